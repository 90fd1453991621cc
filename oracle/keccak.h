// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// Keccak-256 with the legacy 0x01 domain padding, as used by p3-keccak's Keccak256Hash over
// tiny-keccak 2.0.2 `Keccak::v256` (reference call sites: basic/src/bin/valida.rs:26,367-371).
// Restated from the published Keccak-f[1600] specification (FIPS-202 permutation, pre-SHA-3 padding).
// Pinned by tests/test_oracle_keccak.py: KATs keccak256("")/("abc") and, with pad byte 0x06,
// bit-equality with hashlib.sha3_256 over random lengths (same permutation, same sponge).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {

static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void keccak_f1600(uint64_t A[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
        // rho + pi: B[y, 2x+3y] = rot(A[x,y], r[x,y]);  index = x + 5*y
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                int src = x + 5 * y;
                int dst = y + 5 * ((2 * x + 3 * y) % 5);
                B[dst] = rotl64(A[src], KECCAK_ROT[src]);
            }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= KECCAK_RC[round];
    }
}

// Sponge with rate 136 bytes, 32-byte output, selectable domain pad byte (0x01 Keccak, 0x06 SHA-3).
static inline void keccak256_pad(const uint8_t* in, size_t len, uint8_t out[32], uint8_t pad) {
    uint64_t A[25];
    std::memset(A, 0, sizeof A);
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; i++) { uint64_t w; std::memcpy(&w, in + 8 * i, 8); A[i] ^= w; }
        keccak_f1600(A);
        in += rate; len -= rate;
    }
    uint8_t block[136];
    std::memset(block, 0, sizeof block);
    std::memcpy(block, in, len);
    block[len] ^= pad;
    block[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) { uint64_t w; std::memcpy(&w, block + 8 * i, 8); A[i] ^= w; }
    keccak_f1600(A);
    std::memcpy(out, A, 32);
}
static inline void keccak256(const uint8_t* in, size_t len, uint8_t out[32]) { keccak256_pad(in, len, out, 0x01); }

}  // namespace orc
