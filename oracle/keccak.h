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

// Same permutation as the loop form it replaced (theta, rho+pi, chi, iota per FIPS-202), with the 25 lanes in locals and the
// rho/pi destinations written out, so that the compiler keeps the state in registers (about 2x faster: the CPU baseline of
// bench.py spends a fifth of its time here).  Pinned by the KATs and the sha3_256 cross-check of tests/test_oracle_primitives.py.
static inline void keccak_f1600(uint64_t A[25]) {
    uint64_t a00 = A[0], a01 = A[1], a02 = A[2], a03 = A[3], a04 = A[4], a05 = A[5], a06 = A[6], a07 = A[7], a08 = A[8], a09 = A[9],
             a10 = A[10], a11 = A[11], a12 = A[12], a13 = A[13], a14 = A[14], a15 = A[15], a16 = A[16], a17 = A[17], a18 = A[18], a19 = A[19],
             a20 = A[20], a21 = A[21], a22 = A[22], a23 = A[23], a24 = A[24];
    for (int round = 0; round < 24; round++) {
        // theta
        const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                       c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
        const uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
        a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
        a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
        a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
        a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
        a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
        // rho + pi: B[y + 5*((2x + 3y) % 5)] = rot(A[x + 5y], KECCAK_ROT[x + 5y])
        const uint64_t b00 = a00,               b10 = rotl64(a01, 1),  b20 = rotl64(a02, 62), b05 = rotl64(a03, 28), b15 = rotl64(a04, 27),
                       b16 = rotl64(a05, 36),   b01 = rotl64(a06, 44), b11 = rotl64(a07, 6),  b21 = rotl64(a08, 55), b06 = rotl64(a09, 20),
                       b07 = rotl64(a10, 3),    b17 = rotl64(a11, 10), b02 = rotl64(a12, 43), b12 = rotl64(a13, 25), b22 = rotl64(a14, 39),
                       b23 = rotl64(a15, 41),   b08 = rotl64(a16, 45), b18 = rotl64(a17, 15), b03 = rotl64(a18, 21), b13 = rotl64(a19, 8),
                       b14 = rotl64(a20, 18),   b24 = rotl64(a21, 2),  b09 = rotl64(a22, 61), b19 = rotl64(a23, 56), b04 = rotl64(a24, 14);
        // chi
        a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
        a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        // iota
        a00 ^= KECCAK_RC[round];
    }
    A[0] = a00; A[1] = a01; A[2] = a02; A[3] = a03; A[4] = a04; A[5] = a05; A[6] = a06; A[7] = a07; A[8] = a08; A[9] = a09;
    A[10] = a10; A[11] = a11; A[12] = a12; A[13] = a13; A[14] = a14; A[15] = a15; A[16] = a16; A[17] = a17; A[18] = a18; A[19] = a19;
    A[20] = a20; A[21] = a21; A[22] = a22; A[23] = a23; A[24] = a24;
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
