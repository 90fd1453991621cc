// ORACLE — TEST INFRASTRUCTURE ONLY (see field.h header).
// Poseidon<BabyBear, CosetMds<_,16>, 16, 5> with 4+4 full and 22 partial rounds, and
// DuplexChallenger<BabyBear, Perm16, 16> (rate = WIDTH), as configured at
// basic/src/bin/valida.rs:360-365,382,397 and basic/tests/test_prover.rs:418-422,439.
// Restated from the p3-poseidon / p3-mds / p3-challenger design [P3-UNVERIFIED; SURVEY App. A 10-12].
// Round constants are an INPUT (480 words): the Rust caller owns the RNG (thread_rng in every
// reference test, Pcg64 from Seeder("validia seed") in the CLI), so they cross the boundary.
#pragma once
#include "field.h"

namespace orc {

struct Poseidon16 {
    static constexpr int WIDTH = 16, HALF_FULL = 4, PARTIAL = 22, ALPHA = 5;
    static constexpr int NUM_CONSTANTS = WIDTH * (2 * HALF_FULL + PARTIAL);  // 480
    uint32_t rc[NUM_CONSTANTS];
    uint32_t mds[WIDTH][WIDTH];  // out[i] = sum_j mds[i][j] * in[j]

    // CosetMds<F,16>::permute: bowers_g_t (inverse twiddles, unscaled, bit-reversed output) ->
    // multiply position i by 31^i -> bowers_g (bit-reversed input, natural output).
    static void coset_mds_apply(uint32_t v[16]) {
        const int N = 16, LOG_N = 4;
        uint32_t root = two_adic_generator(LOG_N), root_inv = inv(root);
        uint32_t fft_tw[8], ifft_tw[8];
        {
            uint32_t a = 1, b = 1;
            uint32_t ft[8], it[8];
            for (int i = 0; i < 8; i++) { ft[i] = a; it[i] = b; a = mul(a, root); b = mul(b, root_inv); }
            for (int i = 0; i < 8; i++) { fft_tw[reverse_bits_len(i, 3)] = ft[i]; ifft_tw[reverse_bits_len(i, 3)] = it[i]; }
        }
        // bowers_g_t: layers log_half_block_size = LOG_N-1 .. 0, DIT butterflies
        for (int lh = LOG_N - 1; lh >= 0; lh--) {
            int half = 1 << lh, nblocks = N >> (lh + 1);
            for (int blk = 0; blk < nblocks; blk++) {
                uint32_t tw = ifft_tw[blk];
                int start = blk << (lh + 1);
                for (int hi = start; hi < start + half; hi++) {
                    int lo = hi + half;
                    uint32_t a = v[hi], b = mul(v[lo], tw);
                    v[hi] = add(a, b); v[lo] = sub(a, b);
                }
            }
        }
        uint32_t w = 1;
        for (int i = 0; i < N; i++) { v[i] = mul(v[i], w); w = mul(w, GEN); }
        // bowers_g: layers 0 .. LOG_N-1, DIF butterflies
        for (int lh = 0; lh < LOG_N; lh++) {
            int half = 1 << lh, nblocks = N >> (lh + 1);
            for (int blk = 0; blk < nblocks; blk++) {
                uint32_t tw = fft_tw[blk];
                int start = blk << (lh + 1);
                for (int hi = start; hi < start + half; hi++) {
                    int lo = hi + half;
                    uint32_t a = v[hi], b = v[lo];
                    v[hi] = add(a, b); v[lo] = mul(sub(a, b), tw);
                }
            }
        }
    }
    void set_default_mds() {
        for (int j = 0; j < WIDTH; j++) {
            uint32_t e[16] = {0};
            e[j] = 1;
            coset_mds_apply(e);
            for (int i = 0; i < WIDTH; i++) mds[i][j] = e[i];
        }
    }
    static uint32_t sbox(uint32_t x) { uint32_t x2 = mul(x, x), x4 = mul(x2, x2); return mul(x4, x); }
    void mds_layer(uint32_t s[16]) const {
        uint32_t o[16];
        for (int i = 0; i < 16; i++) {
            uint64_t acc = 0;
            for (int j = 0; j < 16; j++) acc += ((uint64_t)mds[i][j] * s[j]) % P;
            o[i] = (uint32_t)(acc % P);
        }
        std::memcpy(s, o, sizeof o);
    }
    void permute(uint32_t s[16]) const {
        int r = 0;
        auto constants = [&](int round) { for (int i = 0; i < 16; i++) s[i] = add(s[i], rc[round * 16 + i]); };
        for (int k = 0; k < HALF_FULL; k++, r++) { constants(r); for (int i = 0; i < 16; i++) s[i] = sbox(s[i]); mds_layer(s); }
        for (int k = 0; k < PARTIAL; k++, r++) { constants(r); s[0] = sbox(s[0]); mds_layer(s); }
        for (int k = 0; k < HALF_FULL; k++, r++) { constants(r); for (int i = 0; i < 16; i++) s[i] = sbox(s[i]); mds_layer(s); }
    }
};

// Deterministic stand-in for the caller's RNG (documented in DESIGN.md): SplitMix64 seeded with
// ASCII "valida" = 0x76616c696461, 31-bit rejection sampling (< p), 480 canonical words.
static inline void default_round_constants(uint32_t out[480]) {
    uint64_t state = 0x76616c696461ULL;
    int n = 0;
    while (n < 480) {
        state += 0x9E3779B97F4A7C15ULL;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z = z ^ (z >> 31);
        uint32_t cand = (uint32_t)(z >> 33);
        if (cand < P) out[n++] = cand;
    }
}

struct Challenger {
    const Poseidon16* perm;
    uint32_t state[16];
    std::vector<uint32_t> input, output;
    explicit Challenger(const Poseidon16* p) : perm(p) { std::memset(state, 0, sizeof state); }
    void duplexing() {
        assert(input.size() <= 16);
        for (size_t i = 0; i < input.size(); i++) state[i] = input[i];
        input.clear();
        perm->permute(state);
        output.assign(state, state + 16);
    }
    void observe(uint32_t v) {
        output.clear();
        input.push_back(v);
        if (input.size() == 16) duplexing();
    }
    void observe_digest(const uint32_t d[8]) { for (int i = 0; i < 8; i++) observe(d[i]); }
    uint32_t sample() {
        if (!input.empty() || output.empty()) duplexing();
        uint32_t r = output.back();
        output.pop_back();
        return r;
    }
    Ext5 sample_ext() { Ext5 e; for (int i = 0; i < 5; i++) e.c[i] = sample(); return e; }
    uint32_t sample_bits(int bits) { return sample() & ((1u << bits) - 1); }
    bool check_witness(int bits, uint32_t witness) { observe(witness); return sample_bits(bits) == 0; }
    // GrindingChallenger::grind — the reference searches with rayon find_any (nondeterministic
    // witness); the smallest witness is the deterministic member of that set.
    uint32_t grind(int bits) {
        for (uint32_t w = 0; w < P; w++) {
            Challenger c = *this;
            if (c.check_witness(bits, w)) { bool ok = check_witness(bits, w); assert(ok); (void)ok; return w; }
        }
        assert(false); return 0;
    }
};

}  // namespace orc
