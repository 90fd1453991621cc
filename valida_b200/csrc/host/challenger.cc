#include "challenger.h"
#include "../ctx.h"
#include <cstring>

namespace vgh {

// CosetMds<F,16>::default(): unscaled inverse Bowers network, multiply slot i by 31^i, forward Bowers
// network (p3-mds coset_mds).  Realised here as the 16x16 matrix obtained by pushing the unit vectors
// through the transform written as explicit sums (a different formulation from the butterfly network):
//   u_q = sum_j v_j * w^(-j*rev(q))        (inverse DFT, bit-reversed output slots, no 1/N)
//   y_q = u_q * 31^q
//   out_k = sum_q y_q * w^(rev(q)*k)       (forward DFT reading bit-reversed input slots)
static void default_mds(uint32_t mds[16][16]) {
    uint32_t w = bb::two_adic_generator_monty(4), winv = bb::inv(w), g = bb::to_monty(31);
    uint32_t gp[16];
    gp[0] = bb::R1;
    for (int i = 1; i < 16; i++) gp[i] = bb::mul(gp[i - 1], g);
    for (int k = 0; k < 16; k++)
        for (int j = 0; j < 16; j++) {
            uint32_t acc = 0;
            for (int q = 0; q < 16; q++) {
                int rq = (int)bb::reverse_bits((uint32_t)q, 4);
                uint32_t a = bb::pow(winv, (uint64_t)(j * rq) % 16);
                uint32_t b = bb::pow(w, (uint64_t)(rq * k) % 16);
                acc = bb::add(acc, bb::mul(bb::mul(a, gp[q]), b));
            }
            mds[k][j] = acc;
        }
}

void Poseidon16::set(const uint32_t rc_canonical[480], const uint32_t* mds_canonical_or_null) {
    for (int i = 0; i < 480; i++) rc[i] = bb::to_monty(rc_canonical[i] % bb::P);
    if (mds_canonical_or_null) {
        for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) mds[i][j] = bb::to_monty(mds_canonical_or_null[i * 16 + j] % bb::P);
    } else {
        default_mds(mds);
    }
}

static inline uint32_t sbox5(uint32_t x) { uint32_t x2 = bb::sqr(x), x4 = bb::sqr(x2); return bb::mul(x4, x); }

void Poseidon16::permute(uint32_t s[16]) const {
    // dense 16 x 16 product with lazy sums: four raw 64-bit products (each < p^2 < 2^62) per Montgomery reduction
    auto mds_layer = [&]() {
        uint32_t o[16];
        for (int i = 0; i < 16; i++) {
            uint32_t acc = 0;
            for (int j = 0; j < 16; j += 4) {
                const uint64_t t = (uint64_t)mds[i][j] * s[j] + (uint64_t)mds[i][j + 1] * s[j + 1] + (uint64_t)mds[i][j + 2] * s[j + 2] + (uint64_t)mds[i][j + 3] * s[j + 3];
                acc = bb::add(acc, bb::monty_reduce64(t));
            }
            o[i] = acc;
        }
        std::memcpy(s, o, sizeof o);
    };
    int round = 0;
    for (int phase = 0; phase < 3; phase++) {
        int n = phase == 1 ? 22 : 4;
        for (int k = 0; k < n; k++, round++) {
            for (int i = 0; i < 16; i++) s[i] = bb::add(s[i], rc[round * 16 + i]);
            if (phase == 1) s[0] = sbox5(s[0]);
            else for (int i = 0; i < 16; i++) s[i] = sbox5(s[i]);
            mds_layer();
        }
    }
}

void Challenger::duplexing() {
    for (size_t i = 0; i < input.size(); i++) state[i] = input[i];
    input.clear();
    perm->permute(state);
    output.assign(state, state + 16);
}
void Challenger::observe(uint32_t v) {
    output.clear();
    input.push_back(v);
    if (input.size() == 16) duplexing();
}
void Challenger::observe_digest_canonical(const uint32_t d[8]) { for (int i = 0; i < 8; i++) observe(bb::to_monty(d[i])); }
uint32_t Challenger::sample() {
    if (!input.empty() || output.empty()) duplexing();
    uint32_t r = output.back();
    output.pop_back();
    return r;
}
bb::E5 Challenger::sample_ext() { bb::E5 e; for (int i = 0; i < 5; i++) e.c[i] = sample(); return e; }
uint32_t Challenger::sample_bits(int bits) { return bb::from_monty(sample()) & ((1u << bits) - 1); }
bool Challenger::check_witness(int bits, uint32_t w) { observe(w); return sample_bits(bits) == 0; }
uint32_t Challenger::grind(int bits) {
    for (uint32_t w = 0; w < bb::P; w++) {
        Challenger c = *this;
        uint32_t wm = bb::to_monty(w);
        if (c.check_witness(bits, wm)) { check_witness(bits, wm); return wm; }
    }
    return 0;
}

}  // namespace vgh

extern "C" int32_t vgpu_set_challenger(vgpu_ctx* ctx, const uint32_t round_constants[480], const uint32_t* mds_16x16_or_null) {
    if (!ctx || !round_constants) return -1;
    std::memcpy(ctx->poseidon_rc, round_constants, sizeof ctx->poseidon_rc);
    if (ctx->d_poseidon) { vg_free(ctx, ctx->d_poseidon); ctx->d_poseidon = nullptr; }
    if (mds_16x16_or_null) std::memcpy(ctx->poseidon_mds, mds_16x16_or_null, sizeof ctx->poseidon_mds);
    ctx->challenger_set = true;
    ctx->poseidon_has_mds = mds_16x16_or_null != nullptr;
    return 0;
}
