// K6 + K7 — fused constraint / quotient sweep with the chunk split, one kernel per chip.
// Replaces quotient() / quotient_values() (machine/src/quotient.rs:18-238), ProverConstraintFolder
// (machine/src/folding_builder.rs:32-125), eval_permutation_constraints (machine/src/chip.rs:210-289)
// and p3-uni-stark's decompose_and_flatten / ZerofierOnCoset for log_quotient_degree = 1.
//
// The reference gathers LDE rows through a bit-reversed view element by element; here one thread owns
// one STORAGE row of the committed (bit-reversed) LDEs; lanes (2r, 2r+1) hold the natural rows
// (j, j+h) with j = bitrev(r): x and -x.  That pair is exactly what the even/odd chunk split needs,
// so the quotient values never touch HBM: each lane folds all constraints of its row, divides by Z_H,
// the pair exchanges its ext5 value with one shuffle, and the even / odd lane writes the even / odd
// chunk limbs.  Every trace load and every chunk store is a fully coalesced 4-byte-per-lane access
// (the chunk matrix is left in bit-reversed row order; the quotient commit's iNTT reads it that way).
// ncu on the first version (pair per thread, natural-order scatter): 3.2x the algorithmic DRAM reads
// and 9x the writes; see profiles/r01_summary.md.
// alpha-folding: acc = sum_i c_i * alpha^(N-1-i) with precomputed powers — the same value as the
// reference's Horner recurrence acc = acc*alpha + c_i, at 5 instead of 25 multiplications for the
// base-field constraints.  The sum is accumulated LAZILY (bb::Lazy5: raw 64-bit products, one IMAD.WIDE
// per limb and constraint, a fold every fourth) and reduced once per row; the powers sit in the kernel
// parameters (constant bank), so a base-field constraint costs ~7 instructions instead of 50.
// The selectors need 1/((x-1)(x-g^-1)) on both rows of a pair: the product over the pair is symmetric,
// vg_selector_inverses() inverts it for every pair of a height with the Montgomery batch trick (one
// Fermat inversion per 8 pairs) instead of one Fermat inversion (~46 multiplications) per row.
#include "ctx.h"
#include "devchip.h"
#include "airs.cuh"
#include <cstring>
#include <cstdlib>
#include <memory>

namespace {

using bb::E5;
using air::F;

constexpr uint32_t Q_MAX_CONSTRAINTS = 128;   // bitwise: 88 base + interactions + 3

struct QParams {
    // Base pointers are VIRTUAL: base + (global storage row) is the element, whether the matrix is whole or this rank's row
    // shard (then base = shard - first row).  The *_n bases serve the "next" rows: natural row i + 2 of every row of a shard
    // lies in ONE other rank's shard (rows of a shard share i mod comm_size), read through its peer pointer over NVLink.
    const uint32_t* main; const uint32_t* main_n; uint64_t mcs;
    const uint32_t* prep; const uint32_t* prep_n; uint64_t pcs;
    const uint32_t* perm; const uint32_t* perm_n; uint64_t qcs;
    uint32_t* out; uint64_t ocs;            // h x 10 chunk matrix (virtual base: + chunk row)
    const uint32_t* selinv;                 // selinv[r] = 1 / ((x-1)(x-glast)(-x-1)(-x-glast)), x = s * w^bitrev(r)  (virtual base: + pair)
    uint32_t log_h;
    uint64_t row_begin, row_end;            // storage rows of the LDE swept by this launch (a rank's range when the sweep is split)
    uint32_t s;                             // coset shift (Montgomery)
    uint32_t glast;                         // g_subgroup^-1
    uint32_t zh[2], zinv[2];                // Z_H on even / odd natural rows, and inverses
    uint32_t odd_scale;                     // 1 / (2 s)
    uint32_t half;                          // 1 / 2
    E5 cumsum;
    const uint32_t* root_lo; const uint32_t* root_hi;
    uint32_t apow[Q_MAX_CONSTRAINTS][5];    // apow[i] = alpha^(N-1-i)
    DevChip chip;                           // interaction descriptors + LogUp randomness: read through the constant bank (uniform loads),
                                            // not through dependent global loads (ncu r1b: 27-54 % of the stall samples sat on those);
                                            // staging them in shared memory instead measured 6 % slower (9.08 vs 8.57 ms per proof)
};

struct DevBuilder {
    using V = air::F;
    const uint32_t* lrow; const uint32_t* nrow; uint64_t cs;   // pointers already offset to the row
    F first, last, trans;
    const uint32_t (*apow)[5]; uint32_t idx; bb::Lazy5 acc;
    __device__ __forceinline__ F L(int c) const { return F{__ldg(lrow + (uint64_t)c * cs)}; }
    __device__ __forceinline__ F N(int c) const { return F{__ldg(nrow + (uint64_t)c * cs)}; }
    __device__ __forceinline__ E5 pw() const { E5 a; for (int l = 0; l < 5; l++) a.c[l] = apow[idx][l]; return a; }
    __device__ __forceinline__ void z(F x) { acc.fma_base(pw(), x.v); idx++; }
    __device__ __forceinline__ void z_ext(const E5& x) { acc.fma_ext(pw(), x, bb::e5_dbl(x)); idx++; }
};

__device__ __forceinline__ uint32_t qroot_pow(const QParams& p, uint64_t e) {
    e &= ((1ull << VG_LOG_NMAX) - 1);
    return bb::mul(__ldg(p.root_lo + (e & (VG_POW_LO - 1))), __ldg(p.root_hi + (e >> VG_POW_LO_BITS)));
}
__device__ __forceinline__ uint32_t dev_pair_col(const DevPairCol& pc, const uint32_t* mrow, uint64_t mcs, const uint32_t* prow, uint64_t pcs) {
    uint32_t v = pc.constant;
    for (uint32_t t = 0; t < pc.n_terms; t++) {
        uint32_t x = pc.is_prep[t] ? __ldg(prow + (uint64_t)pc.column[t] * pcs) : __ldg(mrow + (uint64_t)pc.column[t] * mcs);
        v = bb::add(v, bb::mul(x, pc.weight[t]));
    }
    return v;
}
__device__ __forceinline__ E5 load_e5(const uint32_t* row, uint64_t cs, uint32_t m) {
    E5 r;
#pragma unroll
    for (int l = 0; l < 5; l++) r.c[l] = __ldg(row + (uint64_t)(5 * m + l) * cs);
    return r;
}

// MINB: resident-CTA target (register cap) of the variant.  The sweep is latency bound (ncu r1b: issue slots 43-57 % busy at
// 5 CTAs per SM), so more resident warps beat fewer spills: measured 11.4 / 8.9 / 8.6 ms per proof at 2-4 / 6 / 8 CTAs per SM.
template <int CHIP, int MINB>
__global__ void __launch_bounds__(128, MINB) quotient_kernel(const __grid_constant__ QParams p) {
    const uint64_t h = 1ull << p.log_h, H = 2 * h;
    const uint64_t rho_raw = p.row_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;     // storage row of the committed LDEs
    const bool active = rho_raw < p.row_end;
    const uint64_t rho = active ? rho_raw : p.row_begin + (rho_raw & 1);           // idle lanes shadow the first pair of the range (shuffles need every lane)
    const uint32_t e = (uint32_t)(rho & 1);                                        // 0: x = +x0 (natural row j), 1: x = -x0 (natural row j + h)
    const uint64_t r = rho >> 1;
    const uint32_t j = bb::reverse_bits((uint32_t)r, (int)p.log_h);
    // next row: natural (i + 2) mod 2h  ->  pair bitrev((j + 2) mod h), element e ^ carry
    const uint64_t t = (uint64_t)j + 2;
    const uint32_t a = (uint32_t)(t & (h - 1));
    const uint32_t swap = (uint32_t)((t >> p.log_h) & 1);
    const uint64_t nrow = 2 * (uint64_t)bb::reverse_bits(a, (int)p.log_h) + (e ^ swap);
    const uint32_t x0 = bb::mul(p.s, qroot_pow(p, (uint64_t)j << (VG_LOG_NMAX - p.log_h - 1)));
    const uint32_t x = e ? bb::neg(x0) : x0;
    // selectors 1/(x-1), 1/(x-glast): both lanes of a pair invert the same symmetric product
    uint32_t inv_first, inv_last;
    {
        const uint32_t nx0 = bb::neg(x0);
        const uint32_t d0 = bb::sub(x0, bb::R1), d1 = bb::sub(x0, p.glast), d2 = bb::sub(nx0, bb::R1), d3 = bb::sub(nx0, p.glast);
        const uint32_t p01 = bb::mul(d0, d1), p23 = bb::mul(d2, d3);
        const uint32_t all = __ldg(p.selinv + r);
        const uint32_t i01 = bb::mul(all, p23), i23 = bb::mul(all, p01);
        inv_first = e ? bb::mul(i23, d3) : bb::mul(i01, d1);
        inv_last = e ? bb::mul(i23, d2) : bb::mul(i01, d0);
    }
    const DevChip& chip = p.chip;
    const uint32_t k = chip.n_interactions;
    const uint32_t parity = (uint32_t)(((uint64_t)j + (e ? h : 0)) & 1);
    const uint32_t zh = p.zh[parity];
    DevBuilder b;
    b.lrow = p.main + rho; b.nrow = p.main_n + nrow; b.cs = p.mcs;
    b.first = F{bb::mul(zh, inv_first)};
    b.last = F{bb::mul(zh, inv_last)};
    b.trans = F{bb::sub(x, p.glast)};
    b.apow = p.apow; b.idx = 0; b.acc.init();
    air::eval_chip<CHIP>(b);
    {   // eval_permutation_constraints
        const uint32_t* ql = p.perm + rho; const uint32_t* qn = p.perm_n + nrow;
        const uint32_t* pl = p.prep ? p.prep + rho : nullptr; const uint32_t* pn = p.prep ? p.prep_n + nrow : nullptr;
        const E5 phi_local = load_e5(ql, p.qcs, k), phi_next = load_e5(qn, p.qcs, k);
        E5 rhs = bb::e5_zero(), phi0 = bb::e5_zero();
        for (uint32_t m = 0; m < k; m++) {
            const DevInteraction& it = chip.interactions[m];
            bb::Lazy5 ra; ra.init();
            for (uint32_t f = 0; f < it.n_fields; f++) ra.fma_base(chip.betas[f], dev_pair_col(it.fields[f], b.lrow, p.mcs, pl, p.pcs));
            const E5 rlc = bb::e5_add(it.alpha, ra.value());
            const E5 pm_l = load_e5(ql, p.qcs, m), pm_n = load_e5(qn, p.qcs, m);
            b.z_ext(bb::e5_sub_base(bb::e5_mul(rlc, pm_l), bb::R1));
            const uint32_t mult_l = dev_pair_col(it.count, b.lrow, p.mcs, pl, p.pcs), mult_n = dev_pair_col(it.count, b.nrow, p.mcs, pn, p.pcs);
            const E5 tl = bb::e5_mul_base(pm_l, mult_l), tn = bb::e5_mul_base(pm_n, mult_n);
            if (it.is_send) { phi0 = bb::e5_add(phi0, tl); rhs = bb::e5_add(rhs, tn); }
            else { phi0 = bb::e5_sub(phi0, tl); rhs = bb::e5_sub(rhs, tn); }
        }
        b.z_ext(bb::e5_mul_base(bb::e5_sub(bb::e5_sub(phi_next, phi_local), rhs), b.trans.v));
        b.z_ext(bb::e5_mul_base(bb::e5_sub(phi_local, phi0), b.first.v));
        b.z_ext(bb::e5_mul_base(bb::e5_sub(phi_local, p.cumsum), b.last.v));
    }
    const E5 q = bb::e5_mul_base(b.acc.value(), p.zinv[parity]);
    // decompose_and_flatten across the lane pair: even = (q(x) + q(-x))/2, odd = (q(x) - q(-x)) / (2 s g^j)
    E5 other;
#pragma unroll
    for (int l = 0; l < 5; l++) other.c[l] = __shfl_xor_sync(0xffffffffu, q.c[l], 1);
    E5 outv;
    if (e == 0) {
        outv = bb::e5_mul_base(bb::e5_add(q, other), p.half);
    } else {
        const uint32_t ginv_j = qroot_pow(p, (1ull << VG_LOG_NMAX) - ((uint64_t)j << (VG_LOG_NMAX - p.log_h - 1)));
        outv = bb::e5_mul_base(bb::e5_sub(other, q), bb::mul(p.odd_scale, ginv_j));
    }
    if (active) {
        // chunk row j is written at row r = bitrev(j): the chunk matrix leaves this kernel in bit-reversed row order
        // (coalesced), which the commit's inverse transform consumes directly
#pragma unroll
        for (int l = 0; l < 5; l++) p.out[(uint64_t)(5 * e + l) * p.ocs + r] = outv.c[l];
    }
}

// selinv[r] = 1 / ((x0-1)(x0-glast)(-x0-1)(-x0-glast)), x0 = s * w_2h^bitrev(r), for the pairs [begin, begin + count).
// Montgomery batch trick over SEL_BATCH pairs per thread, pairs of one thread a grid apart (coalesced).
constexpr int SEL_BATCH = 8;
__global__ void __launch_bounds__(256) selector_inverse_kernel(uint32_t* __restrict__ out, uint64_t begin, uint64_t count, uint32_t log_h, uint32_t s, uint32_t glast,
                                                               const uint32_t* __restrict__ root_lo, const uint32_t* __restrict__ root_hi) {
    const uint64_t stride = (count + SEL_BATCH - 1) / SEL_BATCH;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    uint32_t v[SEL_BATCH], pref[SEL_BATCH];
    uint32_t acc = bb::R1;
#pragma unroll
    for (int i = 0; i < SEL_BATCH; i++) {
        const uint64_t r = begin + t + (uint64_t)i * stride;
        v[i] = bb::R1;
        if (t + (uint64_t)i * stride < count) {
            const uint32_t j = bb::reverse_bits((uint32_t)r, (int)log_h);
            uint64_t e = ((uint64_t)j << (VG_LOG_NMAX - log_h - 1)) & ((1ull << VG_LOG_NMAX) - 1);
            const uint32_t x0 = bb::mul(s, bb::mul(__ldg(root_lo + (e & (VG_POW_LO - 1))), __ldg(root_hi + (e >> VG_POW_LO_BITS))));
            const uint32_t nx0 = bb::neg(x0);
            v[i] = bb::mul(bb::mul(bb::sub(x0, bb::R1), bb::sub(x0, glast)), bb::mul(bb::sub(nx0, bb::R1), bb::sub(nx0, glast)));
        }
        pref[i] = acc;
        acc = bb::mul(acc, v[i]);
    }
    uint32_t inv = bb::inv(acc);
#pragma unroll
    for (int i = SEL_BATCH - 1; i >= 0; i--) {
        if (t + (uint64_t)i * stride < count) out[begin + t + (uint64_t)i * stride] = bb::mul(inv, pref[i]);
        inv = bb::mul(inv, v[i]);
    }
}

struct CountBuilder {
    using V = air::F;
    F first{0}, last{0}, trans{0};
    uint32_t n = 0;
    BB_HD F L(int) const { return F{0}; }
    BB_HD F N(int) const { return F{0}; }
    BB_HD void z(F) { n++; }
};
template <int CHIP> uint32_t count_base() { CountBuilder c; air::eval_chip<CHIP>(c); return c.n; }

template <int CHIP> void launch(const QParams& p, uint64_t h, cudaStream_t st) {
    (void)h;
    static const int minb = [] { const char* e = getenv("VGPU_QUOTIENT_MINB"); return e ? atoi(e) : 8; }();   // tuning knob (profiles/)
    const unsigned grid = (unsigned)((p.row_end - p.row_begin + 127) / 128);
    if (minb == 6) quotient_kernel<CHIP, 6><<<grid, 128, 0, st>>>(p);
    else quotient_kernel<CHIP, 8><<<grid, 128, 0, st>>>(p);
}

}  // namespace

uint32_t vg_chip_base_constraints(uint32_t chip_id) {
    switch (chip_id) {
        case 0: return count_base<0>(); case 3: return count_base<3>(); case 4: return count_base<4>(); case 5: return count_base<5>();
        case 7: return count_base<7>(); case 8: return count_base<8>(); case 9: return count_base<9>(); case 10: return count_base<10>();
        case 11: return count_base<11>(); case 13: return count_base<13>(); default: return 0;
    }
}

extern "C" int32_t vgpu_quotient(vgpu_ctx* ctx, const vgpu_chip_desc* chip, uint32_t log_degree, const vgpu_dmat* prep_lde,
                                 const vgpu_dmat* main_lde, const vgpu_dmat* perm_lde, const uint32_t cumulative_sum[5],
                                 const uint32_t perm_challenges[15], const uint32_t alpha[5], vgpu_dmat** out_chunks) {
    if (!chip || !main_lde || !perm_lde || !out_chunks) VG_FAIL(ctx, "quotient: null argument");
    VG_TRY(vg_enter(ctx));
    const uint64_t h = 1ull << log_degree;
    if (main_lde->gh != 2 * h || perm_lde->gh != 2 * h) VG_FAIL(ctx, "quotient: LDE height must be 2 * 2^log_degree");
    if (main_lde->gw != chip->width || perm_lde->gw != 5 * (chip->n_interactions + 1)) VG_FAIL(ctx, "quotient: LDE width does not match the chip");
    if (chip->chip_id >= VGPU_NUM_CHIPS) VG_FAIL(ctx, "quotient: unknown chip id %u", chip->chip_id);
    // split proof: the committed LDEs of a tall chip are row shards (all three the same run of rows), of a short chip whole
    const bool split = main_lde->dist == VG_ROWS;
    if ((perm_lde->dist == VG_ROWS) != split || (prep_lde && (prep_lde->dist == VG_ROWS) != split)) VG_FAIL(ctx, "quotient: the LDEs are not distributed alike");
    if (main_lde->dist == VG_COLS) VG_FAIL(ctx, "quotient: column shares are internal to a commit");
    // alpha powers for N = base + k + 3 constraints
    const uint32_t N = vg_chip_base_constraints(chip->chip_id) + chip->n_interactions + 3;
    if (N > Q_MAX_CONSTRAINTS) { VG_FAIL(ctx, "quotient: %u constraints exceed the parameter table (%u)", N, Q_MAX_CONSTRAINTS); }
    auto pp = std::make_unique<QParams>();
    QParams& p = *pp;
    std::memset(&p, 0, sizeof p);
    VG_TRY(vg_build_devchip(ctx, chip, perm_challenges, &p.chip));
    E5 al; for (int i = 0; i < 5; i++) al.c[i] = bb::to_monty(alpha[i] % bb::P);
    { E5 a = bb::e5_one(); for (uint32_t i = 0; i < N; i++) { for (int l = 0; l < 5; l++) p.apow[N - 1 - i][l] = a.c[l]; a = bb::e5_mul(a, al); } }
    p.row_begin = split ? main_lde->row0 : 0;
    p.row_end = split ? main_lde->row0 + main_lde->h : 2 * h;
    if ((p.row_begin & 1) || (p.row_end & 1)) VG_FAIL(ctx, "quotient: a row shard must hold whole (x, -x) pairs");
    int next_rank = ctx->comm_rank;
    if (split) {   // rank holding natural rows i + 2 of this shard's rows: reverse_bits((reverse_bits(rank) + 2) mod G)
        int lg = 0; while ((1 << lg) < ctx->comm_size) lg++;
        const uint32_t rho = bb::reverse_bits((uint32_t)ctx->comm_rank, lg);
        next_rank = (int)bb::reverse_bits((rho + 2) & (uint32_t)(ctx->comm_size - 1), lg);
    }
    // virtual bases of a matrix: local rows, and the rows of `next_rank`
    auto base_of = [&](const vgpu_dmat* m) { return m->d - (split ? m->row0 : 0); };
    auto next_of = [&](const vgpu_dmat* m) -> const uint32_t* {
        if (!split || next_rank == ctx->comm_rank) return base_of(m);
        if (!m->symm) return nullptr;
        return vg_peer_ptr(ctx, m->d, next_rank) - (uint64_t)next_rank * m->h;
    };
    vgpu_dmat* out = nullptr;
    VG_TRY(split ? vg_dmat_alloc_dist(ctx, VG_ROWS, h, 10, false, &out) : vg_dmat_alloc(ctx, h, 10, &out));
    out->bitrev_rows = true;
    p.main = base_of(main_lde); p.main_n = next_of(main_lde); p.mcs = main_lde->col_stride;
    p.prep = prep_lde ? base_of(prep_lde) : nullptr; p.prep_n = prep_lde ? next_of(prep_lde) : nullptr; p.pcs = prep_lde ? prep_lde->col_stride : 0;
    p.perm = base_of(perm_lde); p.perm_n = next_of(perm_lde); p.qcs = perm_lde->col_stride;
    if (!p.main_n || !p.perm_n || (prep_lde && !p.prep_n)) { vgpu_dmat_free(out); VG_FAIL(ctx, "quotient: a row shard read by a peer must live in the symmetric heap"); }
    p.out = out->d - p.row_begin / 2; p.ocs = out->col_stride;
    p.log_h = log_degree;
    p.s = bb::to_monty(bb::GEN_CANON);
    uint32_t g_sub = bb::two_adic_generator_monty((int)log_degree);
    p.glast = bb::inv(g_sub);
    uint32_t s_pow_n = p.s;
    for (uint32_t i = 0; i < log_degree; i++) s_pow_n = bb::sqr(s_pow_n);
    p.zh[0] = bb::sub(s_pow_n, bb::R1);
    p.zh[1] = bb::sub(bb::neg(s_pow_n), bb::R1);
    p.zinv[0] = bb::inv(p.zh[0]); p.zinv[1] = bb::inv(p.zh[1]);
    p.half = bb::inv(bb::to_monty(2));
    p.odd_scale = bb::mul(p.half, bb::inv(p.s));
    for (int i = 0; i < 5; i++) p.cumsum.c[i] = bb::to_monty(cumulative_sum[i] % bb::P);
    p.root_lo = ctx->root_table.lo; p.root_hi = ctx->root_table.hi;
    const uint64_t pb = p.row_begin / 2, pc = (p.row_end - p.row_begin) / 2;     // pairs swept here
    uint32_t* selinv = nullptr;
    { int32_t rc = vg_alloc(ctx, (void**)&selinv, pc * 4); if (rc) { vgpu_dmat_free(out); return rc; } }
    p.selinv = selinv - pb;
    KScope* ks = new KScope(ctx, KC_QUOTIENT, 4.0 * (double)(p.row_end - p.row_begin) * (main_lde->gw + perm_lde->gw + (prep_lde ? prep_lde->gw : 0)) + 20.0 * (double)(p.row_end - p.row_begin));
    {
        const uint64_t stride = (pc + SEL_BATCH - 1) / SEL_BATCH;
        selector_inverse_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, ctx->stream>>>(selinv - pb, pb, pc, log_degree, p.s, p.glast, p.root_lo, p.root_hi);
        ctx->launches++;
    }
    switch (chip->chip_id) {
        case 0: launch<0>(p, h, ctx->stream); break;   case 1: launch<1>(p, h, ctx->stream); break;
        case 2: launch<2>(p, h, ctx->stream); break;   case 3: launch<3>(p, h, ctx->stream); break;
        case 4: launch<4>(p, h, ctx->stream); break;   case 5: launch<5>(p, h, ctx->stream); break;
        case 6: launch<6>(p, h, ctx->stream); break;   case 7: launch<7>(p, h, ctx->stream); break;
        case 8: launch<8>(p, h, ctx->stream); break;   case 9: launch<9>(p, h, ctx->stream); break;
        case 10: launch<10>(p, h, ctx->stream); break; case 11: launch<11>(p, h, ctx->stream); break;
        case 12: launch<12>(p, h, ctx->stream); break; case 13: launch<13>(p, h, ctx->stream); break;
    }
    delete ks;
    vg_free(ctx, selinv);
    VG_LAUNCH_CHECK(ctx);
    *out_chunks = out;
    return 0;
}
