// K8/K9/K10 — kernels of TwoAdicFriPcs::open_multi_batches (derive/src/lib.rs:391-392) and of the
// p3-fri commit phase:
//   inverse denominators 1/(x - z) over a whole bit-reversed coset (one ext5 batch inversion per
//     height and point; the barycentric weights of K8 reuse them on the first-half rows),
//   K8  out-of-domain evaluation p(z) of every column from the first h storage rows (= g*H),
//   K9  reduced openings  ro[i] += alpha^off * (sum_c alpha^c p_c(x_i) - sum_c alpha^c p_c(z)) / (x_i - z),
//   K10 fold_even_odd with the next height's reduced openings added in.
// Ext5 vectors are limb-major (limb l of element i at v[l*cs + i]) so that every access is coalesced.
#include "ctx.h"
#include "devchip.h"
#include "open.h"

namespace {

using bb::E5;

__device__ __forceinline__ uint32_t oroot_pow(const uint32_t* lo, const uint32_t* hi, uint64_t e) {
    e &= ((1ull << VG_LOG_NMAX) - 1);
    return bb::mul(__ldg(lo + (e & (VG_POW_LO - 1))), __ldg(hi + (e >> VG_POW_LO_BITS)));
}
__device__ __forceinline__ E5 ld5(const uint32_t* v, uint64_t cs, uint64_t i) {
    E5 r;
#pragma unroll
    for (int l = 0; l < 5; l++) r.c[l] = v[(uint64_t)l * cs + i];
    return r;
}
__device__ __forceinline__ void st5(uint32_t* v, uint64_t cs, uint64_t i, const E5& x) {
#pragma unroll
    for (int l = 0; l < 5; l++) v[(uint64_t)l * cs + i] = x.c[l];
}

// out[i] = x_i - z,  x_i = s * w_H^bitrev(i)   (storage order of a committed LDE of height H = 2^log_h)
__global__ void __launch_bounds__(256) coset_minus_point_kernel(uint32_t* out, uint64_t H, uint32_t log_h, uint32_t s, E5 z, const uint32_t* lo, const uint32_t* hi) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H) return;
    uint32_t nat = bb::reverse_bits((uint32_t)i, (int)log_h);
    uint32_t x = bb::mul(s, oroot_pow(lo, hi, (uint64_t)nat << (VG_LOG_NMAX - log_h)));
    E5 d = bb::e5_neg(z);
    d.c[0] = bb::add(d.c[0], x);
    st5(out, H, i, d);
}

// K8: partial sums S[c][p] = sum_{rho in block rows} e[rho][c] * x_rho * invden_p[rho] over the first h rows.
constexpr int BARY_COLS = 8, BARY_THREADS = 256;
struct BaryParams {
    const uint32_t* mat; uint64_t mcs; uint64_t h; uint32_t log_H; uint32_t w;
    const uint32_t* invden[2]; uint64_t ics; uint32_t npoints;
    uint32_t s; const uint32_t* lo; const uint32_t* hi;
    uint32_t* partial;       // [gridDim.x][w][npoints][5]
};
__global__ void __launch_bounds__(BARY_THREADS) bary_kernel(BaryParams p) {
    const uint32_t c0 = blockIdx.y * BARY_COLS;
    const uint32_t nc = min((uint32_t)BARY_COLS, p.w - c0);
    E5 acc[BARY_COLS][2];
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++) { acc[c][0] = bb::e5_zero(); acc[c][1] = bb::e5_zero(); }
    for (uint64_t r = (uint64_t)blockIdx.x * BARY_THREADS + threadIdx.x; r < p.h; r += (uint64_t)gridDim.x * BARY_THREADS) {
        uint32_t nat = bb::reverse_bits((uint32_t)r, (int)p.log_H);
        uint32_t x = bb::mul(p.s, oroot_pow(p.lo, p.hi, (uint64_t)nat << (VG_LOG_NMAX - p.log_H)));
        E5 wgt[2];
        for (uint32_t q = 0; q < p.npoints; q++) wgt[q] = bb::e5_mul_base(ld5(p.invden[q], p.ics, r), x);
#pragma unroll
        for (int c = 0; c < BARY_COLS; c++) {
            if ((uint32_t)c < nc) {
                uint32_t e = __ldg(p.mat + (uint64_t)(c0 + c) * p.mcs + r);
                acc[c][0] = bb::e5_add(acc[c][0], bb::e5_mul_base(wgt[0], e));
                if (p.npoints > 1) acc[c][1] = bb::e5_add(acc[c][1], bb::e5_mul_base(wgt[1], e));
            }
        }
    }
    // block reduction through shared memory
    __shared__ uint32_t red[BARY_THREADS / 32][BARY_COLS * 2 * 5];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int l = 0; l < 5; l++) {
                uint32_t v = acc[c][q].c[l];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = bb::add(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) red[wid][(c * 2 + q) * 5 + l] = v;
            }
    __syncthreads();
    if (threadIdx.x < BARY_COLS * 2 * 5) {
        uint32_t v = 0;
        for (int w = 0; w < BARY_THREADS / 32; w++) v = bb::add(v, red[w][threadIdx.x]);
        int c = threadIdx.x / 10, q = (threadIdx.x / 5) % 2, l = threadIdx.x % 5;
        if ((uint32_t)c < nc && (uint32_t)q < p.npoints)
            p.partial[(((uint64_t)blockIdx.x * p.w + c0 + c) * p.npoints + q) * 5 + l] = v;
    }
}

// K9
struct RoParams {
    const uint32_t* mat; uint64_t mcs; uint64_t H; uint32_t w;
    const E5* apow;                   // alpha^c, c < w
    const uint32_t* invden[2]; uint64_t ics; uint32_t npoints;
    E5 sum_y[2]; E5 alpha_off[2];     // per point: sum_c alpha^c y_c and alpha^offset
    uint32_t* ro; uint64_t rcs;       // accumulator, limb-major, height H
};
__global__ void __launch_bounds__(256) reduced_opening_kernel(RoParams p) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.H) return;
    E5 red = bb::e5_zero();
    for (uint32_t c = 0; c < p.w; c++) red = bb::e5_add(red, bb::e5_mul_base(p.apow[c], __ldg(p.mat + (uint64_t)c * p.mcs + i)));
    E5 acc = ld5(p.ro, p.rcs, i);
    for (uint32_t q = 0; q < p.npoints; q++) {
        E5 t = bb::e5_mul(bb::e5_sub(red, p.sum_y[q]), ld5(p.invden[q], p.ics, i));
        acc = bb::e5_add(acc, bb::e5_mul(t, p.alpha_off[q]));
    }
    st5(p.ro, p.rcs, i, acc);
}

// K10: out[i] = (lo + hi)/2 + (beta/2) * g_inv^bitrev(i) * (lo - hi)  (+ add[i])
__global__ void __launch_bounds__(256) fri_fold_kernel(const uint32_t* cur, uint64_t ccs, uint64_t half, uint32_t log_half, E5 half_beta, uint32_t one_half,
                                                      const uint32_t* add, uint64_t acs, uint32_t* out, uint64_t ocs, const uint32_t* lo, const uint32_t* hi) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    E5 a, b;
#pragma unroll
    for (int l = 0; l < 5; l++) { uint2 t = *reinterpret_cast<const uint2*>(cur + (uint64_t)l * ccs + 2 * i); a.c[l] = t.x; b.c[l] = t.y; }
    uint32_t nat = bb::reverse_bits((uint32_t)i, (int)log_half);
    // g_inv^nat, g = two_adic_generator(log_half + 1)
    uint64_t e = (uint64_t)nat << (VG_LOG_NMAX - log_half - 1);
    uint32_t gp = e ? oroot_pow(lo, hi, (1ull << VG_LOG_NMAX) - e) : bb::R1;
    E5 pw = bb::e5_mul_base(half_beta, gp);
    E5 r = bb::e5_add(bb::e5_mul_base(bb::e5_add(a, b), one_half), bb::e5_mul(pw, bb::e5_sub(a, b)));
    if (add) r = bb::e5_add(r, ld5(add, acs, i));
    st5(out, ocs, i, r);
}

// out[i] = sum over blocks of partial[b][i]
__global__ void __launch_bounds__(256) bary_reduce_kernel(const uint32_t* __restrict__ partial, uint32_t nblocks, uint32_t n, uint32_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t acc = 0;
    for (uint32_t b = 0; b < nblocks; b++) acc = bb::add(acc, partial[(uint64_t)b * n + i]);
    out[i] = acc;
}

__global__ void __launch_bounds__(256) gather_words_kernel(const uint32_t* const* ptrs, uint64_t n, uint32_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = *ptrs[i];
}

}  // namespace

int32_t vg_inverse_denominators(vgpu_ctx* ctx, uint32_t log_H, const E5& z, uint32_t* out) {
    uint64_t H = 1ull << log_H;
    KScope ks(ctx, KC_INVDEN, 20.0 * (double)H);
    coset_minus_point_kernel<<<(unsigned)((H + 255) / 256), 256, 0, ctx->stream>>>(out, H, log_H, bb::to_monty(bb::GEN_CANON), z, ctx->root_table.lo, ctx->root_table.hi);
    VG_LAUNCH_CHECK(ctx);
    return vg_ext_batch_inverse(ctx, out, H, H, 1);
}

// p_c(z_q) for every column c and point q (q < npoints <= 2) of a committed LDE (height H = 2h).
int32_t vg_eval_columns(vgpu_ctx* ctx, const vgpu_dmat* lde, uint32_t npoints, const E5* z, const uint32_t* const* invden, std::vector<E5>* ys /* [q][c] */) {
    uint64_t H = lde->h, h = H / 2;
    uint32_t log_H = 0; while ((1ull << log_H) < H) log_H++;
    uint32_t w = (uint32_t)lde->w;
    unsigned bx = (unsigned)std::min<uint64_t>((h + BARY_THREADS - 1) / BARY_THREADS, 8 * (uint64_t)ctx->sm_count);
    unsigned by = (w + BARY_COLS - 1) / BARY_COLS;
    uint32_t* partial = nullptr;
    size_t pn = (size_t)bx * w * npoints * 5;
    VG_TRY(vg_alloc(ctx, (void**)&partial, pn * 4));
    BaryParams p{};
    p.mat = lde->d; p.mcs = lde->col_stride; p.h = h; p.log_H = log_H; p.w = w;
    p.invden[0] = invden[0]; p.invden[1] = npoints > 1 ? invden[1] : invden[0]; p.ics = H; p.npoints = npoints;
    p.s = bb::to_monty(bb::GEN_CANON); p.lo = ctx->root_table.lo; p.hi = ctx->root_table.hi; p.partial = partial;
    {
        KScope ks(ctx, KC_BARY, 4.0 * (double)h * w);
        bary_kernel<<<dim3(bx, by), BARY_THREADS, 0, ctx->stream>>>(p);
    }
    VG_LAUNCH_CHECK(ctx);
    const uint32_t nout = w * npoints * 5;
    uint32_t* reduced = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&reduced, nout * 4));
    bary_reduce_kernel<<<(nout + 255) / 256, 256, 0, ctx->stream>>>(partial, bx, nout, reduced);
    VG_LAUNCH_CHECK(ctx);
    std::vector<uint32_t> hp(nout);
    VG_CUDA(ctx, cudaMemcpyAsync(hp.data(), reduced, nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    vg_free(ctx, partial); vg_free(ctx, reduced);
    // p(z) = -(z^h - s^h) / (h s^h) * S
    uint32_t log_h = log_H - 1;
    uint32_t s = bb::to_monty(bb::GEN_CANON), sh = s;
    for (uint32_t i = 0; i < log_h; i++) sh = bb::sqr(sh);
    uint32_t denom_inv = bb::inv(bb::mul(bb::to_monty((uint32_t)(h % bb::P)), sh));
    ys->assign((size_t)npoints * w, bb::e5_zero());
    for (uint32_t q = 0; q < npoints; q++) {
        E5 zh = bb::e5_exp_pow2(z[q], (int)log_h);
        E5 norm = bb::e5_neg(bb::e5_mul_base(bb::e5_sub_base(zh, sh), denom_inv));
        for (uint32_t c = 0; c < w; c++) {
            E5 S;
            for (int l = 0; l < 5; l++) S.c[l] = hp[((size_t)c * npoints + q) * 5 + l];
            (*ys)[(size_t)q * w + c] = bb::e5_mul(S, norm);
        }
    }
    return 0;
}

int32_t vg_reduced_opening_accumulate(vgpu_ctx* ctx, const vgpu_dmat* lde, const E5* d_apow, uint32_t npoints, const uint32_t* const* invden,
                                      const E5* sum_y, const E5* alpha_off, uint32_t* ro) {
    RoParams p{};
    p.mat = lde->d; p.mcs = lde->col_stride; p.H = lde->h; p.w = (uint32_t)lde->w; p.apow = d_apow;
    p.npoints = npoints; p.ics = lde->h;
    for (uint32_t q = 0; q < npoints; q++) { p.invden[q] = invden[q]; p.sum_y[q] = sum_y[q]; p.alpha_off[q] = alpha_off[q]; }
    p.ro = ro; p.rcs = lde->h;
    {
        KScope ks(ctx, KC_REDUCED_OPENING, (double)lde->h * (4.0 * lde->w + 40.0));
        reduced_opening_kernel<<<(unsigned)((lde->h + 255) / 256), 256, 0, ctx->stream>>>(p);
    }
    VG_LAUNCH_CHECK(ctx);
    return 0;
}

int32_t vg_fri_fold(vgpu_ctx* ctx, const uint32_t* cur, uint64_t n, const E5& beta, const uint32_t* add_or_null, uint32_t* out) {
    uint64_t half = n / 2;
    uint32_t log_half = 0; while ((1ull << log_half) < half) log_half++;
    uint32_t one_half = bb::inv(bb::to_monty(2));
    E5 half_beta = bb::e5_mul_base(beta, one_half);
    KScope ks(ctx, KC_FRI_FOLD, 60.0 * (double)half);
    fri_fold_kernel<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>(cur, n, half, log_half, half_beta, one_half, add_or_null, half, out, half, ctx->root_table.lo, ctx->root_table.hi);
    VG_LAUNCH_CHECK(ctx);
    return 0;
}

int32_t vg_gather_words(vgpu_ctx* ctx, const std::vector<const uint32_t*>& ptrs, std::vector<uint32_t>* out) {
    size_t n = ptrs.size();
    out->assign(n, 0);
    if (!n) return 0;
    const uint32_t** dptr = nullptr; uint32_t* dout = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&dptr, n * sizeof(void*)));
    VG_TRY(vg_alloc(ctx, (void**)&dout, n * 4));
    VG_CUDA(ctx, cudaMemcpyAsync(dptr, ptrs.data(), n * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    gather_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dptr, n, dout);
    VG_LAUNCH_CHECK(ctx);
    VG_CUDA(ctx, cudaMemcpyAsync(out->data(), dout, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    vg_free(ctx, dptr); vg_free(ctx, dout);
    return 0;
}
