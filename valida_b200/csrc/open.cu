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
__global__ void __launch_bounds__(256) coset_minus_point_kernel(uint32_t* out, uint64_t H, uint64_t begin, uint64_t count, uint32_t log_h, uint32_t s, E5 z, const uint32_t* lo, const uint32_t* hi) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    i += begin;
    uint32_t nat = bb::reverse_bits((uint32_t)i, (int)log_h);
    uint32_t x = bb::mul(s, oroot_pow(lo, hi, (uint64_t)nat << (VG_LOG_NMAX - log_h)));
    E5 d = bb::e5_neg(z);
    d.c[0] = bb::add(d.c[0], x);
    st5(out, H, i, d);
}

// K8: out-of-domain evaluation.  With x/(x - z) = 1 + z/(x - z) the barycentric sum splits into
//     S_c(z) = sum_i p_c(x_i) + z * sum_i p_c(x_i) / (x_i - z)
// so the kernel needs no domain points at all: per column a base-field sum and, per point, an ext5 dot product with
// the inverse denominators.  Dot products are accumulated LAZILY: raw 64-bit products of Montgomery words
// (< 2^62 each) are added three at a time into a 64-bit accumulator which is then folded with
// hi*2^32 + lo == hi*R1 + lo (mod p) — one IMAD.WIDE per term instead of a full modular multiply-add.
// A CTA stages the inverse denominators of a tile of rows in shared memory once and every warp sweeps that tile
// for its own BARY_COLS columns, so the denominators are read from HBM once per group of 32 columns.
constexpr int BARY_COLS = 2, BARY_WARPS = 16, BARY_THREADS = 32 * BARY_WARPS, BARY_TILE = 1024, BARY_GROUP = BARY_COLS * BARY_WARPS;
constexpr int BARY_OUT = 11;                       // per column: 2 points x 5 limbs, then the plain column sum
struct BaryParams {
    const uint32_t* mat; uint64_t mcs; uint64_t h; uint32_t w;   // h rows starting at row_begin
    uint64_t row_begin;
    const uint32_t* invden[2]; uint64_t ics; uint32_t npoints;
    uint32_t* partial;       // [gridDim.x][w][BARY_OUT]
};
__device__ __forceinline__ uint64_t lazy_fold(uint64_t a) { return (a & 0xffffffffull) + (a >> 32) * (uint64_t)bb::R1; }
__global__ void __launch_bounds__(BARY_THREADS, 1) bary_kernel(BaryParams p) {
    __shared__ uint32_t dsm[2][5][BARY_TILE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c0 = blockIdx.y * BARY_GROUP + wid * BARY_COLS;
    const uint32_t* col[BARY_COLS];
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++) col[c] = p.mat + (uint64_t)min(c0 + c, p.w - 1) * p.mcs;   // surplus columns shadow the last one (never written)
    uint64_t acc[BARY_COLS][2][5];
    uint64_t sum[BARY_COLS];
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++) { sum[c] = 0; for (int q = 0; q < 2; q++) for (int l = 0; l < 5; l++) acc[c][q][l] = 0; }
    const uint64_t ntiles = (p.h + BARY_TILE - 1) / BARY_TILE;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t rows = (uint32_t)min((uint64_t)BARY_TILE, p.h - t * BARY_TILE);
        const uint64_t row0 = p.row_begin + t * BARY_TILE;
        __syncthreads();
        for (uint32_t q = 0; q < p.npoints; q++)
#pragma unroll
            for (int l = 0; l < 5; l++)
                for (uint32_t r = threadIdx.x; r < rows; r += BARY_THREADS) dsm[q][l][r] = __ldg(p.invden[q] + (uint64_t)l * p.ics + row0 + r);
        __syncthreads();
        if (c0 < p.w) {
            // six row-slices per step: twelve independent global loads in flight per lane, a fold per three products
            for (uint32_t r = lane; r < rows; r += 192) {
                uint32_t e[6][BARY_COLS], rr[6];
#pragma unroll
                for (int u = 0; u < 6; u++) {
                    const uint32_t ru = r + 32 * u;
                    rr[u] = ru < rows ? ru : r;
#pragma unroll
                    for (int c = 0; c < BARY_COLS; c++) e[u][c] = ru < rows ? __ldg(col[c] + row0 + ru) : 0u;   // a zero term adds nothing
                }
#pragma unroll
                for (int g = 0; g < 6; g += 3) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        if ((uint32_t)q < p.npoints) {
#pragma unroll
                            for (int l = 0; l < 5; l++) {
                                const uint32_t d0 = dsm[q][l][rr[g]], d1 = dsm[q][l][rr[g + 1]], d2 = dsm[q][l][rr[g + 2]];
#pragma unroll
                                for (int c = 0; c < BARY_COLS; c++)
                                    acc[c][q][l] = lazy_fold(acc[c][q][l] + (uint64_t)e[g][c] * d0 + (uint64_t)e[g + 1][c] * d1 + (uint64_t)e[g + 2][c] * d2);
                            }
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
                    for (int u = 0; u < 6; u++) sum[c] += e[u][c];
            }
        }
    }
    if (c0 >= p.w) return;
    // reduce to field elements, then across the warp
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++) {
        uint32_t v[BARY_OUT];
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int l = 0; l < 5; l++) v[q * 5 + l] = bb::monty_reduce64(acc[c][q][l]);
        v[10] = (uint32_t)(sum[c] % bb::P);
#pragma unroll
        for (int k = 0; k < BARY_OUT; k++) {
            uint32_t x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = bb::add(x, __shfl_xor_sync(0xffffffffu, x, o));
            if (lane == 0 && c0 + c < p.w) p.partial[((uint64_t)blockIdx.x * p.w + c0 + c) * BARY_OUT + k] = x;
        }
    }
}

// K9
struct RoParams {
    const uint32_t* mat; uint64_t mcs; uint64_t H; uint32_t w;
    const E5* apow;                   // alpha^c, c < w
    const uint32_t* invden[2]; uint64_t ics; uint32_t npoints;
    E5 sum_y[2]; E5 alpha_off[2];     // per point: sum_c alpha^c y_c and alpha^offset
    uint32_t* ro; uint64_t rcs;       // accumulator, limb-major, height H
    uint64_t row_begin, row_end;      // rows swept by this launch
};
// one thread per LDE row; sum_c alpha^c p_c(x_i) accumulated lazily (see K8), three columns per fold
__global__ void __launch_bounds__(256) reduced_opening_kernel(RoParams p) {
    uint64_t i = p.row_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.row_end) return;
    uint64_t a[5] = {0, 0, 0, 0, 0};
    const uint32_t* m = p.mat + i;
    uint32_t c = 0;
    for (; c + 6 <= p.w; c += 6) {   // six loads in flight, two folds
        uint32_t e[6];
#pragma unroll
        for (int u = 0; u < 6; u++) e[u] = __ldg(m + (uint64_t)(c + u) * p.mcs);
#pragma unroll
        for (int g = 0; g < 6; g += 3) {
            const E5 a0 = p.apow[c + g], a1 = p.apow[c + g + 1], a2 = p.apow[c + g + 2];
#pragma unroll
            for (int l = 0; l < 5; l++) a[l] = lazy_fold(a[l] + (uint64_t)a0.c[l] * e[g] + (uint64_t)a1.c[l] * e[g + 1] + (uint64_t)a2.c[l] * e[g + 2]);
        }
    }
    for (; c + 3 <= p.w; c += 3) {
        const uint32_t e0 = __ldg(m + (uint64_t)c * p.mcs), e1 = __ldg(m + (uint64_t)(c + 1) * p.mcs), e2 = __ldg(m + (uint64_t)(c + 2) * p.mcs);
        const E5 a0 = p.apow[c], a1 = p.apow[c + 1], a2 = p.apow[c + 2];
#pragma unroll
        for (int l = 0; l < 5; l++) a[l] = lazy_fold(a[l] + (uint64_t)a0.c[l] * e0 + (uint64_t)a1.c[l] * e1 + (uint64_t)a2.c[l] * e2);
    }
    for (; c < p.w; c++) {
        const uint32_t e0 = __ldg(m + (uint64_t)c * p.mcs);
        const E5 a0 = p.apow[c];
#pragma unroll
        for (int l = 0; l < 5; l++) a[l] = lazy_fold(a[l] + (uint64_t)a0.c[l] * e0);
    }
    E5 red;
#pragma unroll
    for (int l = 0; l < 5; l++) red.c[l] = bb::monty_reduce64(a[l]);
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

static int32_t inverse_denominators_range(vgpu_ctx* ctx, uint32_t log_H, const E5& z, uint32_t* out, uint64_t begin, uint64_t count) {
    uint64_t H = 1ull << log_H;
    KScope ks(ctx, KC_INVDEN, 20.0 * (double)count);
    coset_minus_point_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(out, H, begin, count, log_H, bb::to_monty(bb::GEN_CANON), z, ctx->root_table.lo, ctx->root_table.hi);
    VG_LAUNCH_CHECK(ctx);
    return vg_ext_batch_inverse(ctx, out + begin, H, count, 1);
}
// When the row sweeps of this height are split across ranks, a rank only ever reads 1/(x - z) on its range of the
// reduced-opening sweep (all H rows) and on its range of the barycentric sweep (the first H/2 rows); rank 0's second
// range lies inside its first.
int32_t vg_inverse_denominators(vgpu_ctx* ctx, uint32_t log_H, const E5& z, uint32_t* out) {
    uint64_t H = 1ull << log_H;
    if (!vg_split_rows(ctx, H / 2)) return inverse_denominators_range(ctx, log_H, z, out, 0, H);
    const uint64_t G = (uint64_t)ctx->comm_size, r = (uint64_t)ctx->comm_rank;
    VG_TRY(inverse_denominators_range(ctx, log_H, z, out, r * (H / G), H / G));
    if (r > 0) VG_TRY(inverse_denominators_range(ctx, log_H, z, out, r * (H / 2 / G), H / 2 / G));
    return 0;
}

// p_c(z_q) for every column c and point q (q < npoints <= 2) of a committed LDE (height H = 2h).
int32_t vg_eval_columns(vgpu_ctx* ctx, const vgpu_dmat* lde, uint32_t npoints, const E5* z, const uint32_t* const* invden, std::vector<E5>* ys /* [q][c] */) {
    uint64_t H = lde->h, h = H / 2;
    uint32_t log_H = 0; while ((1ull << log_H) < H) log_H++;
    uint32_t w = (uint32_t)lde->w;
    // split across ranks: each rank sums its contiguous range of the h rows, the per-rank sums meet in one small all-gather
    const bool split = vg_split_rows(ctx, h);
    const uint64_t rows = split ? h / ctx->comm_size : h, row_begin = split ? rows * ctx->comm_rank : 0;
    const uint64_t ntiles = (rows + BARY_TILE - 1) / BARY_TILE;
    unsigned by = (w + BARY_GROUP - 1) / BARY_GROUP;
    unsigned bx = (unsigned)std::min<uint64_t>(ntiles, std::max<uint64_t>(1, 2 * (uint64_t)ctx->sm_count / by));
    uint32_t* partial = nullptr;
    size_t pn = (size_t)bx * w * BARY_OUT;
    VG_TRY(vg_alloc(ctx, (void**)&partial, pn * 4));
    BaryParams p{};
    p.mat = lde->d; p.mcs = lde->col_stride; p.h = rows; p.row_begin = row_begin; p.w = w;
    p.invden[0] = invden[0]; p.invden[1] = npoints > 1 ? invden[1] : invden[0]; p.ics = H; p.npoints = npoints;
    p.partial = partial;
    {
        KScope ks(ctx, KC_BARY, 4.0 * (double)rows * w);
        bary_kernel<<<dim3(bx, by), BARY_THREADS, 0, ctx->stream>>>(p);
    }
    VG_LAUNCH_CHECK(ctx);
    const uint32_t nout = w * BARY_OUT;
    uint32_t* reduced = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&reduced, (size_t)nout * 4 * (split ? ctx->comm_size + 1 : 1)));
    uint32_t* mine = split ? reduced + (size_t)nout * (1 + ctx->comm_rank) : reduced;     // [total | rank 0 | rank 1 | ...]
    bary_reduce_kernel<<<(nout + 255) / 256, 256, 0, ctx->stream>>>(partial, bx, nout, mine);
    VG_LAUNCH_CHECK(ctx);
    if (split) {
        VG_TRY(vg_comm_allgather_inplace(ctx, reduced + nout, nout));
        bary_reduce_kernel<<<(nout + 255) / 256, 256, 0, ctx->stream>>>(reduced + nout, (uint32_t)ctx->comm_size, nout, reduced);
        VG_LAUNCH_CHECK(ctx);
    }
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
            E5 D;   // sum_i p_c(x_i) / (x_i - z_q)
            for (int l = 0; l < 5; l++) D.c[l] = hp[(size_t)c * BARY_OUT + q * 5 + l];
            const E5 S = bb::e5_add_base(bb::e5_mul(z[q], D), hp[(size_t)c * BARY_OUT + 10]);
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
    // split across ranks: a rank accumulates only its range of rows; vg_reduced_openings_complete() joins the ranges
    const bool split = vg_split_rows(ctx, lde->h / 2);
    const uint64_t rows = split ? lde->h / ctx->comm_size : lde->h;
    p.row_begin = split ? rows * ctx->comm_rank : 0; p.row_end = p.row_begin + rows;
    {
        KScope ks(ctx, KC_REDUCED_OPENING, (double)rows * (4.0 * lde->w + 40.0));
        reduced_opening_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(p);
    }
    VG_LAUNCH_CHECK(ctx);
    return 0;
}

// after the last vg_reduced_opening_accumulate of a height: every rank gets the whole vector (5 limbs x H)
int32_t vg_reduced_openings_complete(vgpu_ctx* ctx, uint32_t* ro, uint64_t H) {
    if (!vg_split_rows(ctx, H / 2)) return 0;
    VG_TRY(vg_comm_group_begin(ctx));
    for (int l = 0; l < 5; l++) VG_TRY(vg_comm_allgather_inplace(ctx, ro + (uint64_t)l * H, H / ctx->comm_size));
    return vg_comm_group_end(ctx);
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
