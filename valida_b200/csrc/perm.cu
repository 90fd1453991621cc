// K5 — LogUp permutation trace on the device: generate_permutation_trace (machine/src/chip.rs:121-208)
// with generate_rlc_elements (291-331), reduce_row (335-352) and
// batch_multiplicative_inverse_allowing_zero (util/src/lib.rs:21-43; zeros stay zero).
// The reference walks rows serially with a Vec per row; here:
//   (a) one thread per row builds every interaction's denominator  alpha_bus + sum_j beta^j * field_j
//       from the column-major main/preprocessed traces (coalesced column reads);
//   (b) the ext5 inversion is batched per thread over 8 rows spaced one grid apart (Montgomery trick,
//       one Frobenius-norm inversion per 8 elements), so every access stays coalesced;
//   (c) the signed running sum phi is a 3-kernel block scan over the five limbs (addition mod p is
//       limb-wise in F_p[X]/(X^5-2)).
// Output layout = RowMajorMatrix<Challenge>::flatten_to_base, column-major on device: column 5m+l.
#include "ctx.h"
#include "devchip.h"
#include <cstdlib>
#include <memory>

namespace {

using bb::E5;

__device__ __forceinline__ uint32_t pair_col_eval(const DevPairCol& pc, const uint32_t* main, uint64_t mcs, const uint32_t* prep, uint64_t pcs, uint64_t row) {
    uint32_t v = pc.constant;
    for (uint32_t t = 0; t < pc.n_terms; t++) {
        uint32_t x = pc.is_prep[t] ? __ldg(prep + (uint64_t)pc.column[t] * pcs + row) : __ldg(main + (uint64_t)pc.column[t] * mcs + row);
        v = bb::add(v, bb::mul(x, pc.weight[t]));
    }
    return v;
}

// The chip descriptor travels in the kernel parameters (constant bank): descriptor reads are uniform constant loads.
__global__ void __launch_bounds__(256) perm_denominators_kernel(const __grid_constant__ DevChip chip_, const uint32_t* __restrict__ main, uint64_t mcs,
                                                               const uint32_t* __restrict__ prep, uint64_t pcs, uint64_t h, uint32_t* __restrict__ perm, uint64_t qcs) {
    const DevChip* chip = &chip_;
    uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= h) return;
    for (uint32_t m = 0; m < chip->n_interactions; m++) {
        const DevInteraction& it = chip->interactions[m];
        bb::Lazy5 ra; ra.init();        // sum_j beta^j * field_j as raw 64-bit products, reduced once
        for (uint32_t j = 0; j < it.n_fields; j++) ra.fma_base(chip->betas[j], pair_col_eval(it.fields[j], main, mcs, prep, pcs, n));
        const E5 rlc = bb::e5_add(it.alpha, ra.value());
#pragma unroll
        for (int l = 0; l < 5; l++) perm[(uint64_t)(5 * m + l) * qcs + n] = rlc.c[l];
    }
}

constexpr int INV_BATCH = 8;
// In-place inverse of `count` ext5 columns (column group g uses base columns 5g..5g+4); zero stays zero.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) ext_batch_inverse_kernel(uint32_t* __restrict__ data, uint64_t cs, uint64_t h, uint32_t groups) {
    uint64_t stride = (h + INV_BATCH - 1) / INV_BATCH;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= stride) return;
    for (uint32_t g = 0; g < groups; g++) {
        uint32_t* base = data + (uint64_t)(5 * g) * cs;
        E5 pref[INV_BATCH];
        E5 acc = bb::e5_one();
#pragma unroll
        for (int i = 0; i < INV_BATCH; i++) {
            uint64_t n = t + (uint64_t)i * stride;
            pref[i] = acc;
            if (n < h) {
                E5 d;
#pragma unroll
                for (int l = 0; l < 5; l++) d.c[l] = base[(uint64_t)l * cs + n];
                if (!bb::e5_is_zero(d)) acc = bb::e5_mul(acc, d);
            }
        }
        E5 inv = bb::e5_inv(acc);
#pragma unroll
        for (int i = INV_BATCH - 1; i >= 0; i--) {
            uint64_t n = t + (uint64_t)i * stride;
            if (n < h) {
                E5 d;
#pragma unroll
                for (int l = 0; l < 5; l++) d.c[l] = base[(uint64_t)l * cs + n];
                if (!bb::e5_is_zero(d)) {
                    E5 r = bb::e5_mul(inv, pref[i]);
                    inv = bb::e5_mul(inv, d);
#pragma unroll
                    for (int l = 0; l < 5; l++) base[(uint64_t)l * cs + n] = r.c[l];
                }
            }
        }
    }
}

// term[n] = sum_m (+-) q[n][m] * count_m(n), written into the phi columns (5k..5k+4)
__global__ void __launch_bounds__(256) perm_terms_kernel(const __grid_constant__ DevChip chip_, const uint32_t* __restrict__ main, uint64_t mcs,
                                                        const uint32_t* __restrict__ prep, uint64_t pcs, uint64_t h, uint32_t* __restrict__ perm, uint64_t qcs) {
    const DevChip* chip = &chip_;
    uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= h) return;
    uint32_t k = chip->n_interactions;
    E5 term = bb::e5_zero();
    for (uint32_t m = 0; m < k; m++) {
        const DevInteraction& it = chip->interactions[m];
        uint32_t mult = pair_col_eval(it.count, main, mcs, prep, pcs, n);
        E5 q;
#pragma unroll
        for (int l = 0; l < 5; l++) q.c[l] = perm[(uint64_t)(5 * m + l) * qcs + n];
        E5 t = bb::e5_mul_base(q, mult);
        term = it.is_send ? bb::e5_add(term, t) : bb::e5_sub(term, t);
    }
#pragma unroll
    for (int l = 0; l < 5; l++) perm[(uint64_t)(5 * k + l) * qcs + n] = term.c[l];
}

// Inclusive prefix sums mod p.  Each block scans SCAN_CHUNK consecutive elements of column blockIdx.y.
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 8, SCAN_CHUNK = SCAN_THREADS * SCAN_PER_THREAD;
__global__ void __launch_bounds__(SCAN_THREADS) scan_chunks_kernel(uint32_t* __restrict__ data, uint64_t cs, uint64_t n, uint32_t* __restrict__ chunk_sums, uint64_t sums_cs) {
    __shared__ uint32_t buf[SCAN_CHUNK];
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    uint32_t* col = data + (uint64_t)blockIdx.y * cs;
    uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK;
    for (int i = threadIdx.x; i < SCAN_CHUNK; i += SCAN_THREADS) buf[i] = base + i < n ? col[base + i] : 0;
    __syncthreads();
    uint32_t loc[SCAN_PER_THREAD];
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) { run = bb::add(run, buf[threadIdx.x * SCAN_PER_THREAD + i]); loc[i] = run; }
    // exclusive scan of per-thread totals across the block
    uint32_t v = run;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = bb::add(v, u); }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w = bb::add(w, u); }
        if (lane < SCAN_THREADS / 32) wsum[lane] = w;
    }
    __syncthreads();
    uint32_t offset = bb::sub(v, run);                    // exclusive within warp
    if (wid > 0) offset = bb::add(offset, wsum[wid - 1]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; i++) buf[threadIdx.x * SCAN_PER_THREAD + i] = bb::add(loc[i], offset);
    __syncthreads();
    for (int i = threadIdx.x; i < SCAN_CHUNK; i += SCAN_THREADS) if (base + i < n) col[base + i] = buf[i];
    if (threadIdx.x == SCAN_THREADS - 1 && chunk_sums) chunk_sums[(uint64_t)blockIdx.y * sums_cs + blockIdx.x] = bb::add(loc[SCAN_PER_THREAD - 1], offset);
}
// Serial-over-chunks inclusive scan of a short array by ONE block per column (used on the chunk sums).
__global__ void __launch_bounds__(1024) scan_small_kernel(uint32_t* __restrict__ data, uint64_t cs, uint64_t n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    uint32_t* col = data + (uint64_t)blockIdx.y * cs;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += 1024) {
        uint64_t i = base + threadIdx.x;
        uint32_t x = i < n ? col[i] : 0, v = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = bb::add(v, u); }
        if (lane == 31) wsum[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w = bb::add(w, u); }
            wsum[lane] = w;
        }
        __syncthreads();
        uint32_t r = bb::add(v, carry_s);
        if (wid > 0) r = bb::add(r, wsum[wid - 1]);
        if (i < n) col[i] = r;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = r;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) scan_add_offsets_kernel(uint32_t* __restrict__ data, uint64_t cs, uint64_t n, const uint32_t* __restrict__ chunk_sums, uint64_t sums_cs) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t chunk = i / SCAN_CHUNK;
    if (chunk == 0) return;
    uint32_t* col = data + (uint64_t)blockIdx.y * cs;
    col[i] = bb::add(col[i], chunk_sums[(uint64_t)blockIdx.y * sums_cs + chunk - 1]);
}

}  // namespace

// Inclusive prefix sum (mod p) of `ncols` columns of length n, in place.
int32_t vg_prefix_sum_columns(vgpu_ctx* ctx, uint32_t* data, uint64_t cs, uint64_t n, uint32_t ncols) {
    uint64_t chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    uint32_t* sums = nullptr;
    if (chunks > 1) VG_TRY(vg_alloc(ctx, (void**)&sums, chunks * ncols * 4));
    scan_chunks_kernel<<<dim3((unsigned)chunks, ncols), SCAN_THREADS, 0, ctx->stream>>>(data, cs, n, sums, chunks);
    VG_LAUNCH_CHECK(ctx);
    if (chunks > 1) {
        scan_small_kernel<<<dim3(1, ncols), 1024, 0, ctx->stream>>>(sums, chunks, chunks);
        VG_LAUNCH_CHECK(ctx);
        scan_add_offsets_kernel<<<dim3((unsigned)((n + 255) / 256), ncols), 256, 0, ctx->stream>>>(data, cs, n, sums, chunks);
        VG_LAUNCH_CHECK(ctx);
        vg_free(ctx, sums);
    }
    return 0;
}

int32_t vg_ext_batch_inverse(vgpu_ctx* ctx, uint32_t* data, uint64_t cs, uint64_t h, uint32_t groups) {
    if (!groups || !h) return 0;
    uint64_t stride = (h + INV_BATCH - 1) / INV_BATCH;
    static const int minb = [] { const char* e = getenv("VGPU_EXTINV_MINB"); return e ? atoi(e) : 5; }();   // tuning knob (profiles/)
    if (minb == 8) ext_batch_inverse_kernel<8><<<(unsigned)((stride + 127) / 128), 128, 0, ctx->stream>>>(data, cs, h, groups);
    else if (minb == 6) ext_batch_inverse_kernel<6><<<(unsigned)((stride + 127) / 128), 128, 0, ctx->stream>>>(data, cs, h, groups);
    else ext_batch_inverse_kernel<5><<<(unsigned)((stride + 127) / 128), 128, 0, ctx->stream>>>(data, cs, h, groups);
    VG_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int32_t vgpu_perm_trace(vgpu_ctx* ctx, const vgpu_chip_desc* chip, const vgpu_dmat* main, const vgpu_dmat* prep_or_null,
                                   const uint32_t challenges[15], vgpu_dmat** out_perm, uint32_t cumulative_sum_out[5]) {
    if (!chip || !main || !out_perm) VG_FAIL(ctx, "perm_trace: null argument");
    if (main->bitrev_rows) VG_FAIL(ctx, "perm_trace: main trace rows are stored bit-reversed");
    if (main->w != chip->width) VG_FAIL(ctx, "perm_trace: main width %llu != chip width %u", (unsigned long long)main->w, chip->width);
    if (chip->preprocessed_width && (!prep_or_null || prep_or_null->w != chip->preprocessed_width || prep_or_null->h != main->h)) {
        // interactions of BasicMachine never read preprocessed columns, but the shape must still be coherent when given
        if (prep_or_null) VG_FAIL(ctx, "perm_trace: preprocessed trace shape mismatch");
    }
    VG_TRY(vg_dmat_materialize(ctx, main));
    VG_TRY(vg_dmat_materialize(ctx, prep_or_null));
    auto dchip_h = std::make_unique<DevChip>();
    VG_TRY(vg_build_devchip(ctx, chip, challenges, dchip_h.get()));
    const DevChip& dchip = *dchip_h;
    uint64_t h = main->h;
    uint32_t k = chip->n_interactions;
    vgpu_dmat* perm = nullptr;
    int32_t rc = vg_dmat_alloc(ctx, h, 5 * (k + 1), &perm);
    if (rc) return rc;
    const uint32_t* pd = prep_or_null ? prep_or_null->d : nullptr;
    uint64_t pcs = prep_or_null ? prep_or_null->col_stride : 0;
    unsigned blocks = (unsigned)((h + 255) / 256);
    KScope ks(ctx, KC_PERM, 4.0 * (double)h * (chip->width + 5.0 * (k + 1)));
    if (k) {
        perm_denominators_kernel<<<blocks, 256, 0, ctx->stream>>>(dchip, main->d, main->col_stride, pd, pcs, h, perm->d, perm->col_stride);
        VG_LAUNCH_CHECK(ctx);
        VG_TRY(vg_ext_batch_inverse(ctx, perm->d, perm->col_stride, h, k));
    }
    perm_terms_kernel<<<blocks, 256, 0, ctx->stream>>>(dchip, main->d, main->col_stride, pd, pcs, h, perm->d, perm->col_stride);
    VG_LAUNCH_CHECK(ctx);
    VG_TRY(vg_prefix_sum_columns(ctx, perm->d + (uint64_t)(5 * k) * perm->col_stride, perm->col_stride, h, 5));
    if (cumulative_sum_out) {
        uint32_t cs[5];
        for (int l = 0; l < 5; l++)
            VG_CUDA(ctx, cudaMemcpyAsync(&cs[l], perm->d + (uint64_t)(5 * k + l) * perm->col_stride + (h - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
        VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int l = 0; l < 5; l++) cumulative_sum_out[l] = bb::from_monty(cs[l]);
    }
    *out_perm = perm;
    return 0;
}
