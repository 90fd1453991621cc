// K12 — host <-> device staging: RowMajorMatrix<Val> (machine/src/config.rs:17-22 boundary type)
// to/from the device's column-major Montgomery store.  One H2D/D2H copy of the row-major image plus
// a shared-memory tile transpose: 64 rows x w words are read as one contiguous run (coalesced) and
// written as w runs of 64 consecutive rows (256 B segments).
#include "ctx.h"

namespace {
constexpr int TR = 64;      // rows per tile
constexpr int TW = 128;     // max columns per tile

__global__ void __launch_bounds__(256) rm_to_cm_kernel(const uint32_t* __restrict__ rm, uint64_t h, uint64_t w, uint32_t* __restrict__ cm, uint64_t cs, int to_monty, uint64_t c0, uint32_t wc) {
    __shared__ uint32_t tile[TR][TW + 1];
    uint64_t r0 = (uint64_t)blockIdx.x * TR;
    uint32_t rows = (uint32_t)(h - r0 < TR ? h - r0 : TR);
    if (wc == w) {   // whole rows are contiguous: one run of rows*w words
        uint32_t tot = rows * wc;
        const uint32_t* src = rm + r0 * w;
        for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) tile[i / wc][i % wc] = src[i];
    } else {
        for (uint32_t i = threadIdx.x; i < rows * wc; i += blockDim.x) { uint32_t r = i / wc, c = i % wc; tile[r][c] = rm[(r0 + r) * w + c0 + c]; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < rows * wc; i += blockDim.x) {
        uint32_t c = i / rows, r = i % rows;
        uint32_t v = tile[r][c];
        cm[(c0 + c) * cs + r0 + r] = to_monty ? bb::to_monty(v) : v;
    }
}
__global__ void __launch_bounds__(256) cm_to_rm_kernel(const uint32_t* __restrict__ cm, uint64_t cs, uint64_t h, uint64_t w, uint32_t* __restrict__ rm, int from_monty, uint64_t c0, uint32_t wc) {
    __shared__ uint32_t tile[TR][TW + 1];
    uint64_t r0 = (uint64_t)blockIdx.x * TR;
    uint32_t rows = (uint32_t)(h - r0 < TR ? h - r0 : TR);
    for (uint32_t i = threadIdx.x; i < rows * wc; i += blockDim.x) {
        uint32_t c = i / rows, r = i % rows;
        uint32_t v = cm[(c0 + c) * cs + r0 + r];
        tile[r][c] = from_monty ? bb::from_monty(v) : v;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < rows * wc; i += blockDim.x) { uint32_t r = i / wc, c = i % wc; rm[(r0 + r) * w + c0 + c] = tile[r][c]; }
}
}  // namespace

static int32_t transpose_in(vgpu_ctx* ctx, const uint32_t* stage, uint64_t h, uint64_t w, int32_t repr, vgpu_dmat* dst) {
    for (uint64_t c0 = 0; c0 < w; c0 += TW) {
        uint32_t wc = (uint32_t)(w - c0 < TW ? w - c0 : TW);
        KScope ks(ctx, KC_TRANSPOSE, 8.0 * (double)h * wc);
        rm_to_cm_kernel<<<(unsigned)((h + TR - 1) / TR), 256, 0, ctx->stream>>>(stage, h, w, dst->d, dst->col_stride, repr == VGPU_REPR_CANONICAL, c0, wc);
        VG_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int32_t vg_upload_rowmajor(vgpu_ctx* ctx, const uint32_t* host, uint64_t h, uint64_t w, int32_t repr, vgpu_dmat* dst) {
    if (h == 0 || w == 0) return 0;
    uint32_t* stage = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&stage, h * w * 4));
    VG_CUDA(ctx, cudaMemcpyAsync(stage, host, h * w * 4, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = transpose_in(ctx, stage, h, w, repr, dst);
    vg_free(ctx, stage);
    return rc;
}

// Pipelined form used by vgpu_prove: the copy is queued on the context's copy stream (so it overlaps the kernels of
// matrices that arrived earlier) and the transpose is deferred to the matrix's first use on the compute stream.
int32_t vg_upload_begin(vgpu_ctx* ctx, const uint32_t* host, uint64_t h, uint64_t w, int32_t repr, vgpu_dmat* dst) {
    if (h == 0 || w == 0) return 0;
    if (!ctx->copy_stream) VG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    VG_TRY(vg_alloc(ctx, (void**)&dst->pend_stage, h * w * 4));
    {   // the staging block may be a cached one whose last user is still queued on the compute stream
        cudaEvent_t fence;
        if (!ctx->event_pool.empty()) { fence = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else VG_CUDA(ctx, cudaEventCreate(&fence));
        VG_CUDA(ctx, cudaEventRecord(fence, ctx->stream));
        VG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, fence, 0));
        ctx->event_pool.push_back(fence);
    }
    VG_CUDA(ctx, cudaMemcpyAsync(dst->pend_stage, host, h * w * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (!ctx->event_pool.empty()) { dst->pend_ev = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
    else VG_CUDA(ctx, cudaEventCreate(&dst->pend_ev));
    VG_CUDA(ctx, cudaEventRecord(dst->pend_ev, ctx->copy_stream));
    dst->pend_repr = repr;
    return 0;
}

int32_t vg_dmat_materialize(vgpu_ctx* ctx, const vgpu_dmat* cm) {
    vgpu_dmat* m = const_cast<vgpu_dmat*>(cm);   // completing a pending upload does not change the matrix's value
    if (!m || !m->pend_stage) return 0;
    VG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->pend_ev, 0));
    int32_t rc = transpose_in(ctx, m->pend_stage, m->h, m->w, m->pend_repr, m);
    vg_free(ctx, m->pend_stage);                 // reused only by later work on the compute stream, i.e. after the transpose
    m->pend_stage = nullptr;
    ctx->event_pool.push_back(m->pend_ev);
    m->pend_ev = nullptr;
    return rc;
}

int32_t vg_download_rowmajor(vgpu_ctx* ctx, const vgpu_dmat* src, int32_t repr, uint32_t* host) {
    uint64_t h = src->h, w = src->w;
    if (h == 0 || w == 0) return 0;
    uint32_t* stage = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&stage, h * w * 4));
    for (uint64_t c0 = 0; c0 < w; c0 += TW) {
        uint32_t wc = (uint32_t)(w - c0 < TW ? w - c0 : TW);
        KScope ks(ctx, KC_TRANSPOSE, 8.0 * (double)h * wc);
        cm_to_rm_kernel<<<(unsigned)((h + TR - 1) / TR), 256, 0, ctx->stream>>>(src->d, src->col_stride, h, w, stage, repr == VGPU_REPR_CANONICAL, c0, wc);
        VG_LAUNCH_CHECK(ctx);
    }
    VG_CUDA(ctx, cudaMemcpyAsync(host, stage, h * w * 4, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    vg_free(ctx, stage);
    return 0;
}
