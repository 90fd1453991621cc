// Internal context / device-matrix definitions shared by the kernels' host wrappers.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include <cuda_runtime.h>
#include "../../include/valida_b200.h"
#include "bb.cuh"

constexpr int VG_LOG_NMAX = 27;                    // BabyBear two-adicity: largest transform/LDE size
constexpr int VG_POW_LO_BITS = 12;                 // two-level power tables: base^e = lo[e & 4095] * hi[e >> 12]
constexpr uint32_t VG_POW_LO = 1u << VG_POW_LO_BITS;

struct PowTable {            // device tables of Montgomery words
    uint32_t* lo = nullptr;  // base^j, j < 4096
    uint32_t* hi = nullptr;  // scale * base^(4096 j), j < hi_len
    uint32_t hi_len = 0;
    uint32_t base = 0;       // the base itself (Montgomery), without the scale
};

// kernel classes for the optional per-launch CUDA-event timing (bench.py's roofline line)
enum KClass { KC_NTT = 0, KC_LEAF_HASH, KC_COMPRESS, KC_FRI_LEAF, KC_TRANSPOSE, KC_PERM, KC_QUOTIENT, KC_INVDEN, KC_BARY, KC_REDUCED_OPENING, KC_FRI_FOLD, KC_OTHER, KC_COUNT };
struct KTimer { cudaEvent_t a, b; int cls; double bytes; };

struct vgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    PowTable root_table;                                        // base = two_adic_generator(27)
    uint32_t* root3 = nullptr;                                   // 3 x 512 words: w^(i), w^(512 i), w^(2^18 i) — a 6 KB, L1-resident form of the same table
    std::map<std::pair<uint32_t, uint32_t>, PowTable> shift_tables;  // (shift, scale) canonical -> table
    std::vector<void*> owned;                                   // freed at destroy
    // Poseidon challenger instance (host side; the transcript is sequential and tiny)
    uint32_t poseidon_rc[480];
    uint32_t poseidon_mds[256];
    bool challenger_set = false;
    bool poseidon_has_mds = false;
    void* challenger = nullptr;                                  // vgh::Challenger* (host/challenger.h)
    void* poseidon = nullptr;                                    // vgh::Poseidon16*
    std::vector<std::pair<const char*, float>> phases;          // last prove: per-phase milliseconds
    // size-keyed cache of device buffers: a proof repeats the same allocation sizes every step, so after the
    // first step no driver allocator call is made (single stream => reuse in enqueue order is safe)
    std::multimap<size_t, void*> free_bufs;
    std::map<void*, size_t> live_bufs;
    size_t cached_bytes = 0, live_bytes = 0, peak_bytes = 0;
    // multi-GPU (host/comm.cc): one rank per GPU, NCCL communicator bound to ctx->stream
    void* nccl = nullptr;                                        // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    bool sharding = false;                                       // commit / FRI-commit work split across ranks
    cudaStream_t copy_stream = nullptr;                         // H2D copies of a pipelined vgpu_prove (staging.cu)
    bool ntt_attrs_set = false, bary_attrs_set = false;          // cudaFuncSetAttribute is per device: tracked per context, not per process
    bool ktiming = false;
    std::vector<KTimer> ktimers;
    std::vector<cudaEvent_t> event_pool;
};

struct vgpu_dmat {
    vgpu_ctx* ctx = nullptr;
    uint32_t* d = nullptr;       // column-major: element (r, c) at d[c * col_stride + r], Montgomery form
    uint64_t h = 0, w = 0, col_stride = 0;
    bool owns = true;
    bool bitrev_rows = false;    // row r of the logical matrix is stored at reverse_bits(r) (quotient-chunk output order)
    // pipelined upload: the row-major image is (being) copied into pend_stage on the copy stream; the transpose into
    // `d` runs on the context's stream at first use (vg_dmat_materialize)
    uint32_t* pend_stage = nullptr;
    cudaEvent_t pend_ev = nullptr;
    int32_t pend_repr = 0;
};

#define VG_FAIL(ctx, ...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); (ctx)->err = _b; return -1; } while (0)
#define VG_CUDA(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { VG_FAIL(ctx, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); } } while (0)
#define VG_TRY(expr) do { int32_t _r = (expr); if (_r != 0) return _r; } while (0)
#define VG_LAUNCH_CHECK(ctx) do { (ctx)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { VG_FAIL(ctx, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); } } while (0)

// RAII scope: when ctx->ktiming is on, brackets the launches inside it with a CUDA event pair on ctx->stream.
struct KScope {
    vgpu_ctx* ctx; bool on;
    KScope(vgpu_ctx* c, int cls, double bytes) : ctx(c), on(c->ktiming) {
        if (!on) return;
        KTimer t; t.cls = cls; t.bytes = bytes;
        for (cudaEvent_t* e : {&t.a, &t.b}) {
            if (!c->event_pool.empty()) { *e = c->event_pool.back(); c->event_pool.pop_back(); }
            else cudaEventCreate(e);
        }
        cudaEventRecord(t.a, c->stream);
        c->ktimers.push_back(t);
    }
    ~KScope() { if (on) cudaEventRecord(ctx->ktimers.back().b, ctx->stream); }
};

int32_t vg_alloc(vgpu_ctx* ctx, void** p, size_t bytes);
void vg_free(vgpu_ctx* ctx, void* p);
int32_t vg_dmat_alloc(vgpu_ctx* ctx, uint64_t h, uint64_t w, vgpu_dmat** out);
int32_t vg_get_shift_table(vgpu_ctx* ctx, uint32_t shift_canonical, uint32_t scale_canonical, uint64_t max_exp, const PowTable** out);

// host/comm.cc — every rank calls these in the same order with the same sizes
inline bool vg_sharded(const vgpu_ctx* ctx) { return ctx->sharding && ctx->comm_size > 1; }
// row-wise sweeps over n rows are cut into comm_size contiguous ranges when every range keeps >= 4096 rows
inline bool vg_split_rows(const vgpu_ctx* ctx, uint64_t n) { return vg_sharded(ctx) && n >= (uint64_t)ctx->comm_size * 4096; }
void vg_shard_range(uint64_t total, int nranks, int rank, uint64_t* begin, uint64_t* end);   // contiguous, balanced
int32_t vg_comm_group_begin(vgpu_ctx* ctx);
int32_t vg_comm_group_end(vgpu_ctx* ctx);
// buf holds comm_size consecutive blocks of `words_per_rank` u32; this rank's block is already filled
int32_t vg_comm_allgather_inplace(vgpu_ctx* ctx, uint32_t* buf, uint64_t words_per_rank);
int32_t vg_comm_bcast(vgpu_ctx* ctx, uint32_t* buf, uint64_t words, int root);
void vg_comm_free(vgpu_ctx* ctx);

// ntt.cu
int32_t vg_ntt_nat2nat(vgpu_ctx* ctx, const uint32_t* src, uint64_t src_cs, uint32_t* dst, uint64_t dst_cs, int log_n, uint64_t w,
                       bool inverse, const PowTable* coset_or_null, uint32_t* tmp, uint64_t tmp_cs);
int32_t vg_coset_lde(vgpu_ctx* ctx, const uint32_t* src, uint64_t src_cs, uint64_t h, uint64_t w, uint32_t shift_canonical,
                     uint32_t* dst, uint64_t dst_cs, bool bit_reversed, bool src_bitrev = false);
// staging.cu
int32_t vg_upload_begin(vgpu_ctx* ctx, const uint32_t* host, uint64_t h, uint64_t w, int32_t repr, vgpu_dmat* dst);   // async copy only
int32_t vg_dmat_materialize(vgpu_ctx* ctx, const vgpu_dmat* m);                                                       // no-op unless an upload is pending
int32_t vg_upload_rowmajor(vgpu_ctx* ctx, const uint32_t* host, uint64_t h, uint64_t w, int32_t repr, vgpu_dmat* dst);
int32_t vg_download_rowmajor(vgpu_ctx* ctx, const vgpu_dmat* src, int32_t repr, uint32_t* host);
