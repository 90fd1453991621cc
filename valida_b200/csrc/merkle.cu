// K3/K4 — Keccak-256 Merkle commitment over mixed-height column-major LDE matrices (sm_100a).
// Replaces FieldMerkleTreeMmcs<BabyBear, SerializingHasher32<Keccak256Hash>,
// CompressionFunctionFromHasher<_,_,2,8>, 8>::commit (basic/src/bin/valida.rs:367-374) as reached
// from TwoAdicFriPcs::commit_shifted_batches (derive/src/lib.rs:309,330,355,372).
//  * leaf kernel: one thread per LDE row; the row of every matrix of that height is streamed from
//    the column-major store (coalesced across the warp), converted Montgomery -> canonical, absorbed
//    little-endian into a register-resident 1600-bit state; digest words are reduced mod p;
//  * node kernel: one thread per parent, 64-byte compression (one permutation), with the
//    "inject shorter matrices" rule: node = compress(compress(l, r), hash(rows at that height)).
// Digests are stored canonical, 8 words (32 B) per node, all layers kept for the opening phase.
#include "ctx.h"
#include "keccak.cuh"
#include "merkle.h"
#include <algorithm>
#include <numeric>

namespace {

constexpr int RATE_WORDS = 34;   // 136-byte rate

// Sponge over `nwords` canonical words fetched by `fetch(i)`; Keccak pad 0x01 .. 0x80.
template <class Fetch>
__device__ __forceinline__ void keccak256_words(uint32_t nwords, Fetch fetch, uint32_t out[8]) {
    uint2 A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = make_uint2(0, 0);
    uint32_t nblocks = nwords / RATE_WORDS + 1;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t base = b * RATE_WORDS;
#pragma unroll
        for (int i = 0; i < RATE_WORDS / 2; i++) {
            uint32_t g0 = base + 2 * i, g1 = g0 + 1;
            uint32_t w0 = g0 < nwords ? fetch(g0) : (g0 == nwords ? 1u : 0u);
            uint32_t w1 = g1 < nwords ? fetch(g1) : (g1 == nwords ? 1u : 0u);
            if (i == RATE_WORDS / 2 - 1 && b == nblocks - 1) w1 ^= 0x80000000u;
            A[i].x ^= w0; A[i].y ^= w1;
        }
        kk::keccak_f(A);
    }
    out[0] = A[0].x; out[1] = A[0].y; out[2] = A[1].x; out[3] = A[1].y;
    out[4] = A[2].x; out[5] = A[2].y; out[6] = A[3].x; out[7] = A[3].y;
}

__device__ __forceinline__ uint32_t wrap_mod_p(uint32_t w) {   // F::from_wrapped_u32
    w = bb::umin32(w, w - bb::P);
    return bb::umin32(w, w - bb::P);
}

// colptr[i] = device pointer to column i of the concatenated row (all matrices of this height).
// rows [row0, row0 + nrows) of the layer (a rank's share when the tree is split across GPUs)
__global__ void __launch_bounds__(128) leaf_hash_kernel(const uint32_t* const* __restrict__ colptr, uint32_t nwords, uint64_t row0, uint64_t nrows, uint32_t* __restrict__ digests) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    r += row0;
    uint32_t d[8];
    keccak256_words(nwords, [&](uint32_t i) { return bb::from_monty(__ldg(colptr[i] + r)); }, d);
    uint4* o = reinterpret_cast<uint4*>(digests + r * 8);
    o[0] = make_uint4(wrap_mod_p(d[0]), wrap_mod_p(d[1]), wrap_mod_p(d[2]), wrap_mod_p(d[3]));
    o[1] = make_uint4(wrap_mod_p(d[4]), wrap_mod_p(d[5]), wrap_mod_p(d[6]), wrap_mod_p(d[7]));
}

__device__ __forceinline__ void compress_pair(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint32_t d[8];
    keccak256_words(16, [&](uint32_t i) { return i < 8 ? l[i] : r[i - 8]; }, d);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = wrap_mod_p(d[i]);
}

// next[i] = compress(prev[2i], prev[2i+1]) ; if inject != null: next[i] = compress(next[i], inject[i])
// parents [i0, i0 + n_next) of the layer
__global__ void __launch_bounds__(128) compress_layer_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ inject, uint64_t i0, uint64_t n_next, uint32_t* __restrict__ next) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_next) return;
    i += i0;
    uint32_t l[8], r[8], o[8];
    const uint4* p = reinterpret_cast<const uint4*>(prev + i * 16);
    uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = b.x; l[5] = b.y; l[6] = b.z; l[7] = b.w;
    r[0] = c.x; r[1] = c.y; r[2] = c.z; r[3] = c.w; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = d.w;
    compress_pair(l, r, o);
    if (inject) {
        const uint4* q = reinterpret_cast<const uint4*>(inject + i * 8);
        uint4 e = __ldg(q), f = __ldg(q + 1);
        uint32_t t[8] = {e.x, e.y, e.z, e.w, f.x, f.y, f.z, f.w};
        uint32_t o2[8];
        compress_pair(o, t, o2);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = o2[k];
    }
    uint4* w = reinterpret_cast<uint4*>(next + i * 8);
    w[0] = make_uint4(o[0], o[1], o[2], o[3]);
    w[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

int32_t hash_rows(vgpu_ctx* ctx, const std::vector<const vgpu_dmat*>& mats, uint64_t row0, uint64_t nrows, uint32_t* digests) {
    std::vector<const uint32_t*> cols;
    for (auto* m : mats) for (uint64_t c = 0; c < m->w; c++) cols.push_back(m->d + c * m->col_stride);
    const uint32_t** dcols = nullptr;
    VG_TRY(vg_alloc(ctx, (void**)&dcols, cols.size() * sizeof(void*)));
    VG_CUDA(ctx, cudaMemcpyAsync(dcols, cols.data(), cols.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    // the host vector must outlive the async copy from pageable memory: cudaMemcpyAsync from pageable memory
    // stages synchronously, so it is safe to let `cols` go once the call returns.
    {
        KScope ks(ctx, KC_LEAF_HASH, (double)nrows * (4.0 * cols.size() + 32.0));
        leaf_hash_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, ctx->stream>>>(dcols, (uint32_t)cols.size(), row0, nrows, digests);
    }
    VG_LAUNCH_CHECK(ctx);
    vg_free(ctx, dcols);
    return 0;
}

// FRI commit-phase leaf: the pair (v[2i], v[2i+1]) of ext5 values flattened to 10 base words
// (ExtensionMmcs over a width-2 matrix); v is limb-major: limb l of element e at v[l * cs + e].
__global__ void __launch_bounds__(128) fri_leaf_hash_kernel(const uint32_t* __restrict__ v, uint64_t cs, uint64_t i0, uint64_t npairs, uint32_t* __restrict__ digests) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    i += i0;
    uint32_t w[10];
#pragma unroll
    for (int l = 0; l < 5; l++) {
        uint2 t = __ldg(reinterpret_cast<const uint2*>(v + (uint64_t)l * cs + 2 * i));
        w[l] = bb::from_monty(t.x); w[5 + l] = bb::from_monty(t.y);
    }
    uint32_t d[8];
    keccak256_words(10, [&](uint32_t k) { return w[k]; }, d);
    uint4* o = reinterpret_cast<uint4*>(digests + i * 8);
    o[0] = make_uint4(wrap_mod_p(d[0]), wrap_mod_p(d[1]), wrap_mod_p(d[2]), wrap_mod_p(d[3]));
    o[1] = make_uint4(wrap_mod_p(d[4]), wrap_mod_p(d[5]), wrap_mod_p(d[6]), wrap_mod_p(d[7]));
}

}  // namespace

// Tree layers split across the ranks of the communicator: a layer of `len` nodes with len >= comm_size is cut into
// comm_size contiguous shares; a rank derives its share of every such layer from its own share of the layer below
// (parents [k*len/G, (k+1)*len/G) need exactly children [k*2len/G, (k+1)*2len/G)), so no exchange is needed on the
// way up.  One grouped all-gather then completes every layer on every rank, and the layers shorter than comm_size
// are computed by all ranks.  Unsplit (single GPU / sharding off): share = the whole layer.
struct Share { uint64_t begin, count; bool split; };
static Share share_of(const vgpu_ctx* ctx, uint64_t len) {
    Share s; int32_t split = 0;
    vgpu_tree_share(len, vg_sharded(ctx) ? ctx->comm_size : 1, ctx->comm_rank, &s.begin, &s.count, &split);
    s.split = split != 0;
    return s;
}
// complete the split layers [first, last) of a tree whose layer i starts at layer_ptr[i] and has layer_len[i] nodes
static int32_t gather_split_layers(vgpu_ctx* ctx, const std::vector<uint32_t*>& layer_ptr, const std::vector<uint64_t>& layer_len, size_t first, size_t last) {
    if (first >= last) return 0;
    VG_TRY(vg_comm_group_begin(ctx));
    for (size_t i = first; i < last; i++) VG_TRY(vg_comm_allgather_inplace(ctx, layer_ptr[i], layer_len[i] / (uint64_t)ctx->comm_size * 8));
    return vg_comm_group_end(ctx);
}

// Single-matrix tree over ext5 pairs (p3-fri commit phase): digests = [leaf layer | ... | root].
int32_t vg_fri_layer_commit(vgpu_ctx* ctx, const uint32_t* v, uint64_t cs, uint64_t npairs, uint32_t* digests,
                            std::vector<uint32_t*>* layer_ptr, std::vector<uint64_t>* layer_len, uint32_t root_out[8]) {
    Share sh = share_of(ctx, npairs);
    {
        KScope ks(ctx, KC_FRI_LEAF, (double)sh.count * 72.0);
        fri_leaf_hash_kernel<<<(unsigned)((sh.count + 127) / 128), 128, 0, ctx->stream>>>(v, cs, sh.begin, sh.count, digests);
    }
    VG_LAUNCH_CHECK(ctx);
    uint32_t* layer = digests;
    uint64_t len = npairs;
    layer_ptr->clear(); layer_len->clear();
    layer_ptr->push_back(layer); layer_len->push_back(len);
    size_t n_split = sh.split ? 1 : 0;
    bool gathered = !sh.split;
    while (len > 1) {
        uint64_t next_len = len / 2;
        uint32_t* next = layer + len * 8;
        sh = share_of(ctx, next_len);
        if (!sh.split && !gathered) { VG_TRY(gather_split_layers(ctx, *layer_ptr, *layer_len, 0, n_split)); gathered = true; }
        {
            KScope ks(ctx, KC_COMPRESS, (double)sh.count * 96.0);
            compress_layer_kernel<<<(unsigned)((sh.count + 127) / 128), 128, 0, ctx->stream>>>(layer, nullptr, sh.begin, sh.count, next);
        }
        VG_LAUNCH_CHECK(ctx);
        layer_ptr->push_back(next); layer_len->push_back(next_len);
        if (sh.split) n_split++;
        layer = next; len = next_len;
    }
    if (!gathered) VG_TRY(gather_split_layers(ctx, *layer_ptr, *layer_len, 0, n_split));
    VG_CUDA(ctx, cudaMemcpyAsync(root_out, layer, 32, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Build the mixed-height tree over the bit-reversed LDE matrices of `pd` (see merkle.h for `need`).
int32_t vg_merkle_build(vgpu_ctx* ctx, vgpu_prover_data* pd, const std::vector<uint64_t>& heights,
                        const std::function<int32_t(const std::vector<size_t>&)>& need) {
    size_t n = heights.size();
    if (!n) VG_FAIL(ctx, "commit: no matrices");
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return heights[a] > heights[b]; });
    uint64_t max_h = heights[order[0]];
    if (max_h & (max_h - 1)) VG_FAIL(ctx, "commit: heights must be powers of two");
    pd->max_height = max_h;
    // all layers in one allocation: max_h + max_h/2 + ... + 1 = 2*max_h - 1 digests
    VG_TRY(vg_alloc(ctx, (void**)&pd->digests, (2 * max_h) * 32));
    size_t pos = 0;
    std::vector<size_t> idx;
    std::vector<const vgpu_dmat*> group;
    auto take = [&](uint64_t height) -> int32_t {   // the matrices of that height, extended on demand
        idx.clear(); group.clear();
        while (pos < n && heights[order[pos]] == height) idx.push_back(order[pos++]);
        if (idx.empty()) return 0;
        if (need) VG_TRY(need(idx));
        for (size_t i : idx) {
            if (!pd->ldes[i] || pd->ldes[i]->h != height) VG_FAIL(ctx, "commit: matrix %zu was not extended to height %llu", i, (unsigned long long)height);
            group.push_back(pd->ldes[i]);
        }
        return 0;
    };
    VG_TRY(take(max_h));
    uint32_t* layer = pd->digests;
    pd->layer_ptr.clear(); pd->layer_len.clear();
    pd->layer_ptr.push_back(layer); pd->layer_len.push_back(max_h);
    Share sh = share_of(ctx, max_h);
    VG_TRY(hash_rows(ctx, group, sh.begin, sh.count, layer));
    size_t n_split = sh.split ? 1 : 0;
    bool gathered = !sh.split;
    uint32_t* inject_buf = nullptr;
    uint64_t len = max_h;
    while (len > 1) {
        uint64_t next_len = len / 2;
        sh = share_of(ctx, next_len);
        if (!sh.split && !gathered) { VG_TRY(gather_split_layers(ctx, pd->layer_ptr, pd->layer_len, 0, n_split)); gathered = true; }
        VG_TRY(take(next_len));
        const uint32_t* inj = nullptr;
        if (!group.empty()) {
            if (!inject_buf) VG_TRY(vg_alloc(ctx, (void**)&inject_buf, (max_h / 2) * 32));
            VG_TRY(hash_rows(ctx, group, sh.begin, sh.count, inject_buf));
            inj = inject_buf;
        }
        uint32_t* next = layer + len * 8;
        {
            KScope ks(ctx, KC_COMPRESS, (double)sh.count * (inj ? 128.0 : 96.0));
            compress_layer_kernel<<<(unsigned)((sh.count + 127) / 128), 128, 0, ctx->stream>>>(layer, inj, sh.begin, sh.count, next);
        }
        VG_LAUNCH_CHECK(ctx);
        pd->layer_ptr.push_back(next); pd->layer_len.push_back(next_len);
        if (sh.split) n_split++;
        layer = next; len = next_len;
    }
    if (!gathered) VG_TRY(gather_split_layers(ctx, pd->layer_ptr, pd->layer_len, 0, n_split));
    if (inject_buf) vg_free(ctx, inject_buf);
    if (pos != n) VG_FAIL(ctx, "commit: a matrix height does not match any tree layer");
    VG_CUDA(ctx, cudaMemcpyAsync(pd->root, layer, 32, cudaMemcpyDeviceToHost, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
