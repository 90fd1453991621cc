// <ValMmcs as Mmcs>::ProverData, device resident: committed (bit-reversed) LDE matrices in the
// caller's order + the digest layers (canonical words, 8 per node).
// Split proof: a layer of more than comm_size nodes is cut into comm_size contiguous runs and a rank computes and KEEPS
// only its run (its sub-tree); the layer of exactly comm_size nodes — the sub-roots — is all-gathered (comm_size x 32 bytes,
// the only collective of a tree) and the layers above it are computed by every rank.
#pragma once
#include "ctx.h"
#include <functional>

struct VgTree {
    uint32_t* digests = nullptr;           // the stored parts of all layers, leaf layer first
    std::vector<uint32_t*> layer_ptr;      // layer_ptr[i] -> node layer_begin[i] of layer i
    std::vector<uint64_t> layer_len;       // nodes of the whole layer
    std::vector<uint64_t> layer_begin, layer_count;   // the run of nodes stored on this rank
    // address of node `j` of layer `lvl` if THIS rank is the one that reports it in a query answer (the owner of a split
    // layer's run; rank 0 for the layers every rank holds), else null
    const uint32_t* node(const vgpu_ctx* ctx, size_t lvl, uint64_t j) const {
        if (layer_count[lvl] == layer_len[lvl]) return ctx->comm_rank == 0 || !vg_sharded(ctx) ? layer_ptr[lvl] + j * 8 : nullptr;
        return j >= layer_begin[lvl] && j < layer_begin[lvl] + layer_count[lvl] ? layer_ptr[lvl] + (j - layer_begin[lvl]) * 8 : nullptr;
    }
};

struct vgpu_prover_data {
    vgpu_ctx* ctx = nullptr;
    std::vector<vgpu_dmat*> ldes;          // owned; VG_ROWS shards for the tall matrices of a split proof
    VgTree tree;
    uint64_t max_height = 0;
    uint32_t root[8] = {0};
};

// heights[i] = height of committed matrix i.  `need(group)` is called with the indices of one height group right before
// its rows are hashed and must fill pd->ldes[i] (the commit extends a matrix only when the tree reaches its height, so
// that uploads of shorter matrices and hashing of taller ones overlap); null = every pd->ldes[i] is already there.
int32_t vg_merkle_build(vgpu_ctx* ctx, vgpu_prover_data* pd, const std::vector<uint64_t>& heights,
                        const std::function<int32_t(const std::vector<size_t>&)>& need);
// Tree over the sibling pairs of an ext5 vector (p3-fri commit phase).  v: limb-major, limb stride cs, holding pairs
// [pair0, pair0 + local_pairs) of the npairs of the layer (all of them, or this rank's run).  The root arrives in root_out
// once the context's stream has been synchronised (the caller needs it for the transcript anyway).
int32_t vg_fri_layer_commit(vgpu_ctx* ctx, const uint32_t* v, uint64_t cs, uint64_t npairs, bool v_is_shard, VgTree* tree, uint32_t root_out[8]);
void vg_tree_free(vgpu_ctx* ctx, VgTree* t);
