// <ValMmcs as Mmcs>::ProverData, device resident: committed (bit-reversed) LDE matrices in the
// caller's order + every digest layer (canonical words, 8 per node).
#pragma once
#include "ctx.h"
#include <functional>

struct vgpu_prover_data {
    vgpu_ctx* ctx = nullptr;
    std::vector<vgpu_dmat*> ldes;          // owned
    uint32_t* digests = nullptr;           // all layers, leaf layer first
    std::vector<uint32_t*> layer_ptr;      // layer_ptr[i] -> layer i (len layer_len[i] digests)
    std::vector<uint64_t> layer_len;
    uint64_t max_height = 0;
    uint32_t root[8] = {0};
};

// heights[i] = height of committed matrix i.  `need(group)` is called with the indices of one height group right before
// its rows are hashed and must fill pd->ldes[i] (the commit extends a matrix only when the tree reaches its height, so
// that uploads of shorter matrices and hashing of taller ones overlap); null = every pd->ldes[i] is already there.
int32_t vg_merkle_build(vgpu_ctx* ctx, vgpu_prover_data* pd, const std::vector<uint64_t>& heights,
                        const std::function<int32_t(const std::vector<size_t>&)>& need);
int32_t vg_fri_layer_commit(vgpu_ctx* ctx, const uint32_t* v, uint64_t cs, uint64_t npairs, uint32_t* digests,
                            std::vector<uint32_t*>* layer_ptr, std::vector<uint64_t>* layer_len, uint32_t root_out[8]);
