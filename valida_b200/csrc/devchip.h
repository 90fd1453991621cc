// Device image of a vgpu_chip_desc with the LogUp randomness folded in (Montgomery words):
// alphas per interaction = r1^(bus+1) (generate_rlc_elements, machine/src/chip.rs:291-331),
// betas[j] = r2^j (chip.rs:133).
#pragma once
#include "ctx.h"
#include <cstddef>

struct DevPairCol {
    uint32_t constant;
    uint32_t n_terms;
    uint32_t is_prep[VGPU_MAX_TERMS], column[VGPU_MAX_TERMS], weight[VGPU_MAX_TERMS];
};
struct DevInteraction {
    uint32_t n_fields;
    DevPairCol fields[VGPU_MAX_FIELDS];
    DevPairCol count;
    uint32_t is_send;
    bb::E5 alpha;
};
struct DevChip {
    uint32_t chip_id, width, prep_width, n_interactions;
    DevInteraction interactions[VGPU_MAX_INTERACTIONS];
    bb::E5 betas[VGPU_MAX_FIELDS];
};

int32_t vg_build_devchip(vgpu_ctx* ctx, const vgpu_chip_desc* chip, const uint32_t challenges_canonical[15], DevChip* out);
int32_t vg_upload_devchip(vgpu_ctx* ctx, const vgpu_chip_desc* chip, const uint32_t challenges_canonical[15], DevChip** out_device);
int32_t vg_prefix_sum_columns(vgpu_ctx* ctx, uint32_t* data, uint64_t cs, uint64_t n, uint32_t ncols);
int32_t vg_ext_batch_inverse(vgpu_ctx* ctx, uint32_t* data, uint64_t cs, uint64_t h, uint32_t groups);

#ifdef __CUDACC__
// Stage the part of a device-resident DevChip a kernel reads (header, the first n_interactions interactions, betas) in
// shared memory: every thread of the CTA then reads descriptors as broadcast LDS instead of chains of dependent global
// loads (ncu r1b: 27-54 % of the quotient kernel's stall samples) — measured faster than the constant bank, whose
// indexed loads of a 4.6 KB structure miss the constant cache.  Caller must __syncthreads() afterwards.
__device__ __forceinline__ void devchip_to_shared(DevChip* s, const DevChip* __restrict__ g) {
    constexpr uint32_t HEAD = offsetof(DevChip, interactions) / 4, INTER = sizeof(DevInteraction) / 4;
    constexpr uint32_t BETA0 = offsetof(DevChip, betas) / 4, BETAS = sizeof(((DevChip*)nullptr)->betas) / 4;
    const uint32_t k = __ldg(&g->n_interactions);
    const uint32_t n1 = HEAD + k * INTER;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s);
    for (uint32_t i = threadIdx.x; i < n1; i += blockDim.x) dst[i] = __ldg(src + i);
    for (uint32_t i = threadIdx.x; i < BETAS; i += blockDim.x) dst[BETA0 + i] = __ldg(src + BETA0 + i);
}
#endif
