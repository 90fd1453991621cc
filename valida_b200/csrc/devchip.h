// Device image of a vgpu_chip_desc with the LogUp randomness folded in (Montgomery words):
// alphas per interaction = r1^(bus+1) (generate_rlc_elements, machine/src/chip.rs:291-331),
// betas[j] = r2^j (chip.rs:133).
#pragma once
#include "ctx.h"

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
int32_t vg_perm_trace_enqueue(vgpu_ctx* ctx, const vgpu_chip_desc* chip, const vgpu_dmat* main, const vgpu_dmat* prep_or_null,
                              const uint32_t challenges[15], vgpu_dmat** out_perm, uint32_t* d_totals, uint32_t* n_totals);
uint32_t vg_perm_totals_ranks(const vgpu_ctx* ctx);
