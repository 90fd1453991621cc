#pragma once
#include "ctx.h"
#include <vector>

int32_t vg_inverse_denominators(vgpu_ctx* ctx, uint32_t log_H, const bb::E5& z, uint32_t* out /* 5*H words, limb-major */);
int32_t vg_eval_columns(vgpu_ctx* ctx, const vgpu_dmat* lde, uint32_t npoints, const bb::E5* z, const uint32_t* const* invden, std::vector<bb::E5>* ys);
int32_t vg_reduced_opening_accumulate(vgpu_ctx* ctx, const vgpu_dmat* lde, const bb::E5* d_apow, uint32_t npoints, const uint32_t* const* invden,
                                      const bb::E5* sum_y, const bb::E5* alpha_off, uint32_t* ro);
int32_t vg_fri_fold(vgpu_ctx* ctx, const uint32_t* cur, uint64_t n, const bb::E5& beta, const uint32_t* add_or_null, uint32_t* out);
int32_t vg_gather_words(vgpu_ctx* ctx, const std::vector<const uint32_t*>& ptrs, std::vector<uint32_t>* out);
uint32_t vg_chip_base_constraints(uint32_t chip_id);
int32_t vg_reduced_openings_complete(vgpu_ctx* ctx, uint32_t* ro, uint64_t H);   // multi-GPU: join the per-rank row ranges
