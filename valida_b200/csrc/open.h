#pragma once
#include "ctx.h"
#include <vector>

int32_t vg_inverse_denominators(vgpu_ctx* ctx, uint32_t log_H, const bb::E5& z, uint32_t* out /* 5*H words, limb-major */);
// out-of-domain evaluation in two halves: enqueue the column sums (no host sync), then host arithmetic on the copied-back sums
int32_t vg_eval_columns_enqueue(vgpu_ctx* ctx, const vgpu_dmat* lde, uint32_t npoints, const uint32_t* const* invden, uint32_t* d_out /* vg_eval_columns_words(w) */);
uint32_t vg_eval_columns_words(uint32_t w);
void vg_eval_columns_finish(const uint32_t* sums, uint64_t H, uint32_t w, uint32_t npoints, const bb::E5* z, std::vector<bb::E5>* ys /* [q][c] */);
int32_t vg_reduced_opening_accumulate(vgpu_ctx* ctx, const vgpu_dmat* lde, const bb::E5* apow_off /* alpha^(off_0 + c), host */, const bb::E5& alpha_w, uint32_t npoints,
                                      const uint32_t* const* invden, const bb::E5* sum_y, uint32_t* ro);
int32_t vg_fri_fold(vgpu_ctx* ctx, const uint32_t* cur, uint64_t n, const bb::E5& beta, const uint32_t* add_or_null, uint32_t* out);
int32_t vg_gather_words(vgpu_ctx* ctx, const std::vector<const uint32_t*>& ptrs, std::vector<uint32_t>* out);
uint32_t vg_chip_base_constraints(uint32_t chip_id);
int32_t vg_reduced_openings_complete(vgpu_ctx* ctx, uint32_t* ro, uint64_t H);   // multi-GPU: join the per-rank row ranges
