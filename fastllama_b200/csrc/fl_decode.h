// fl_decode.h -- internal launcher interface of the fused decode kernels (fl_decode_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "fl_cuda.h"

int flk_mv_fused_supported(int type, int K, int mtot);
int flk_mv_fused(cudaStream_t st, const fl_mv_args *args);
int flk_attn_decode(cudaStream_t st, const float *q, const float *kcache, const float *vcache, float *out, const int *n_past,
                    int k_row_stride, int n_head, int head_dim, int n_ctx, float scale, const uint16_t *exp_tab);

// fl_token_kernel.cu: the persistent per-token kernel
int flk_token_plan_create(const fl_token_step *steps, int n_steps, const uint16_t *silu_tab, const uint16_t *exp_tab, const void *rope_cs, unsigned *epoch_counter, void **out);
int flk_token_plan_launch(cudaStream_t st, void *plan);
int flk_token_plan_destroy(void *plan);
int flk_token_plan_profile(void *plan, unsigned long long *out, size_t max_words, int *n_ctas);
int flk_token_plan_error(void *plan);
int flk_token_plan_profile2(void *plan, unsigned *out, size_t max_words);
// fl_runtime.cu: the peer-mapped buffers of fl_comm_shared_alloc (nullptr when there are none)
const void *const *fl_shared_peers(int *rank, int *world);
