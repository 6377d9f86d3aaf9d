// kernels.cuh -- internal launchers shared by model.cu
#pragma once
#include "common.cuh"

int k_im2col(mdc_ctx* ctx, int dtype, const float* x, void* P, int B, int C, int img, int p, cudaStream_t s);
int k_set_cls_rows(mdc_ctx* ctx, float* h, const float* cls, int B, int rows_per_img, int D, cudaStream_t s);
int k_encoder_tail(mdc_ctx* ctx, int dtype, const float* h, const float* w, const float* b, float eps, const float* enc_pos,
                   float* enc_out, void* memory, int B, int n, int D, int out_dim, cudaStream_t s);
int k_add_pos(mdc_ctx* ctx, int dtype, const float* enc_out, const float* pos, void* mem, int64_t total, int64_t per_img, cudaStream_t s);
