// model.cu -- orchestration of the encoder (timm VisionTransformer + channel pooling, model.py:14-23)
// and the once-per-image cross-attention K/V build.  Pure launch sequencing on the caller's stream:
// no allocation, no synchronisation, capturable in a CUDA graph.
#include "common.cuh"
#include "kernels.cuh"

namespace {

struct EncWs { float* h; void* u; void* qkv; void* a; void* mlp; void* patches; };

size_t enc_ws_bytes(const mdc_dims& d, int B) {
  size_t M = (size_t)B * (d.n_patches + 1), es = esize(d.precision);
  size_t mlp = M * d.enc_mlp * es, pat = (size_t)B * d.n_patches * d.in_chans * d.patch * d.patch * es;
  return align_up(M * d.enc_dim * 4, 256) + align_up(M * d.enc_dim * es, 256) + align_up(M * 3 * d.enc_dim * es, 256) +
         align_up(M * d.enc_dim * es, 256) + align_up(mlp > pat ? mlp : pat, 256);
}

EncWs enc_carve(const mdc_dims& d, int B, void* ws) {
  size_t M = (size_t)B * (d.n_patches + 1), es = esize(d.precision);
  char* p = (char*)ws; EncWs w;
  w.h = (float*)p; p += align_up(M * d.enc_dim * 4, 256);
  w.u = p; p += align_up(M * d.enc_dim * es, 256);
  w.qkv = p; p += align_up(M * 3 * d.enc_dim * es, 256);
  w.a = p; p += align_up(M * d.enc_dim * es, 256);
  w.mlp = p; w.patches = p;     // the patch matrix is dead before the first MLP runs
  return w;
}

}  // namespace

extern "C" size_t mdc_encode_workspace_bytes(const mdc_model* m, int B) {
  if (!m || B <= 0) return 0;
  return enc_ws_bytes(m->d, B);
}

extern "C" int mdc_encode(mdc_model* m, const float* image, int B, float* enc_out, void* memory, void* workspace,
                          size_t workspace_bytes, void* stream) {
  MDC_CHECK_ARG(m && image && workspace && B > 0 && (enc_out || memory));
  MDC_CHECK_DEVICE(m->ctx);
  MDC_CHECK_ARG(workspace_bytes >= enc_ws_bytes(m->d, B));
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d; cudaStream_t s = (cudaStream_t)stream;
  const int dt = d.precision, D = d.enc_dim, n = d.n_patches, M = B * (n + 1), Kp = d.in_chans * d.patch * d.patch;
  const void** g = m->w; const void** blk0 = m->w + MDC_ENC_GLOBAL_SLOTS;
  const void** dg = blk0 + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  EncWs w = enc_carve(d, B, workspace);
  const float attn_scale = 1.0f / sqrtf((float)(D / d.enc_heads));

  // patch embedding as a GEMM: stride == kernel, so im2col is a pure re-tiling (A.1)
  MDC_TRY(k_im2col(ctx, dt, image, w.patches, B, d.in_chans, d.img_size, d.patch, s));
  MDC_TRY(k_set_cls_rows(ctx, w.h, (const float*)g[MDC_CLS], B, n + 1, D, s));
  MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_PATCH, w.patches, Kp, g[MDC_W_PATCH], Kp, w.h, D, (const float*)g[MDC_B_PATCH],
                   (const float*)g[MDC_POS], n, B * n, D, Kp, s));
  for (int i = 0; i < d.enc_depth; ++i) {
    const void** bw = blk0 + i * MDC_ENC_BLOCK_SLOTS;
    MDC_TRY(mdc_layernorm(ctx, w.h, D, (const float*)bw[MDC_N1_W], (const float*)bw[MDC_N1_B], 1e-6f, w.u, D, dt, M, D, s));
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_BIAS, w.u, D, bw[MDC_QKV_W], D, w.qkv, 3 * D, (const float*)bw[MDC_QKV_B], nullptr, 0, M, 3 * D, D, s));
    MDC_TRY(mdc_strip_attention(ctx, dt, w.qkv, 3 * D, w.a, D, B, n + 1, d.enc_heads, D / d.enc_heads, attn_scale, 0, s));
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_LS_RESIDUAL, w.a, D, bw[MDC_PROJ_W], D, w.h, D, (const float*)bw[MDC_PROJ_B],
                     (const float*)bw[MDC_LS1], 0, M, D, D, s));
    MDC_TRY(mdc_layernorm(ctx, w.h, D, (const float*)bw[MDC_N2_W], (const float*)bw[MDC_N2_B], 1e-6f, w.u, D, dt, M, D, s));
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_BIAS_GELU, w.u, D, bw[MDC_FC1_W], D, w.mlp, d.enc_mlp, (const float*)bw[MDC_FC1_B], nullptr, 0, M,
                     d.enc_mlp, D, s));
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_LS_RESIDUAL, w.mlp, d.enc_mlp, bw[MDC_FC2_W], d.enc_mlp, w.h, D, (const float*)bw[MDC_FC2_B],
                     (const float*)bw[MDC_LS2], 0, M, D, d.enc_mlp, s));
  }
  MDC_TRY(k_encoder_tail(ctx, dt, w.h, (const float*)g[MDC_NORM_W], (const float*)g[MDC_NORM_B], 1e-6f, (const float*)dg[MDC_ENC_POS],
                         enc_out, memory, B, n, D, d.dim, s));
  return 0;
}

extern "C" int mdc_memory_from_encoder_out(mdc_model* m, const float* enc_out, int B, void* memory, void* stream) {
  MDC_CHECK_ARG(m && enc_out && memory && B > 0);
  MDC_CHECK_DEVICE(m->ctx);
  const mdc_dims& d = m->d;
  const void** dg = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  int64_t per = (int64_t)d.n_patches * d.dim;
  return k_add_pos(m->ctx, d.precision, enc_out, (const float*)dg[MDC_ENC_POS], memory, per * B, per, (cudaStream_t)stream);
}

// cross_kv buffer: the plain tensor [layer][B*S][K(dim) | V(dim)] (what the per-operation decode kernels read), followed -- when the
// fused decode kernel covers the geometry -- by the same values re-arranged per (layer, image, head, 16-key chunk) into 2 KB cells
// that the fused kernel fetches with one bulk copy each (decode_cluster.cu).
static size_t ckv_plain_bytes(const mdc_model* m, int B) {
  return align_up((size_t)m->d.dec_layers * B * m->d.n_patches * 2 * m->d.dim * esize(m->d.precision), 256);
}

extern "C" size_t mdc_cross_kv_bytes(const mdc_model* m, int B) {
  if (!m || B <= 0) return 0;
  return ckv_plain_bytes(m, B) + decode_cluster_ckv_pack_bytes(m, B);
}

extern "C" int mdc_cross_kv_build(mdc_model* m, const void* memory, int B, void* cross_kv, void* stream) {
  MDC_CHECK_ARG(m && memory && cross_kv && B > 0);
  MDC_CHECK_DEVICE(m->ctx);
  const mdc_dims& d = m->d; const int dim = d.dim, S = d.n_patches;
  const void** lw0 = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS + MDC_DEC_GLOBAL_SLOTS;
  const size_t es = esize(d.precision);
  for (int l = 0; l < d.dec_layers; ++l) {
    const void** lw = lw0 + l * MDC_DEC_LAYER_SLOTS;
    const char* Wkv = (const char*)lw[MDC_CA_IN_W] + (size_t)dim * dim * es;      // rows [dim:3dim] = k then v
    const float* bkv = (const float*)lw[MDC_CA_IN_B] + dim;
    char* out = (char*)cross_kv + (size_t)l * B * S * 2 * dim * es;
    MDC_TRY(mdc_gemm(m->ctx, d.precision, MDC_EPI_BIAS, memory, dim, Wkv, dim, out, 2 * dim, bkv, nullptr, 0, B * S, 2 * dim, dim, stream));
  }
  if (decode_cluster_ckv_pack_bytes(m, B))
    MDC_TRY(decode_cluster_ckv_pack(m, cross_kv, B, (char*)cross_kv + ckv_plain_bytes(m, B), (cudaStream_t)stream));
  return 0;
}

// decode-loop weights pre-arranged for the fused decode kernel (decode_cluster.cu): 0 bytes when it does not cover the geometry
extern "C" size_t mdc_decode_pack_bytes(const mdc_model* m) { return m ? decode_cluster_pack_bytes(m) : 0; }

extern "C" int mdc_decode_pack(mdc_model* m, void* packed, void* stream) {
  MDC_CHECK_ARG(m && packed && decode_cluster_pack_bytes(m) > 0);
  MDC_CHECK_DEVICE(m->ctx);
  return decode_cluster_pack(m, packed, (cudaStream_t)stream);
}

// AxialAttention.forward (axial_model.py:28-40): qkv = x . Wqkv^T (no bias) -> strip attention with
// scale 0.125 (dim_head default, regardless of real head width) -> . Wout^T + bout.
extern "C" size_t mdc_axial_workspace_bytes(const mdc_model* m, int B, int n) {
  if (!m || B <= 0 || n <= 0) return 0;
  size_t rows = (size_t)B * n, es = esize(m->d.precision);
  return align_up(rows * m->d.dim * es, 256) + align_up(rows * 3 * m->d.dim * es, 256) + align_up(rows * m->d.dim * es, 256);
}

__global__ void cast_rows_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void uncast_rows_kernel(const bf16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = __bfloat162float(in[i]);
}

extern "C" int mdc_axial_attention(mdc_model* m, const float* x, int B, int n, int softmax_over_queries, float* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  MDC_CHECK_ARG(m && x && out && workspace && B > 0 && n > 0 && m->d.has_axial);
  MDC_CHECK_DEVICE(m->ctx);
  MDC_CHECK_ARG(workspace_bytes >= mdc_axial_workspace_bytes(m, B, n));
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d; cudaStream_t s = (cudaStream_t)stream;
  const int dim = d.dim, heads = 8, dt = d.precision; const size_t es = esize(dt);   // axial_model.py:20 heads=8
  MDC_CHECK_ARG(dim % heads == 0);
  const void** dg = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  MDC_CHECK_ARG(dg[MDC_AX_QKV_W] && dg[MDC_AX_OUT_W] && dg[MDC_AX_OUT_B]);
  size_t rows = (size_t)B * n;
  char* p = (char*)workspace;
  void* xin = p; p += align_up(rows * dim * es, 256);
  void* qkv = p; p += align_up(rows * 3 * dim * es, 256);
  void* att = p;
  const void* a_in = x;
  if (dt == MDC_BF16) {
    cast_rows_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(x, (bf16*)xin, (int64_t)rows * dim); MDC_LAUNCH_CHECK(ctx);
    a_in = xin;
  }
  MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_BIAS, a_in, dim, dg[MDC_AX_QKV_W], dim, qkv, 3 * dim, nullptr, nullptr, 0, (int)rows, 3 * dim, dim, s));
  MDC_TRY(mdc_strip_attention(ctx, dt, qkv, 3 * dim, att, dim, B, n, heads, dim / heads, 0.125f, softmax_over_queries, s));
  if (dt == MDC_F32) {
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_BIAS, att, dim, dg[MDC_AX_OUT_W], dim, out, dim, (const float*)dg[MDC_AX_OUT_B], nullptr, 0, (int)rows, dim, dim, s));
  } else {
    MDC_TRY(mdc_gemm(ctx, dt, MDC_EPI_BIAS, att, dim, dg[MDC_AX_OUT_W], dim, xin, dim, (const float*)dg[MDC_AX_OUT_B], nullptr, 0, (int)rows, dim, dim, s));
    uncast_rows_kernel<<<ctx->sm_count * 4, 256, 0, s>>>((const bf16*)xin, out, (int64_t)rows * dim); MDC_LAUNCH_CHECK(ctx);
  }
  return 0;
}

// axial_model.Decoder.forward front end (axial_model.py:100-103):
//   out[b,i,:] = AxialAttention(embedding[tokens])[b,i,:] + pos[i,:]
__global__ void gather_rows_kernel(const float* __restrict__ emb, const int32_t* __restrict__ tokens, int tokens_ld, int n, int dim,
                                   float* __restrict__ out, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % dim); int64_t r = i / dim; int b = (int)(r / n), t = (int)(r % n);
    out[i] = emb[(int64_t)tokens[(int64_t)b * tokens_ld + t] * dim + c];
  }
}
__global__ void add_rows_kernel(float* __restrict__ x, const float* __restrict__ pos, int64_t total, int64_t per_img) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) x[i] += pos[i % per_img];
}

extern "C" size_t mdc_axial_embed_workspace_bytes(const mdc_model* m, int B, int n) {
  if (!m || B <= 0 || n <= 0) return 0;
  return mdc_axial_workspace_bytes(m, B, n) + align_up((size_t)B * n * m->d.dim * sizeof(float), 256);
}

extern "C" int mdc_axial_embed(mdc_model* m, const int32_t* tokens, int tokens_ld, int B, int n, const float* pos,
                               int softmax_over_queries, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  MDC_CHECK_ARG(m && tokens && out && workspace && B > 0 && n > 0 && tokens_ld >= n);
  MDC_CHECK_DEVICE(m->ctx);
  MDC_CHECK_ARG(workspace_bytes >= mdc_axial_embed_workspace_bytes(m, B, n));
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d; cudaStream_t s = (cudaStream_t)stream;
  const void** dg = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  if (!pos) { MDC_CHECK_ARG(n == d.max_pos); pos = (const float*)dg[MDC_DEC_POS]; }
  size_t axb = mdc_axial_workspace_bytes(m, B, n);
  float* x = (float*)((char*)workspace + axb);
  int64_t total = (int64_t)B * n * d.dim;
  gather_rows_kernel<<<ctx->sm_count * 4, 256, 0, s>>>((const float*)dg[MDC_EMB], tokens, tokens_ld, n, d.dim, x, total); MDC_LAUNCH_CHECK(ctx);
  MDC_TRY(mdc_axial_attention(m, x, B, n, softmax_over_queries, out, workspace, axb, stream));
  add_rows_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(out, pos, total, (int64_t)n * d.dim); MDC_LAUNCH_CHECK(ctx);
  return 0;
}
