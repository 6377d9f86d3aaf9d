// decode.cu -- one autoregressive decode step of the transformer decoder (Decoder.predict /
// nn.TransformerDecoder post-norm layers, model.py:92-127 + torch transformer.py:1158-1197),
// computed ONLY for the new position t (the reference recomputes all 99 positions and the whole
// encoder every step).
//
// Data flow per layer (all activations f32, [B, *]; weights/KV `precision`):
//   qkv  = LNload(x) . Ws^T + bs          -> append k,v to the paged self-KV cache, slot t
//   o    = softmax(q.K[0..t]/sqrt(hd) + padbias) V[0..t]        (padbias: +1.0 at PAD keys, Q7)
//   y1   = o . Wso^T + bso
//   qc   = LNload(xa + y1; norm1) . Wcq^T + bcq
//   oc   = softmax(qc.Kc^T/sqrt(hd)) Vc   (cross K/V built once per image, HBM/L2 resident)
//   y2   = oc . Wco^T + bco
//   f1   = relu(LNload(xb + y2; norm2) . W1^T + b1)
//   y3   = f1 . W2^T + b2                 ; next layer / head consumes LNload(xc + y3; norm3)
// "LNload": residual add + LayerNorm are fused into the CONSUMER's operand load (every CTA
// re-normalises the few rows it needs; CTA column 0 also publishes the normalised rows as the
// next residual), so no projection needs a full-row epilogue and all of them spread over the SMs.
// The head kernel fuses LN3 + vocab projection + greedy / top-k / top-p select + max-prob.
#include "common.cuh"
#include <float.h>

namespace {

constexpr int LIN_THREADS = 256;
constexpr int LIN_WARPS = 8;
constexpr int BT = 16;           // batch rows per CTA

enum { XMODE_PLAIN = 0, XMODE_LN = 1, XMODE_EMBED = 2 };

struct XSrc {
  int mode;
  const float* x; int64_t ldx;                  // PLAIN
  const float* resid; const float* delta;       // LN: LN(resid + delta) rows of width K
  const float* ln_w; const float* ln_b; float eps;
  const int32_t* tokens; int tokens_ld; int t;  // EMBED: emb[tokens[b,t]] + pos[t]
  const float* emb; const float* pos;
  float* xn_out;                                // where CTA column 0 publishes the operand rows (or NULL)
};

// Fill Xs[BT][K] (f32) for batch rows b0..b0+BT-1.
__device__ __forceinline__ void load_x_tile(const XSrc& xs, float* Xs, int b0, int B, int K, bool publish) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < BT; r += LIN_WARPS) {
    const int b = b0 + r;
    float* row = Xs + (size_t)r * K;
    if (b >= B) { for (int c = lane; c < K; c += 32) row[c] = 0.f; continue; }
    if (xs.mode == XMODE_PLAIN) {
      const float* src = xs.x + (int64_t)b * xs.ldx;
      for (int c = lane * 4; c < K; c += 128) *reinterpret_cast<float4*>(row + c) = *reinterpret_cast<const float4*>(src + c);
    } else if (xs.mode == XMODE_EMBED) {
      const int tok = xs.tokens[(int64_t)b * xs.tokens_ld + xs.t];
      const float* e = xs.emb + (int64_t)tok * K; const float* p = xs.pos + (int64_t)xs.t * K;
      for (int c = lane; c < K; c += 32) row[c] = e[c] + p[c];
    } else {
      const float* a = xs.resid + (int64_t)b * K; const float* d = xs.delta + (int64_t)b * K;
      float s = 0.f;
      for (int c = lane; c < K; c += 32) { float v = a[c] + d[c]; row[c] = v; s += v; }
      float mean = warp_sum(s) / (float)K;
      float q = 0.f;
      for (int c = lane; c < K; c += 32) { float v = row[c] - mean; q += v * v; }
      float rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + xs.eps);
      for (int c = lane; c < K; c += 32) row[c] = (row[c] - mean) * rstd * xs.ln_w[c] + xs.ln_b[c];
    }
    if (publish && xs.xn_out) {
      __syncwarp();
      for (int c = lane; c < K; c += 32) xs.xn_out[(int64_t)b * K + c] = row[c];
    }
  }
}

// Y[B,N] = act(Xeff[B,K] . W[N,K]^T + bias).  grid = (N / (LIN_WARPS*RW), ceil(B/BT)).
// A warp owns RW output columns; lanes split K in 8-element (128-bit) slices.
template <typename TW, int RW, bool RELU>
__global__ void __launch_bounds__(LIN_THREADS) dec_linear_kernel(XSrc xs, const TW* __restrict__ W, const float* __restrict__ bias,
                                                                  float* __restrict__ Y, int64_t ldy, int B, int N, int K) {
  extern __shared__ __align__(16) float Xs[];
  const int b0 = blockIdx.y * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * LIN_WARPS + warp) * RW;
  // issue the weight loads for the first K slice before the operand tile is built (overlaps latency)
  load_x_tile(xs, Xs, b0, B, K, blockIdx.x == 0);
  __syncthreads();
  float acc[RW][BT];
#pragma unroll
  for (int c = 0; c < RW; ++c)
#pragma unroll
    for (int r = 0; r < BT; ++r) acc[c][r] = 0.f;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    float w[RW][8];
#pragma unroll
    for (int c = 0; c < RW; ++c) {
      if (n0 + c < N) load8(W + (int64_t)(n0 + c) * K + k0, w[c]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[c][j] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < BT; ++r) {
      float4 xa = *reinterpret_cast<const float4*>(Xs + (size_t)r * K + k0);
      float4 xb = *reinterpret_cast<const float4*>(Xs + (size_t)r * K + k0 + 4);
#pragma unroll
      for (int c = 0; c < RW; ++c) {
        float a = acc[c][r];
        a = fmaf(w[c][0], xa.x, a); a = fmaf(w[c][1], xa.y, a); a = fmaf(w[c][2], xa.z, a); a = fmaf(w[c][3], xa.w, a);
        a = fmaf(w[c][4], xb.x, a); a = fmaf(w[c][5], xb.y, a); a = fmaf(w[c][6], xb.z, a); a = fmaf(w[c][7], xb.w, a);
        acc[c][r] = a;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < RW; ++c) {
    const int n = n0 + c;
    const float bv = (n < N && bias) ? bias[n] : 0.f;
#pragma unroll
    for (int r = 0; r < BT; ++r) {
      float v = warp_sum(acc[c][r]);
      if (lane == (r & 31) && n < N && b0 + r < B) {
        v += bv;
        if (RELU) v = fmaxf(v, 0.f);
        Y[(int64_t)(b0 + r) * ldy + n] = v;
      }
    }
  }
}

// ---- attention of ONE query per (image, head) over a key/value set ---------------------------------
// warp == head.  Scores go to shared memory; softmax max/sum are warp shuffles; P.V splits the warp
// into (key subgroup, 8-channel chunk) so every V access is a 128-bit load.
template <typename TKV, typename KeyPtr, typename Bias>
__device__ __forceinline__ void attend_one(const float* q /*smem, hd, pre-scaled*/, float* sc /*smem, nkeys*/, int nkeys, int hd,
                                           KeyPtr kv_ptr, Bias bias, float* out /*global or smem, hd*/) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int u = lane; u < nkeys; u += 32) {
    const TKV* kp = kv_ptr(u, 0);
    float s = 0.f;
    for (int c = 0; c < hd; c += 8) {
      float kv[8]; load8(kp + c, kv);
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(q[c + j], kv[j], s);
    }
    s += bias(u);
    sc[u] = s; mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int u = lane; u < nkeys; u += 32) { float e = expf(sc[u] - mx); sc[u] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  const int DC = hd / 8, KS = 32 / DC;       // hd in {32,64,128} -> DC in {4,8,16}
  const int dc = lane % DC, ks = lane / DC;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int u = ks; u < nkeys; u += KS) {
    float p = sc[u];
    float vv[8]; load8(kv_ptr(u, 1) + dc * 8, vv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, vv[j], acc[j]);
  }
  for (int o = DC; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (ks == 0) {
    float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) out[dc * 8 + j] = acc[j] * inv;
  }
  __syncwarp();
}

// self-attention: appends this step's k,v to the paged cache, then attends over slots 0..t.
// pool layout: [page][layer][k|v][page_tokens][d]
template <typename TKV>
__global__ void dec_self_attn_kernel(const float* __restrict__ qkv /*[B,3d]*/, TKV* __restrict__ pool,
                                     const int32_t* __restrict__ page_table, int pages_per_seq, int PT, int n_layers, int layer,
                                     const int32_t* __restrict__ tokens, int tokens_ld, int pad_idx, int t, int d, int hd,
                                     float scale, float* __restrict__ o /*[B,d]*/) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31, heads = blockDim.x >> 5;
  float* q = sm + head * hd;
  float* sc = sm + heads * hd + head * (t + 1);
  const float* row = qkv + (int64_t)b * 3 * d;
  const int32_t* pt = page_table + (int64_t)b * pages_per_seq;
  const int64_t plane = (int64_t)PT * d;                    // one k or v plane of a page/layer
  {  // append k_t, v_t (this head's channels) and stage q
    TKV* kdst = pool + (((int64_t)pt[t / PT] * n_layers + layer) * 2) * plane + (int64_t)(t % PT) * d + head * hd;
    for (int c = lane; c < hd; c += 32) {
      q[c] = row[head * hd + c] * scale;
      kdst[c] = from_f<TKV>(row[d + head * hd + c]);
      kdst[plane + c] = from_f<TKV>(row[2 * d + head * hd + c]);
    }
  }
  __syncwarp();
  auto kv_ptr = [&](int u, int which) -> const TKV* {
    return pool + (((int64_t)pt[u / PT] * n_layers + layer) * 2 + which) * plane + (int64_t)(u % PT) * d + head * hd;
  };
  auto bias = [&](int u) -> float { return tokens[(int64_t)b * tokens_ld + u] == pad_idx ? 1.0f : 0.0f; };
  attend_one<TKV>(q, sc, t + 1, hd, kv_ptr, bias, o + (int64_t)b * d + head * hd);
}

// cross-attention over the S memory keys of image b; cross_kv layer plane: [B*S][2d] (K | V)
template <typename TKV>
__global__ void dec_cross_attn_kernel(const float* __restrict__ qc /*[B,d]*/, const TKV* __restrict__ ckv_layer, int S, int d,
                                      int hd, float scale, float* __restrict__ o) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31, heads = blockDim.x >> 5;
  float* q = sm + head * hd;
  float* sc = sm + heads * hd + head * S;
  for (int c = lane; c < hd; c += 32) q[c] = qc[(int64_t)b * d + head * hd + c] * scale;
  __syncwarp();
  const TKV* base = ckv_layer + (int64_t)b * S * 2 * d + head * hd;
  auto kv_ptr = [&](int u, int which) -> const TKV* { return base + (int64_t)u * 2 * d + which * d; };
  auto bias = [&](int) -> float { return 0.f; };
  attend_one<TKV>(q, sc, S, hd, kv_ptr, bias, o + (int64_t)b * d + head * hd);
}

// ---- select -------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;

__device__ __forceinline__ void block_argmax(float v, int idx, float* s_val, int* s_idx, float& out_v, int& out_i) {
  // max value, lowest index on ties (torch.argmax returns the first maximal index)
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o); int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
  __syncthreads();
  if (warp == 0) {
    v = lane < (SEL_THREADS / 32) ? s_val[lane] : -INFINITY; idx = lane < (SEL_THREADS / 32) ? s_idx[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o); int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { s_val[0] = v; s_idx[0] = idx; }
  }
  __syncthreads();
  out_v = s_val[0]; out_i = s_idx[0];
  __syncthreads();
}

__device__ __forceinline__ double block_sum_d(double v, double* s_d) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_d[warp] = v;
  __syncthreads();
  double tot = 0.0;
  for (int i = 0; i < SEL_THREADS / 32; ++i) tot += s_d[i];
  __syncthreads();
  return tot;
}

// lg: V logits in shared memory (may be overwritten with the filtered logits); srt: scratch of
// next_pow2(V) floats.  Returns (token, conf) in thread 0.
// Semantics: transformers top_k_top_p_filtering (inference_p.py:83) -> conf = max softmax prob of the
// filtered logits (inference_p.py:84-86) -> greedy argmax (inference_p.py:77) or inverse-CDF draw with u.
__device__ void select_from_logits(float* lg, float* srt, int V, int Vp2, int top_k, float top_p, bool sample, float u,
                                   int& token, float& conf) {
  __shared__ float s_val[SEL_THREADS / 32]; __shared__ int s_idx[SEL_THREADS / 32];
  __shared__ double s_d[SEL_THREADS / 32]; __shared__ double s_scan[SEL_THREADS]; __shared__ int s_first;
  const int tid = threadIdx.x;
  if (top_k > 0 || top_p < 1.0f) {
    for (int i = tid; i < Vp2; i += SEL_THREADS) srt[i] = i < V ? lg[i] : -INFINITY;
    __syncthreads();
    for (int k = 2; k <= Vp2; k <<= 1)            // bitonic sort, descending
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < Vp2; i += SEL_THREADS) {
          int ixj = i ^ j;
          if (ixj > i) {
            float a = srt[i], b = srt[ixj];
            bool desc = (i & k) == 0;
            if (desc ? (a < b) : (a > b)) { srt[i] = b; srt[ixj] = a; }
          }
        }
        __syncthreads();
      }
    float cut = -INFINITY;
    int kept = V;
    if (top_k > 0) { int k = min(max(top_k, 1), V); cut = srt[k - 1]; }
    if (top_k > 0) {  // entries strictly below the k-th largest are removed (ties at the k-th kept)
      int cnt = 0;
      for (int i = tid; i < V; i += SEL_THREADS) cnt += (srt[i] >= cut);
      double c = block_sum_d((double)cnt, s_d);
      kept = (int)c;
    }
    if (top_p < 1.0f) {
      // ascending cumulative softmax over the kept entries = suffix sums of the descending array
      float mx = srt[0];
      double part = 0.0;
      for (int i = tid; i < kept; i += SEL_THREADS) part += (double)expf(srt[i] - mx);
      double total = block_sum_d(part, s_d);
      // thread 0 walks from the smallest kept entry upwards (V <= 4096; kept is small after top-k)
      if (tid == 0) {
        float cum = 0.f; int removed = 0;
        for (int i = kept - 1; i >= 1; --i) {      // never remove the largest (min_tokens_to_keep = 1)
          cum += (float)((double)expf(srt[i] - mx) / total);
          if (cum <= 1.0f - top_p) removed++; else break;
        }
        s_first = kept - removed;                  // number of entries kept from the top
      }
      __syncthreads();
      int nk = s_first;
      cut = fmaxf(cut, srt[nk - 1]);
      __syncthreads();
    }
    for (int i = tid; i < V; i += SEL_THREADS) if (lg[i] < cut) lg[i] = -INFINITY;
    __syncthreads();
  }
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int i = tid; i < V; i += SEL_THREADS) { float v = lg[i]; if (v > bv) { bv = v; bi = i; } }
  float mx; int amax;
  block_argmax(bv, bi, s_val, s_idx, mx, amax);
  // softmax denominator in fp32 (conf) and, for sampling, probabilities in double (matches the oracle's draw)
  const int EPT = (V + SEL_THREADS - 1) / SEL_THREADS;
  double local = 0.0; float localf = 0.f;
  for (int j = 0; j < EPT; ++j) {
    int i = tid * EPT + j;
    if (i < V) { float e = expf(lg[i] - mx); localf += e; if (sample) local += exp((double)lg[i] - (double)mx); }
  }
  double totf = block_sum_d((double)localf, s_d);
  float cf = 1.0f / (float)totf;
  int tok = amax;
  if (sample) {
    s_scan[tid] = local;
    __syncthreads();
    if (tid == 0) { double run = 0.0; for (int i = 0; i < SEL_THREADS; ++i) { double v = s_scan[i]; s_scan[i] = run; run += v; } s_d[0] = run; s_first = V - 1; }
    __syncthreads();
    double total = s_d[0], thr = (double)u * total, run = s_scan[tid];
    int mine = 0x7fffffff;
    for (int j = 0; j < EPT; ++j) {
      int i = tid * EPT + j;
      if (i < V) { run += exp((double)lg[i] - (double)mx); if (run > thr && mine == 0x7fffffff) mine = i; }
    }
    if (mine != 0x7fffffff) atomicMin(&s_first, mine);
    __syncthreads();
    tok = s_first;
    __syncthreads();
  }
  token = tok; conf = cf;
}

// head: logits = LN3(xc + y3) . Wout^T + bout, then select.  One CTA per image.
template <typename TW>
__global__ void __launch_bounds__(SEL_THREADS) dec_head_select_kernel(XSrc xs, const TW* __restrict__ Wout, const float* __restrict__ bout,
                                                                       int V, int Vp2, int K, int t, float* __restrict__ logits_out,
                                                                       int64_t logits_img_stride, int logits_row, int32_t* __restrict__ tokens,
                                                                       int tokens_ld, int forced, float* __restrict__ confs, int confs_ld,
                                                                       const float* __restrict__ uniforms, int uniforms_ld, int top_k, float top_p) {
  extern __shared__ __align__(16) float sm[];
  float* x = sm;              // K
  float* lg = sm + K;         // V
  float* srt = lg + V;        // Vp2
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const float* a = xs.resid + (int64_t)b * K; const float* d = xs.delta + (int64_t)b * K;
    float s = 0.f;
    for (int c = lane; c < K; c += 32) { float v = a[c] + d[c]; x[c] = v; s += v; }
    float mean = warp_sum(s) / (float)K;
    float q = 0.f;
    for (int c = lane; c < K; c += 32) { float v = x[c] - mean; q += v * v; }
    float rstd = 1.0f / sqrtf(warp_sum(q) / (float)K + xs.eps);
    for (int c = lane; c < K; c += 32) x[c] = (x[c] - mean) * rstd * xs.ln_w[c] + xs.ln_b[c];
  }
  __syncthreads();
  for (int n = warp; n < V; n += SEL_THREADS / 32) {
    float acc = 0.f;
    for (int k0 = lane * 8; k0 < K; k0 += 256) {
      float w[8]; load8(Wout + (int64_t)n * K + k0, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(w[j], x[k0 + j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + bout[n];
      lg[n] = v;
      if (logits_out) logits_out[(int64_t)b * logits_img_stride + (int64_t)logits_row * V + n] = v;
    }
  }
  __syncthreads();
  if (forced && !(confs && (t % 4 == 0))) return;
  const bool sample = (top_k != 0 || top_p != 1.0f) && uniforms != nullptr;
  float u = sample ? uniforms[(int64_t)b * uniforms_ld + t] : 0.f;
  int token; float conf;
  select_from_logits(lg, srt, V, Vp2, top_k, top_p, sample, u, token, conf);
  if (threadIdx.x == 0) {
    if (!forced) tokens[(int64_t)b * tokens_ld + t + 1] = token;
    if (confs && (t % 4 == 0)) confs[(int64_t)b * confs_ld + t / 4] = conf;
  }
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const float* __restrict__ logits, int64_t ld, int V, int Vp2, int top_k,
                                                             float top_p, const float* __restrict__ uniforms, int32_t* __restrict__ token_out,
                                                             float* __restrict__ conf_out) {
  extern __shared__ __align__(16) float sm[];
  float* lg = sm; float* srt = sm + V;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < V; i += SEL_THREADS) lg[i] = logits[(int64_t)b * ld + i];
  __syncthreads();
  const bool sample = (top_k != 0 || top_p != 1.0f) && uniforms != nullptr;
  int token; float conf;
  select_from_logits(lg, srt, V, Vp2, top_k, top_p, sample, sample ? uniforms[b] : 0.f, token, conf);
  if (threadIdx.x == 0) { if (token_out) token_out[b] = token; if (conf_out) conf_out[b] = conf; }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <typename TW>
int launch_linear(mdc_ctx* ctx, const XSrc& xs, const void* W, const float* bias, float* Y, int64_t ldy, int B, int N, int K,
                  bool relu, cudaStream_t s) {
  MDC_CHECK_ARG(K % 8 == 0);
  size_t smem = (size_t)BT * K * sizeof(float);
  int rw = N >= 1536 ? 4 : (N >= 512 ? 2 : 1);
  dim3 grid((N + LIN_WARPS * rw - 1) / (LIN_WARPS * rw), (B + BT - 1) / BT), block(LIN_THREADS);
#define MDC_LIN(RW_, RELU_)                                                                                                  \
  {                                                                                                                          \
    MDC_ENSURE_SMEM((dec_linear_kernel<TW, RW_, RELU_>), smem);                                                              \
    dec_linear_kernel<TW, RW_, RELU_><<<grid, block, smem, s>>>(xs, (const TW*)W, bias, Y, ldy, B, N, K);                       \
  }
  if (rw == 4) { if (relu) MDC_LIN(4, true) else MDC_LIN(4, false) }
  else if (rw == 2) { if (relu) MDC_LIN(2, true) else MDC_LIN(2, false) }
  else { if (relu) MDC_LIN(1, true) else MDC_LIN(1, false) }
#undef MDC_LIN
  MDC_LAUNCH_CHECK(ctx); return 0;
}

struct Scratch { float *xa, *xb, *xc, *y1, *y2, *y3, *qkv, *qc, *o, *oc, *f1; };

size_t scratch_floats(const mdc_dims& d, int B) {
  return (size_t)B * ((size_t)d.dim * 9 + (size_t)d.dim * 3 + d.dec_ffn);
}

Scratch carve(const mdc_dims& d, int B, void* p) {
  float* f = (float*)p; Scratch s; size_t bd = (size_t)B * d.dim;
  s.xa = f; f += bd; s.xb = f; f += bd; s.xc = f; f += bd; s.y1 = f; f += bd; s.y2 = f; f += bd; s.y3 = f; f += bd;
  s.qc = f; f += bd; s.o = f; f += bd; s.oc = f; f += bd; s.qkv = f; f += 3 * bd; s.f1 = f;
  return s;
}

template <typename T>
int decode_step_typed(mdc_model* m, const mdc_decode_state* st, int t, cudaStream_t s) {
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d;
  const int B = st->B, dim = d.dim, hd = dim / d.dec_heads;
  const void** gw = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  const void** lw0 = gw + MDC_DEC_GLOBAL_SLOTS;
  Scratch sc = carve(d, B, st->scratch);
  const float scale = 1.0f / sqrtf((float)hd);
  XSrc prev{};   // how the next consumer obtains the layer input
  if (st->x_override) { prev.mode = XMODE_PLAIN; prev.x = st->x_override + (int64_t)t * dim; prev.ldx = (int64_t)st->x_override_ld * dim; }
  else {
    prev.mode = XMODE_EMBED; prev.tokens = st->tokens; prev.tokens_ld = st->tokens_ld; prev.t = t;
    prev.emb = (const float*)gw[MDC_EMB]; prev.pos = st->pos_override ? st->pos_override : (const float*)gw[MDC_DEC_POS];
  }
  for (int l = 0; l < d.dec_layers; ++l) {
    const void** lw = lw0 + l * MDC_DEC_LAYER_SLOTS;
    // qkv = LNload(prev) . Ws^T + bs; publishes xa
    XSrc x1 = prev; x1.xn_out = sc.xa;
    MDC_TRY(launch_linear<T>(ctx, x1, lw[MDC_SA_IN_W], (const float*)lw[MDC_SA_IN_B], sc.qkv, 3 * dim, B, 3 * dim, dim, false, s));
    {
      size_t smem = (size_t)d.dec_heads * (hd + t + 1) * sizeof(float);
      MDC_ENSURE_SMEM(dec_self_attn_kernel<T>, smem);
      dec_self_attn_kernel<T><<<B, d.dec_heads * 32, smem, s>>>(sc.qkv, (T*)st->kv_pool, st->page_table, st->pages_per_seq, d.page_tokens,
                                                               d.dec_layers, l, st->tokens, st->tokens_ld, d.pad_idx, t, dim, hd, scale, sc.o);
      MDC_LAUNCH_CHECK(ctx);
    }
    XSrc xo{}; xo.mode = XMODE_PLAIN; xo.x = sc.o; xo.ldx = dim;
    MDC_TRY(launch_linear<T>(ctx, xo, lw[MDC_SA_OUT_W], (const float*)lw[MDC_SA_OUT_B], sc.y1, dim, B, dim, dim, false, s));
    // cross-attention query from LN1(xa + y1); publishes xb
    XSrc x2{}; x2.mode = XMODE_LN; x2.resid = sc.xa; x2.delta = sc.y1; x2.ln_w = (const float*)lw[MDC_LN1_W]; x2.ln_b = (const float*)lw[MDC_LN1_B];
    x2.eps = 1e-5f; x2.xn_out = sc.xb;
    MDC_TRY(launch_linear<T>(ctx, x2, lw[MDC_CA_IN_W], (const float*)lw[MDC_CA_IN_B], sc.qc, dim, B, dim, dim, false, s));
    {
      size_t smem = (size_t)d.dec_heads * (hd + d.n_patches) * sizeof(float);
      const T* ckv = (const T*)st->cross_kv + (int64_t)l * B * d.n_patches * 2 * dim;
      MDC_ENSURE_SMEM(dec_cross_attn_kernel<T>, smem);
      dec_cross_attn_kernel<T><<<B, d.dec_heads * 32, smem, s>>>(sc.qc, ckv, d.n_patches, dim, hd, scale, sc.oc);
      MDC_LAUNCH_CHECK(ctx);
    }
    XSrc xco{}; xco.mode = XMODE_PLAIN; xco.x = sc.oc; xco.ldx = dim;
    MDC_TRY(launch_linear<T>(ctx, xco, lw[MDC_CA_OUT_W], (const float*)lw[MDC_CA_OUT_B], sc.y2, dim, B, dim, dim, false, s));
    // FFN
    XSrc x3{}; x3.mode = XMODE_LN; x3.resid = sc.xb; x3.delta = sc.y2; x3.ln_w = (const float*)lw[MDC_LN2_W]; x3.ln_b = (const float*)lw[MDC_LN2_B];
    x3.eps = 1e-5f; x3.xn_out = sc.xc;
    MDC_TRY(launch_linear<T>(ctx, x3, lw[MDC_FF1_W], (const float*)lw[MDC_FF1_B], sc.f1, d.dec_ffn, B, d.dec_ffn, dim, true, s));
    XSrc xf{}; xf.mode = XMODE_PLAIN; xf.x = sc.f1; xf.ldx = d.dec_ffn;
    MDC_TRY(launch_linear<T>(ctx, xf, lw[MDC_FF2_W], (const float*)lw[MDC_FF2_B], sc.y3, dim, B, dim, d.dec_ffn, false, s));
    prev = XSrc{}; prev.mode = XMODE_LN; prev.resid = sc.xc; prev.delta = sc.y3; prev.ln_w = (const float*)lw[MDC_LN3_W];
    prev.ln_b = (const float*)lw[MDC_LN3_B]; prev.eps = 1e-5f;
  }
  {
    const int V = d.vocab, Vp2 = next_pow2(V);
    size_t smem = (size_t)(dim + V + Vp2) * sizeof(float);
    MDC_ENSURE_SMEM(dec_head_select_kernel<T>, smem);
    dec_head_select_kernel<T><<<B, SEL_THREADS, smem, s>>>(prev, (const T*)gw[MDC_OUT_W], (const float*)gw[MDC_OUT_B], V, Vp2, dim, t,
                                                          st->logits, (int64_t)st->logits_ld * V, t + st->logits_row_offset, st->tokens,
                                                          st->tokens_ld, st->forced, st->confs, st->confs_ld, st->uniforms, st->uniforms_ld,
                                                          st->top_k, st->top_p);
    MDC_LAUNCH_CHECK(ctx);
  }
  return 0;
}

}  // namespace

extern "C" size_t mdc_decode_workspace_bytes(const mdc_model* m, int B) {
  if (!m || B <= 0) return 0;
  return align_up(scratch_floats(m->d, B) * sizeof(float), 256);
}

extern "C" size_t mdc_kv_page_bytes(const mdc_model* m) {
  if (!m) return 0;
  return (size_t)m->d.dec_layers * 2 * m->d.page_tokens * m->d.dim * esize(m->d.precision);
}

extern "C" int mdc_decode_steps(mdc_model* m, const mdc_decode_state* st, int t_begin, int t_end, void* stream) {
  MDC_CHECK_ARG(m && st && st->B > 0 && st->tokens && st->kv_pool && st->page_table && st->cross_kv && st->scratch);
  MDC_CHECK_ARG(t_begin >= 0 && t_end >= t_begin);
  MDC_CHECK_ARG(st->x_override || st->pos_override || t_end <= m->d.max_pos);   // Q6: the pos table has max_len-1 rows
  MDC_CHECK_ARG(t_end <= st->pages_per_seq * m->d.page_tokens);
  MDC_CHECK_ARG(st->forced || t_end < st->tokens_ld);
  MDC_CHECK_ARG(st->scratch_bytes >= mdc_decode_workspace_bytes(m, st->B));
  if (st->logits) MDC_CHECK_ARG(t_end + st->logits_row_offset <= st->logits_ld);
  if (st->confs) MDC_CHECK_ARG((t_end + 3) / 4 <= st->confs_ld);
  if ((st->top_k != 0 || st->top_p != 1.0f) && !st->forced) MDC_CHECK_ARG(st->uniforms && t_end <= st->uniforms_ld);
  cudaStream_t s = (cudaStream_t)stream;
  for (int t = t_begin; t < t_end; ++t) {
    if (m->d.precision == MDC_F32) MDC_TRY(decode_step_typed<float>(m, st, t, s));
    else MDC_TRY(decode_step_typed<bf16>(m, st, t, s));
  }
  return 0;
}

extern "C" int mdc_select(mdc_ctx* ctx, const float* logits, int64_t ld, int B, int V, int top_k, float top_p, const float* uniforms,
                          int32_t* token_out, float* conf_out, void* stream) {
  MDC_CHECK_ARG(ctx && logits && B > 0 && V > 0 && V <= 4096 && (token_out || conf_out));
  int Vp2 = next_pow2(V);
  size_t smem = (size_t)(V + Vp2) * sizeof(float);
  select_kernel<<<B, SEL_THREADS, smem, (cudaStream_t)stream>>>(logits, ld, V, Vp2, top_k, top_p, uniforms, token_out, conf_out);
  MDC_LAUNCH_CHECK(ctx); return 0;
}
