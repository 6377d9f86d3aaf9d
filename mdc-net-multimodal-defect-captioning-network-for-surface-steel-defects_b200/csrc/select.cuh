// select.cuh -- greedy / top-k / top-p token select + max-prob over one image's logits held in shared memory.
#pragma once
#include "common.cuh"
#include <float.h>

// block barrier used inside the select; a kernel whose select runs on a subset of its warps overrides it with a named barrier
#ifndef MDC_SEL_SYNC
#define MDC_SEL_SYNC() __syncthreads()
#endif

namespace mdcsel {

constexpr int SEL_THREADS = 256;

__device__ __forceinline__ void block_argmax(float v, int idx, float* s_val, int* s_idx, float& out_v, int& out_i) {
  // max value, lowest index on ties (torch.argmax returns the first maximal index)
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o); int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = v; s_idx[warp] = idx; }
  MDC_SEL_SYNC();
  if (warp == 0) {
    v = lane < (SEL_THREADS / 32) ? s_val[lane] : -INFINITY; idx = lane < (SEL_THREADS / 32) ? s_idx[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o); int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { s_val[0] = v; s_idx[0] = idx; }
  }
  MDC_SEL_SYNC();
  out_v = s_val[0]; out_i = s_idx[0];
  MDC_SEL_SYNC();
}

__device__ __forceinline__ double block_sum_d(double v, double* s_d) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_d[warp] = v;
  MDC_SEL_SYNC();
  double tot = 0.0;
  for (int i = 0; i < SEL_THREADS / 32; ++i) tot += s_d[i];
  MDC_SEL_SYNC();
  return tot;
}

// lg: V logits in shared memory (may be overwritten with the filtered logits); srt: scratch of
// next_pow2(V) floats.  Returns (token, conf) in thread 0.
// Semantics: transformers top_k_top_p_filtering (inference_p.py:83) -> conf = max softmax prob of the
// filtered logits (inference_p.py:84-86) -> greedy argmax (inference_p.py:77) or inverse-CDF draw with u.
__device__ inline void select_from_logits(float* lg, float* srt, int V, int Vp2, int top_k, float top_p, bool sample, float u,
                                   int& token, float& conf, float* token_prob = nullptr) {
  __shared__ float s_val[SEL_THREADS / 32]; __shared__ int s_idx[SEL_THREADS / 32];
  __shared__ double s_d[SEL_THREADS / 32]; __shared__ double s_scan[SEL_THREADS]; __shared__ int s_first;
  const int tid = threadIdx.x;
  if (top_k > 0 || top_p < 1.0f) {
    for (int i = tid; i < Vp2; i += SEL_THREADS) srt[i] = i < V ? lg[i] : -INFINITY;
    MDC_SEL_SYNC();
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
        MDC_SEL_SYNC();
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
      MDC_SEL_SYNC();
      int nk = s_first;
      cut = fmaxf(cut, srt[nk - 1]);
      MDC_SEL_SYNC();
    }
    for (int i = tid; i < V; i += SEL_THREADS) if (lg[i] < cut) lg[i] = -INFINITY;
    MDC_SEL_SYNC();
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
    MDC_SEL_SYNC();
    if (tid == 0) { double run = 0.0; for (int i = 0; i < SEL_THREADS; ++i) { double v = s_scan[i]; s_scan[i] = run; run += v; } s_d[0] = run; s_first = V - 1; }
    MDC_SEL_SYNC();
    double total = s_d[0], thr = (double)u * total, run = s_scan[tid];
    int mine = 0x7fffffff;
    for (int j = 0; j < EPT; ++j) {
      int i = tid * EPT + j;
      if (i < V) { run += exp((double)lg[i] - (double)mx); if (run > thr && mine == 0x7fffffff) mine = i; }
    }
    if (mine != 0x7fffffff) atomicMin(&s_first, mine);
    MDC_SEL_SYNC();
    tok = s_first;
    MDC_SEL_SYNC();
  }
  token = tok; conf = cf;
  if (token_prob) *token_prob = expf(lg[tok] - mx) * cf;      // softmax probability of the selected token (filtered distribution)
}


}  // namespace mdcsel
