// iou.cu -- batched box scoring (north_star (e)): one launch for the whole batch instead of the
// reference's per-image Python loop of ~12 broadcast launches + a .tolist() sync
// (iou_calcualtions.py:45-105, iou_bbox.py:3-63).
//
// One thread = one predicted box: a single 128-bit load of its xyxy, then the image's GT boxes
// (128-bit loads, staged per block in shared memory), M contiguous outputs and the row-max.
// Arithmetic is written with non-contracting intrinsics in the reference's operation order so the
// result is bit-identical to the torch expression (no FMA fusion of w*h into the union).
#include "common.cuh"
#include <float.h>

namespace {

struct Pair { float inter, uni, enc; };

__device__ __forceinline__ Pair pair_terms(float4 p, float4 g) {
  float ap = __fmul_rn(__fsub_rn(p.z, p.x), __fsub_rn(p.w, p.y));
  float ag = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
  float iw = fmaxf(__fsub_rn(fminf(p.z, g.z), fmaxf(p.x, g.x)), 0.f);
  float ih = fmaxf(__fsub_rn(fminf(p.w, g.w), fmaxf(p.y, g.y)), 0.f);
  Pair r;
  r.inter = __fmul_rn(iw, ih);
  r.uni = __fsub_rn(__fadd_rn(ap, ag), r.inter);
  r.enc = __fmul_rn(__fsub_rn(fmaxf(p.z, g.z), fminf(p.x, g.x)), __fsub_rn(fmaxf(p.w, g.w), fminf(p.y, g.y)));
  return r;
}

template <int mode>
__device__ __forceinline__ float score(float4 p, float4 g) {
  Pair t = pair_terms(p, g);
  if (mode == MDC_IOU_EPS) return __fdiv_rn(t.inter, __fadd_rn(t.uni, 1e-6f));
  float iou = __fdiv_rn(t.inter, t.uni);
  if (mode == MDC_IOU_PLAIN) return iou;
  if (mode == MDC_IOU_NAN0) {          // torch.nan_to_num(nan=0.0): nan->0, +-inf -> +-FLT_MAX
    if (isnan(iou)) return 0.f;
    if (isinf(iou)) return iou > 0 ? FLT_MAX : -FLT_MAX;
    return iou;
  }
  return __fsub_rn(iou, __fdiv_rn(__fsub_rn(t.enc, t.uni), t.enc));   // GIoU
}

constexpr int IOU_THREADS = 256;

// I = index type: 32-bit when B*N*M < 2^31 (the kernel is issue-bound, and three 64-bit integer divisions per thread were a third
// of its instructions), 64-bit otherwise.  STAGED: the GT boxes of the block's images and the block's 256 x M outputs go through
// shared memory (coalesced write-back); shapes whose staging does not fit (many GT boxes per image: bbox_iou((N,4),(M,4)) with
// M in the hundreds, or N = 1 with dozens of GT boxes) take the direct form: GT boxes by 128-bit read-only loads, outputs stored
// straight to global memory.
template <int MODE, typename I, bool STAGED>
__global__ void __launch_bounds__(IOU_THREADS) iou_batch_kernel(const float4* __restrict__ pred,
                                                                const float4* __restrict__ gt, int B, int N, int M,
                                                                float* __restrict__ iou_out, float* __restrict__ max_out, int n_gt_cap) {
  extern __shared__ float4 sgt[];   // GT boxes of the images this block touches, then (when iou_out) the block's 256 x M outputs
  float* sout = reinterpret_cast<float*>(sgt + n_gt_cap);
  const I total = (I)B * N;
  const I first = (I)blockIdx.x * IOU_THREADS;
  const int img0 = (int)(first / N);
  if (STAGED) {
    const I last = min(first + (I)IOU_THREADS, total) - 1;
    const int img1 = (int)(last / N);
    const int n_gt = (img1 - img0 + 1) * M;
    for (int i = threadIdx.x; i < n_gt; i += IOU_THREADS) sgt[i] = __ldg(gt + (I)img0 * M + i);
    __syncthreads();
  }
  I idx = first + threadIdx.x;
  if (idx < total) {
    const int img = (int)(idx / N);
    const float4 p = __ldg(pred + idx);
    const float4* g = STAGED ? sgt + (img - img0) * M : gt + (I)img * M;
    float best = -INFINITY;
    bool any_nan = false;
#pragma unroll 4
    for (int j = 0; j < M; ++j) {
      const float v = score<MODE>(p, STAGED ? g[j] : __ldg(g + j));
      if (iou_out) { if (STAGED) sout[threadIdx.x * M + j] = v; else iou_out[idx * M + j] = v; }
      any_nan |= isnan(v);
      best = fmaxf(best, v);
    }
    if (max_out) max_out[idx] = any_nan ? NAN : best;   // torch.max propagates NaN
  }
  if (STAGED && iou_out) {
    // the block's outputs are one contiguous range of the (B,N,M) tensor: written back with consecutive lanes on consecutive floats
    // (M-strided 4-byte stores cost 17 L2 sectors per warp store)
    __syncthreads();
    const int n_out = (int)(min(first + (I)IOU_THREADS, total) - first) * M;
    float* dst = iou_out + first * M;
    for (int i = threadIdx.x; i < n_out; i += IOU_THREADS) dst[i] = sout[i];
  }
}

// giou_loss_with_scores: one warp per image
__global__ void giou_loss_kernel(const float4* __restrict__ pred, const float4* __restrict__ gt, int B, int N, int M,
                                 float penalty, float* __restrict__ loss, float* __restrict__ giou_out,
                                 uint8_t* __restrict__ valid_out) {
  int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (img >= B) return;
  const float4* p = pred + (int64_t)img * N;
  const float4* g = gt + (int64_t)img * M;
  int np = 0, ng = 0;
  for (int i = lane; i < N; i += 32) { float4 b = p[i]; np += (((b.x + b.y) + b.z) + b.w) != 0.f; }
  for (int j = lane; j < M; j += 32) { float4 b = g[j]; ng += (((b.x + b.y) + b.z) + b.w) != 0.f; }
  np = __reduce_add_sync(0xffffffffu, np); ng = __reduce_add_sync(0xffffffffu, ng);
  float sum = 0.f;
  for (int e = lane; e < N * M; e += 32) {
    int i = e / M, j = e % M;
    float4 a = p[i], b = g[j];
    bool ok = ((((a.x + a.y) + a.z) + a.w) != 0.f) && ((((b.x + b.y) + b.z) + b.w) != 0.f);
    float v = ok ? score<MDC_IOU_GIOU>(a, b) : 0.f;
    if (giou_out) giou_out[(int64_t)img * N * M + e] = v;
    if (valid_out) valid_out[(int64_t)img * N * M + e] = ok ? 1 : 0;
    sum += v;
  }
  sum = warp_sum(sum);
  if (lane == 0) {
    float l;
    if (np == 0 && ng > 0) l = penalty * (float)ng;
    else if (np == 0 || ng == 0) l = 0.f;
    else l = 1.0f - sum / (float)(np * ng);
    loss[img] = l;
  }
}

__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += v[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = s / (float)n;
}

}  // namespace

extern "C" int mdc_iou_batch(mdc_ctx* ctx, int mode, const float* pred, const float* gt, int B, int N, int M,
                             float* iou_out, float* max_out, void* stream) {
  MDC_CHECK_ARG(ctx && pred && gt && (iou_out || max_out));
  MDC_CHECK_DEVICE(ctx);
  MDC_CHECK_ARG(mode >= MDC_IOU_EPS && mode <= MDC_IOU_GIOU);
  MDC_CHECK_ARG(B >= 0 && N >= 0 && M >= 0);
  MDC_CHECK_ARG(((uintptr_t)pred & 15) == 0 && ((uintptr_t)gt & 15) == 0);
  if (B == 0 || N == 0) return 0;
  MDC_CHECK_ARG(M > 0);
  int64_t total = (int64_t)B * N;
  int grid = (int)((total + IOU_THREADS - 1) / IOU_THREADS);
  int imgs_per_block = IOU_THREADS / (N > 0 ? N : 1) + 2;       // images a 256-thread block can touch
  if (imgs_per_block > B) imgs_per_block = B;
  const int n_gt_cap = imgs_per_block * M;
  size_t smem = (size_t)n_gt_cap * sizeof(float4) + (iou_out ? (size_t)IOU_THREADS * M * sizeof(float) : 0);
  const bool staged = smem <= 96 * 1024;                         // two blocks per SM keep their staging resident
  if (!staged) smem = 0;
  const bool small = total * M < ((int64_t)1 << 31) - IOU_THREADS * (int64_t)M;
#define MDC_IOU_LAUNCH2(MODE_, I_, ST_)                                                                                                              \
  {                                                                                                                                                  \
    MDC_ENSURE_SMEM((iou_batch_kernel<MODE_, I_, ST_>), smem);                                                                                       \
    iou_batch_kernel<MODE_, I_, ST_><<<grid, IOU_THREADS, smem, (cudaStream_t)stream>>>((const float4*)pred, (const float4*)gt, B, N, M, iou_out, max_out, n_gt_cap); \
  }
#define MDC_IOU_LAUNCH(MODE_)                                                                                                                        \
  if (small) { if (staged) MDC_IOU_LAUNCH2(MODE_, int, true) else MDC_IOU_LAUNCH2(MODE_, int, false) }                                               \
  else { if (staged) MDC_IOU_LAUNCH2(MODE_, int64_t, true) else MDC_IOU_LAUNCH2(MODE_, int64_t, false) }
  switch (mode) {
    case MDC_IOU_EPS: MDC_IOU_LAUNCH(MDC_IOU_EPS) break;
    case MDC_IOU_PLAIN: MDC_IOU_LAUNCH(MDC_IOU_PLAIN) break;
    case MDC_IOU_NAN0: MDC_IOU_LAUNCH(MDC_IOU_NAN0) break;
    default: MDC_IOU_LAUNCH(MDC_IOU_GIOU) break;
  }
#undef MDC_IOU_LAUNCH
#undef MDC_IOU_LAUNCH2
  MDC_LAUNCH_CHECK(ctx); return 0;
}

extern "C" int mdc_giou_loss(mdc_ctx* ctx, const float* pred, const float* gt, int B, int N, int M, float no_detection_penalty,
                             float* loss_per_image, float* giou_out, uint8_t* valid_out, void* stream) {
  MDC_CHECK_ARG(ctx && pred && gt && loss_per_image && B > 0 && N >= 0 && M >= 0);
  MDC_CHECK_DEVICE(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  giou_loss_kernel<<<(B + 7) / 8, 256, 0, s>>>((const float4*)pred, (const float4*)gt, B, N, M, no_detection_penalty,
                                               loss_per_image, giou_out, valid_out);
  MDC_LAUNCH_CHECK(ctx);
  mean_kernel<<<1, 32, 0, s>>>(loss_per_image, B, loss_per_image + B);   // slot B receives the batch mean
  MDC_LAUNCH_CHECK(ctx); return 0;
}
