// metrics.cu -- detection-metric matching: the per-image greedy assignment of predicted boxes to ground-truth boxes
// that torchmetrics' MeanAveragePrecision (reference: train_val_epoch.py:205-231, iou_thresholds = [0.3]) runs through pycocotools'
// COCOeval.evaluateImg, for a whole batch in one launch.
//
// Per image and class: detections in descending score order (stable), each takes the still-unmatched ground-truth box of its class
// with the highest IoU >= threshold (ties: the later box, as COCOeval's `if ious[d,g] < iou: continue` does); IoU in double
// precision on the float32 coordinates, inter / (area_d + area_g - inter) (pycocotools maskApi bbIou, no crowd boxes).
// One thread per image: the work per image is tiny (N <= 100 detections, M <= ~20 boxes) and sequential by construction.
// Third-party arithmetic absent from this image (torchmetrics, pycocotools): restated from the published algorithm, parity unpinned.
#include "common.cuh"

namespace {

constexpr int MAXN = 128;     // detections per image the kernel orders (COCO evaluates at most 100)

__device__ __forceinline__ double box_iou_d(float4 a, float4 b) {
  const double aw = (double)a.z - (double)a.x, ah = (double)a.w - (double)a.y, bw = (double)b.z - (double)b.x, bh = (double)b.w - (double)b.y;
  const double w = fmin((double)a.z, (double)b.z) - fmax((double)a.x, (double)b.x);
  const double h = fmin((double)a.w, (double)b.w) - fmax((double)a.y, (double)b.y);
  if (w <= 0.0 || h <= 0.0) return 0.0;
  const double inter = w * h, uni = aw * ah + bw * bh - inter;
  return inter / uni;
}

__global__ void map_match_kernel(const float4* __restrict__ pred, const float* __restrict__ scores, const int32_t* __restrict__ labels,
                                 const int32_t* __restrict__ n_pred, const float4* __restrict__ gt, const int32_t* __restrict__ gt_labels,
                                 const int32_t* __restrict__ n_gt, int B, int N, int M, double thr, int max_det,
                                 int32_t* __restrict__ match, int32_t* __restrict__ rank_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int np = min(n_pred[b], N), ng = min(n_gt[b], M);
  const float4* p = pred + (int64_t)b * N; const float* sc = scores + (int64_t)b * N; const int32_t* lb = labels + (int64_t)b * N;
  const float4* g = gt + (int64_t)b * M; const int32_t* gl = gt_labels + (int64_t)b * M;
  int32_t* mt = match + (int64_t)b * N; int32_t* rk = rank_out + (int64_t)b * N;
  for (int i = 0; i < N; ++i) { mt[i] = -2; rk[i] = -1; }          // -2: not evaluated (beyond n_pred / max_det)
  // stable descending order by score (insertion sort on indices; np <= MAXN)
  short order[MAXN];
  const int n = min(np, MAXN);
  for (int i = 0; i < n; ++i) {
    int j = i;
    const float s = sc[i];
    while (j > 0 && sc[order[j - 1]] < s) { order[j] = order[j - 1]; --j; }
    order[j] = (short)i;
  }
  unsigned long long taken_lo = 0ull, taken_hi = 0ull;               // matched ground-truth boxes (M <= 128)
  // COCOeval keeps at most max_det detections PER CLASS per image, in score order
  for (int oi = 0; oi < n; ++oi) {
    const int d = order[oi];
    int cls_rank = 0;
    for (int oj = 0; oj < oi; ++oj) cls_rank += (lb[order[oj]] == lb[d]);
    if (cls_rank >= max_det) continue;
    rk[d] = oi;
    double best = fmin(thr, 1.0 - 1e-10);
    int m = -1;
    for (int j = 0; j < ng; ++j) {
      if (gl[j] != lb[d]) continue;
      const bool taken = j < 64 ? ((taken_lo >> j) & 1ull) : ((taken_hi >> (j - 64)) & 1ull);
      if (taken) continue;
      const double v = box_iou_d(p[d], g[j]);
      if (v < best) continue;
      best = v; m = j;
    }
    mt[d] = m;
    if (m >= 0) { if (m < 64) taken_lo |= 1ull << m; else taken_hi |= 1ull << (m - 64); }
  }
}

}  // namespace

extern "C" int mdc_map_match(mdc_ctx* ctx, const float* pred_boxes, const float* scores, const int32_t* labels, const int32_t* n_pred,
                             const float* gt_boxes, const int32_t* gt_labels, const int32_t* n_gt, int B, int N, int M, float iou_threshold,
                             int max_det, int32_t* match_out, int32_t* order_out, void* stream) {
  MDC_CHECK_ARG(ctx && pred_boxes && scores && labels && n_pred && gt_boxes && gt_labels && n_gt && match_out && order_out);
  MDC_CHECK_DEVICE(ctx);
  MDC_CHECK_ARG(B >= 0 && N >= 0 && N <= MAXN && M >= 0 && M <= 128 && max_det > 0);
  MDC_CHECK_ARG(((uintptr_t)pred_boxes & 15) == 0 && ((uintptr_t)gt_boxes & 15) == 0);
  if (B == 0 || N == 0) return 0;
  map_match_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const float4*)pred_boxes, scores, labels, n_pred, (const float4*)gt_boxes, gt_labels, n_gt,
                                                                   B, N, M, (double)iou_threshold, max_det, match_out, order_out);
  MDC_LAUNCH_CHECK(ctx); return 0;
}
