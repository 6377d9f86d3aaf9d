// tokens.cu -- token sequences -> labels / boxes / caption ids on the GPU: the per-sequence Python scans with
// .item() synchronisations that sit between the decode loop and the IoU kernel in the reference.
//   mode MDC_TOK_BBOXES : Tokenizer.decode_bboxes              data_processing.py:556-598 (+ adjust_bboxes_dimensions :547-554)
//   mode MDC_TOK_DECODE : Tokenizer.decode, batched            data_processing.py:317-391
// One warp per sequence: the warp stages the tokens in shared memory (coalesced), lane 0 walks the grammar (it is a sequential
// automaton over <= a few hundred tokens), all lanes zero-fill the padding rows.  Integer work is exact; the de-quantisation
// float32(v) / (num_bins-1) * extent is evaluated with the reference's own two float32 roundings.
#include "common.cuh"

namespace {

constexpr int TOK_WARPS = 4;

__device__ __forceinline__ float dequant(int v, int num_bins, int extent) {
  return __fmul_rn(__fdiv_rn((float)v, (float)(num_bins - 1)), (float)extent);
}

__global__ void __launch_bounds__(TOK_WARPS * 32) decode_tokens_kernel(int mode, const int32_t* __restrict__ tokens, int64_t ld, int B, int L,
                                                                       mdc_token_grammar gr, int max_boxes, int32_t* __restrict__ labels_out,
                                                                       float* __restrict__ boxes_out, int32_t* __restrict__ counts_out,
                                                                       int32_t* __restrict__ caption_out, int32_t* __restrict__ caption_len_out) {
  extern __shared__ int32_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * TOK_WARPS + warp;
  if (b >= B) return;
  int32_t* seq = sm + warp * L;
  for (int i = lane; i < L; i += 32) seq[i] = tokens[(int64_t)b * ld + i];
  __syncwarp();
  int count = 0, cap_len = 0;
  if (lane == 0) {
    float* bx = boxes_out + (int64_t)b * max_boxes * 4;
    int32_t* lb = labels_out ? labels_out + (int64_t)b * max_boxes : nullptr;
    if (mode == MDC_TOK_BBOXES) {
      int start = 0;
      for (int i = 0; i < L; ++i) if (seq[i] == gr.caption_end) { start = i + 1; break; }
      int i = start;
      while (i < L - 4) {
        const int tok = seq[i];
        if (tok >= gr.label_lo && tok <= gr.label_hi) {
          const int x0 = seq[i + 1], y0 = seq[i + 2], x1 = seq[i + 3], y1 = seq[i + 4];
          const bool in = x0 >= 0 && x0 <= gr.coord_max && y0 >= 0 && y0 <= gr.coord_max && x1 >= 0 && x1 <= gr.coord_max && y1 >= 0 && y1 <= gr.coord_max;
          if (in && x1 > x0 && y1 > y0 && count < max_boxes) {
            bx[count * 4 + 0] = dequant(x0, gr.num_bins, gr.width); bx[count * 4 + 1] = dequant(y0, gr.num_bins, gr.height);
            bx[count * 4 + 2] = dequant(x1, gr.num_bins, gr.width); bx[count * 4 + 3] = dequant(y1, gr.num_bins, gr.height);
            if (lb) lb[count] = tok;
            ++count;
          }
          i += 5;
        } else if (tok == gr.eos) break;
        else i += 1;
      }
    } else {
      // drop PADs, cut at the first EOS (compaction in place)
      int n = 0;
      for (int i = 0; i < L; ++i) { const int tok = seq[i]; if (tok != gr.pad) seq[n++] = tok; }
      for (int i = 0; i < n; ++i) if (seq[i] == gr.eos) { n = i; break; }
      int soc = -1, eoc = -1;
      for (int i = 0; i < n; ++i) { if (soc < 0 && seq[i] == gr.caption_start) soc = i; if (eoc < 0 && seq[i] == gr.caption_end) eoc = i; }
      cap_len = -1;                                   // no caption markers: the reference returns "" instead of a word list
      if (soc >= 0 && eoc >= 0) {
        cap_len = 0;
        int32_t* cp = caption_out ? caption_out + (int64_t)b * L : nullptr;
        for (int i = soc + 1; i < eoc; ++i) { if (cp) cp[cap_len] = seq[i]; ++cap_len; }
        const int base = eoc + 1, m = n - base;
        for (int i = 0; i + 4 < m; i += 5) {
          const int tok = seq[base + i];
          const int x0 = seq[base + i + 1], y0 = seq[base + i + 2], x1 = seq[base + i + 3], y1 = seq[base + i + 4];
          const bool in = x0 >= 0 && x0 <= gr.coord_max && y0 >= 0 && y0 <= gr.coord_max && x1 >= 0 && x1 <= gr.coord_max && y1 >= 0 && y1 <= gr.coord_max;
          if (tok >= gr.label_lo && tok <= gr.label_hi && in && count < max_boxes) {
            bx[count * 4 + 0] = dequant(x0, gr.num_bins, gr.width); bx[count * 4 + 1] = dequant(y0, gr.num_bins, gr.height);
            bx[count * 4 + 2] = dequant(x1, gr.num_bins, gr.width); bx[count * 4 + 3] = dequant(y1, gr.num_bins, gr.height);
            if (lb) lb[count] = tok;
            ++count;
          }
        }
      }
    }
    counts_out[b] = count;
    if (caption_len_out) caption_len_out[b] = cap_len;
  }
  count = __shfl_sync(0xffffffffu, count, 0);
  cap_len = __shfl_sync(0xffffffffu, cap_len, 0);
  // zero rows = padding (pad_sequence(padding_value=0) / the (1,4) zero box of an empty sequence)
  for (int i = count * 4 + lane; i < max_boxes * 4; i += 32) boxes_out[(int64_t)b * max_boxes * 4 + i] = 0.f;
  if (labels_out) for (int i = count + lane; i < max_boxes; i += 32) labels_out[(int64_t)b * max_boxes + i] = 0;
  if (caption_out) for (int i = max(cap_len, 0) + lane; i < L; i += 32) caption_out[(int64_t)b * L + i] = gr.pad;
}

}  // namespace

extern "C" int mdc_decode_tokens(mdc_ctx* ctx, int mode, const int32_t* tokens, int64_t tokens_ld, int B, int L, const mdc_token_grammar* grammar,
                                 int max_boxes, int32_t* labels_out, float* boxes_out, int32_t* counts_out, int32_t* caption_out,
                                 int32_t* caption_len_out, void* stream) {
  MDC_CHECK_ARG(ctx && tokens && grammar && boxes_out && counts_out && B > 0 && L > 0 && L <= 4096 && tokens_ld >= L && max_boxes > 0);
  MDC_CHECK_DEVICE(ctx);
  MDC_CHECK_ARG(mode == MDC_TOK_BBOXES || mode == MDC_TOK_DECODE);
  MDC_CHECK_ARG(grammar->num_bins > 1);
  const size_t smem = (size_t)TOK_WARPS * L * sizeof(int32_t);
  MDC_ENSURE_SMEM(decode_tokens_kernel, smem);
  decode_tokens_kernel<<<(B + TOK_WARPS - 1) / TOK_WARPS, TOK_WARPS * 32, smem, (cudaStream_t)stream>>>(
      mode, tokens, tokens_ld, B, L, *grammar, max_boxes, labels_out, boxes_out, counts_out, caption_out, caption_len_out);
  MDC_LAUNCH_CHECK(ctx);
  return 0;
}
