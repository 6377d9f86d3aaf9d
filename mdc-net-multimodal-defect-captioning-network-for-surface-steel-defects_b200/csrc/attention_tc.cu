// attention_tc.cu -- tensor-core strip attention for bf16 (placeholder until the mma path lands):
// reports "unsupported" so mdc_strip_attention uses the fp32-math strip kernel on bf16 storage.
#include "common.cuh"
int attn_tc_supported(int strip_len, int head_dim) { (void)strip_len; (void)head_dim; return 0; }
int attn_tc_launch(mdc_ctx*, const void*, int64_t, void*, int64_t, int, int, int, int, float, cudaStream_t) {
  mdc_set_error("attn_tc_launch: not built"); return -5;
}
