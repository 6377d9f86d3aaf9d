"""Developer check of the fused decode kernel's softmax rescale path.  The product kernel keeps the maximum of an image's first key tile
as the softmax reference and raises it only when a score exceeds it by 64 log2-units -- with ordinary weights never.  The developer
build reads the margin from MDC_DECODE_REF_MARGIN: margin 0 makes the same code a running maximum (the rescale path on every new
maximum), margin 1 rescales now and then.  All three are exact softmax evaluations that differ only in where the bf16 rounding of the
probabilities falls, so each must meet the bf16 contract against the CPU oracle (teacher-forced on the oracle's own trajectory,
B = 16, 98 positions) with about the same error.   python tools/ref_margin_check.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MDC_LIB_PATH", os.path.join(ROOT, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200", "libmdc_b200_dev.so"))
from oracle import cases, mdc_oracle as O       # checker only
import mdcnet_b200 as M
m = cases.build_product_model("P", seed=0, gamma_seed=5)
sd, cfg = cases.state_dict_of(m), cases.oracle_cfg("P")
x = cases.images(16, seed=404)
want_toks, _, want_logits = O.generate(sd, x, cfg, max_len=98, return_logits=True)
m = m.to("cuda").set_precision("bf16")
errs = {}
for margin in ("64", "1", "0"):
    os.environ["MDC_DECODE_REF_MARGIN"] = margin
    with M.decode_options(prefill=False, images_per_cluster=16):
        got = m.predict(x.to("cuda"), want_toks[:, :98].to("cuda"))[:, 1:99].cpu()
    errs[margin] = (got - want_logits).abs().max().item()
    assert torch.isfinite(got).all()
    print(f"reference-maximum margin {margin:>2s}: logits max|d| vs the oracle = {errs[margin]:.3e}")
assert all(e <= 2e-2 for e in errs.values()) and abs(errs["0"] - errs["64"]) < 3e-3, errs
print("ok")
