"""Where do the fused and the generic decode kernels disagree (teacher-forced)?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 99
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
x = cases.images(B, seed=21).to("cuda")
os.environ["MDC_DECODE_BACKEND"] = "generic"
tg, _ = m.generate_tokens(x, T, use_graph=False)
lg = m.predict(x, tg[:, :T].long())
os.environ.pop("MDC_DECODE_BACKEND")
lc = m.predict(x, tg[:, :T].long())
d = (lg - lc).abs()
per_step = d.amax(dim=(0, 2)).cpu()
print("rows", lg.shape, "per-step max diff (row index = step+1):")
print(" ".join(f"{v:.1e}" for v in per_step.tolist()))
worst = d.flatten().argmax().item()
b, r, v = worst // (d.shape[1] * d.shape[2]), (worst // d.shape[2]) % d.shape[1], worst % d.shape[2]
print("worst at image", b, "row", r, "vocab", v, "generic", lg[b, r, v].item(), "fused", lc[b, r, v].item())
print("tokens of that image:", tg[b].tolist())
print("per-image max diff:", " ".join(f"{v:.1e}" for v in d.amax(dim=(1, 2)).cpu().tolist()))
m.set_precision("fp32")
lf = m.predict(x, tg[:, :T].long())
m.set_precision("bf16")
print("fp32 value there:", lf[b, r, v].item())
print("generic vs fp32 at that row: max", (lg[b, r] - lf[b, r]).abs().max().item(), " fused vs fp32: max", (lc[b, r] - lf[b, r]).abs().max().item())
bad = ((lc - lf).abs() > 0.05).nonzero().cpu().tolist()
print("fused entries off by > 0.05 from fp32:", bad[:20], len(bad))
bad = ((lg - lf).abs() > 0.05).nonzero().cpu().tolist()
print("generic entries off by > 0.05 from fp32:", bad[:20], len(bad))
for rep in range(3):
    lc2 = m.predict(x, tg[:, :T].long())
    print("fused rerun", rep, "max diff vs first fused run", (lc2 - lc).abs().max().item(), "entries > 0.05 off fp32:", ((lc2 - lf).abs() > 0.05).nonzero().cpu().tolist()[:6])
