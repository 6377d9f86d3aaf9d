import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
g = torch.load("tests/golden/case_P_gamma.pt")
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
x = cases.images(2).to("cuda")
os.environ["MDC_DECODE_BACKEND"] = "generic"
tg, _, lg = m.generate_tokens(x, 24, return_logits=True, use_graph=False)
os.environ.pop("MDC_DECODE_BACKEND")
tc, _, lc = m.generate_tokens(x, 24, return_logits=True, use_graph=False)
ref = g["logits"].cuda()
for name, t, l in [("generic", tg, lg), ("cluster", tc, lc)]:
    same = (t.cpu().long() == g["tokens"]).all(dim=0)
    fd = int((~same).nonzero()[0]) if (~same).any() else 25
    n = min(24, fd)
    d = (l[:, :n] - ref[:, :n]).abs()
    print(f"{name}: first token diff vs reference at col {fd}; over {n} common steps max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e}")
d = (lg - lc).abs()
print("generic vs cluster per step max:", [f"{v:.1e}" for v in d.amax(dim=(0, 2)).tolist()])
