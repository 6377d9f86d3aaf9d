"""One generate() call at the bench shape (for ncu captures of the decode kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 99
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
x = cases.images(B, seed=5).to("cuda")
for _ in range(2):
    toks, confs = m.generate_tokens(x, T, use_graph=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); toks, confs = m.generate_tokens(x, T, use_graph=False); b.record(); torch.cuda.synchronize()
print(f"B={B} T={T}: {a.elapsed_time(b):.2f} ms total (eager launches)")
