"""Phase stamps (clock64) of CTA 0, first item, of the tcgen05 attention kernel (developer build)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MDC_LIB_PATH", os.path.join(ROOT, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200", "libmdc_b200_dev.so"))
import mdcnet_b200 as M
from tests import gpu_util as G
n_strips, n, H, hd = 64, 197, 8, 64
qkv = torch.randn(n_strips * n, 3 * H * hd).to(torch.bfloat16).cuda()
G.strip_attention(qkv, n_strips, n, H, hd, 0.125); torch.cuda.synchronize()
tr = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["MDC_ATTN_TRACE_PTR"] = str(tr.data_ptr())
G.strip_attention(qkv, n_strips, n, H, hd, 0.125); torch.cuda.synchronize()
t = tr.cpu().tolist(); t0 = t[0]
names = {0: "TMA issue", 1: "MMA: smem full", 2: "QK t0 issued", 3: "QK t1 issued", 4: "MMA: P t0 ready", 5: "MMA: P t1 ready", 6: "PV t0 issued", 7: "PV t1 issued",
         10: "SM t0: S ready", 11: "SM t0: S in regs", 12: "SM t0: max exchanged", 13: "SM t0: P stores issued", 14: "SM t0: P stores done",
         18: "SM t1: S ready", 19: "SM t1: S in regs", 20: "SM t1: max exchanged", 21: "SM t1: P stores issued", 22: "SM t1: P stores done",
         26: "EPI t0: O ready", 28: "EPI t1: O ready", 30: "item done"}
for i in sorted(names, key=lambda k: t[k]):
    if t[i]: print(f"{t[i] - t0:8d}  {names[i]}")
