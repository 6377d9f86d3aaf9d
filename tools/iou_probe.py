"""Batched IoU kernel alone at a bandwidth-relevant size: B images x N predicted x M ground-truth boxes (config-5 shape scaled up in
B).  Algorithmic bytes = B*(N+M)*16 + B*N*M*4 (+ B*N*4 for the row max)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdcnet_b200 as M
L = M._lib
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
N, Mg = 19, 5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
g = torch.Generator(device="cpu").manual_seed(0)
p = (torch.rand(B, N, 4, generator=g) * 160); p[..., 2:] += p[..., :2] + 8
q = (torch.rand(B, Mg, 4, generator=g) * 160); q[..., 2:] += q[..., :2] + 8
p, q = p.to(dev), q.to(dev)
out = torch.empty(B, N, Mg, device=dev); mx = torch.empty(B, N, device=dev)
def run():
    L.check(L.lib().mdc_iou_batch(L.ctx(dev), L.IOU_EPS, L.ptr(p), L.ptr(q), B, N, Mg, L.ptr(out), L.ptr(mx), L.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
nbytes = B * (N + Mg) * 16 + B * N * Mg * 4 + B * N * 4
print(f"iou_batch B={B} N={N} M={Mg}: {ms * 1e3:.1f} us, {nbytes / 1e6:.1f} MB algorithmic -> {nbytes / ms / 1e6:.0f} GB/s")
