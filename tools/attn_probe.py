"""Encoder strip attention (197-token strips, 8 heads x 64) and LayerNorm alone, CUDA-event timing (for same-box A/B runs)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdcnet_b200 as M
L = M._lib
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S, H, HD = 197, 8, 64
qkv = (torch.randn(B * S, 3 * H * HD, device=dev) * 0.7).to(torch.bfloat16)
out = torch.empty(B * S, H * HD, dtype=torch.bfloat16, device=dev)
x = torch.randn(B * S, 512, device=dev); w = torch.ones(512, device=dev); b = torch.zeros(512, device=dev)
u = torch.empty(B * S, 512, dtype=torch.bfloat16, device=dev)
def attn():
    L.check(L.lib().mdc_strip_attention(L.ctx(dev), L.MDC_BF16, L.ptr(qkv), 3 * H * HD, L.ptr(out), H * HD, B, S, H, HD, 0.125, 0, L.stream_ptr()))
def ln():
    L.check(L.lib().mdc_layernorm(L.ctx(dev), L.ptr(x), 512, L.ptr(w), L.ptr(b), 1e-6, L.ptr(u), 512, L.MDC_BF16, B * S, 512, L.stream_ptr()))
for name, fn in (("strip attention", attn), ("layernorm", ln)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    print(f"{name:16s} B={B}: {a.elapsed_time(e) / reps * 1e3:8.1f} us")
