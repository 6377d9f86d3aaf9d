"""Runs the encoder GEMM shapes alone (for ncu captures and CUDA-event timing)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdcnet_b200 as M
L = M._lib
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
shapes = [("qkv", B * 197, 1536, 512, L.EPI_BIAS), ("fc1", B * 197, 2048, 512, L.EPI_BIAS_GELU),
          ("proj", B * 197, 512, 512, L.EPI_LS_RESIDUAL), ("fc2", B * 197, 512, 2048, L.EPI_LS_RESIDUAL),
          ("patch", B * 196, 512, 768, L.EPI_PATCH), ("crosskv", B * 196, 512, 256, L.EPI_BIAS)]
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, Mr, N, K, epi in shapes:
    if only and name not in only:
        continue
    A = (torch.randn(Mr, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev); gamma = torch.ones(N, device=dev)
    if epi == L.EPI_LS_RESIDUAL:
        D = torch.zeros(Mr, N, device=dev); aux = gamma; period = 0
    elif epi == L.EPI_PATCH:
        D = torch.zeros(B * 197, N, device=dev); aux = torch.zeros(196, N, device=dev); period = 196
    else:
        D = torch.empty(Mr, N, dtype=torch.bfloat16, device=dev); aux = None; period = 0
    def run():
        L.check(L.lib().mdc_gemm(L.ctx(dev), L.MDC_BF16, epi, L.ptr(A), K, L.ptr(W), K, L.ptr(D), N, L.ptr(bias), L.ptr(aux), period, Mr, N, K, L.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    # the launches are replayed from a CUDA graph (as the encoder phase runs them): eager launches are host-bound at ~16 us each
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): run()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / (5 * reps)
    print(f"{name:8s} M={Mr} N={N} K={K}: {ms*1e3:8.1f} us  {2.0*Mr*N*K/ms/1e9:8.1f} TFLOP/s")
