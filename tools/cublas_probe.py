import torch
shapes = {"qkv": (12608, 1536, 512), "fc1": (12608, 2048, 512), "proj": (12608, 512, 512), "fc2": (12608, 512, 2048)}
for nm, (M, N, K) in shapes.items():
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16); w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    for _ in range(5): torch.nn.functional.linear(a, w, b)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.nn.functional.linear(a, w, b); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); t = ts[len(ts) // 2]
    print(f"cuBLAS {nm:5s} M={M} N={N} K={K}: {t:8.1f} us {2 * M * N * K / t / 1e6:9.1f} TFLOP/s")
