"""Correctness + timing of the tcgen05 strip attention against a float64 reference: python tools/attn_umma_check.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdcnet_b200 as M
from tests import gpu_util as G
torch.manual_seed(0)
for n_strips, n, H in ((2, 197, 8), (3, 64, 2), (5, 256, 8), (4, 129, 8), (64, 197, 8)):
    hd, scale = 64, 0.125
    qkv = (torch.randn(n_strips * n, 3 * H * hd) * 1.0).to(torch.bfloat16).cuda()
    got = G.strip_attention(qkv, n_strips, n, H, hd, scale)
    torch.cuda.synchronize()
    q, k, v = [t.reshape(n_strips, n, H, hd).transpose(1, 2).double() for t in qkv.chunk(3, dim=-1)]
    want = (torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v).transpose(1, 2).reshape(n_strips * n, H * hd)
    err = (got.double() - want).abs().max().item()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): G.strip_attention(qkv, n_strips, n, H, hd, scale)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()                      # 20 launches in one graph: device time, no host launch overhead in between
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(20): G.strip_attention(qkv, n_strips, n, H, hd, scale)
    g.replay(); torch.cuda.synchronize()
    a.record()
    g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"strips {n_strips} x {n} tokens x {H} heads: max|err| = {err:.3e}   {a.elapsed_time(b) / 20 * 1e3:.1f} us/launch", flush=True)
