"""Decode-kernel variants timed alone (CUDA events, no encoder): python tools/decode_variants.py [T] B,ipc,cps ...
ipc = images per cluster, cps = CTAs per SM (mdc_decode_state.images_per_cluster / ctas_per_sm).  Prints ms per launch,
images/s of the decode loop alone and SM-time per image; checks every variant's tokens against the first one (bitwise)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
T = int(sys.argv[1]) if len(sys.argv) > 1 else 99
cfgs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[2:]] or [(64, 0, 0), (64, 8, 0), (64, 16, 0), (64, 8, 2)]
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
dev = torch.device("cuda", 0)
eng = m._engine(dev)
ref = {}
for B, ipc, cps in cfgs:
    x = cases.images(min(B, 64), seed=5).to(dev)
    if B > 64:
        x = x.repeat((B + 63) // 64, 1, 1, 1)[:B].contiguous()
    _, memory = eng.encode(x, want_enc_out=False, want_memory=True)
    ckv = eng.cross_kv(memory)
    tokens = torch.full((B, T + 1), 302, dtype=torch.int32, device=dev); tokens[:, 0] = 300
    kv, scratch = eng.decode(ckv, tokens, 0, T, max_tokens=T, forced=False, images_per_cluster=ipc, ctas_per_sm=cps)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.decode(ckv, tokens, 0, T, max_tokens=T, forced=False, kv=kv, scratch=scratch, images_per_cluster=ipc, ctas_per_sm=cps)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    same = ""
    if B in ref:
        same = f"  tokens equal to first variant: {torch.equal(ref[B], tokens)}"
    else:
        ref[B] = tokens.clone()
    print(f"B={B:4d} ipc={ipc:2d} cps={cps}: {ms:8.3f} ms/launch  {B / ms * 1e3:9.1f} img/s (decode only)  {ms / T * 1e3:7.1f} us/token{same}", flush=True)
