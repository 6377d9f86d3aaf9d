"""Probe of the batch pipeline's DRAIN, second form: the last L batches of a stream of K decode with fewer images per cluster (more, smaller
clusters: lower latency, more SM-time) on the SMs the finished decode loops have freed.  usage: pipeline_tail_probe2.py K  L,ipc ..."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
from mdcnet_b200.model import GenerationPlan
B, T, K = 64, 99, int(sys.argv[1]) if len(sys.argv) > 1 else 20
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
eng = m._engine(torch.device("cuda", 0))
xs = [cases.images(B, seed=100 + i).to("cuda") for i in range(4)]
for _ in range(3): m.generate_tokens(xs[0], T)
torch.cuda.synchronize()
outs = [m.generate_tokens(xs[i], T)[0] for i in range(4)]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cfgs = [tuple(int(v) for v in c.split(',')) for c in sys.argv[2:]] or [(0, 16), (1, 8), (2, 8), (1, 0)]
depth, ndec = 6, 4
mk = lambda ipc: GenerationPlan(eng, B, T, 0, 1.0, False, False, True, split=True, images_per_cluster=ipc, ctas_per_sm=0)
plans = [mk(16) for _ in range(depth)]
for cfg in cfgs:
    L, ipc = cfg
    tail = [mk(ipc) for _ in range(L)]
    s_enc = torch.cuda.Stream(priority=0)
    s_decs = [torch.cuda.Stream(priority=-1) for _ in range(ndec)]
    def run(K):
        res = []
        for i in range(K):
            p = tail[K - 1 - i] if K - 1 - i < L else plans[i % depth]
            s_dec = s_decs[i % ndec]
            with torch.cuda.stream(s_enc):
                if p.busy: s_enc.wait_event(p.dec_done)
                p.x.copy_(xs[i % 4], non_blocking=True)
                p.enc_graph.replay()
                p.enc_done.record(s_enc)
            with torch.cuda.stream(s_dec):
                s_dec.wait_event(p.enc_done)
                p.dec_graph.replay()
                res.append(p.tokens.clone())
                p.dec_done.record(s_dec)
                p.busy = True
        return res
    run(2 * depth); torch.cuda.synchronize()
    best = 1e9
    for rep in range(4):
        cur = torch.cuda.current_stream()
        a.record()
        s_enc.wait_stream(cur)
        for s in s_decs: s.wait_stream(cur)
        res = run(K)
        for s in s_decs: cur.wait_stream(s)
        cur.wait_stream(s_enc)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    ok = all(torch.equal(r, outs[i % 4]) for i, r in enumerate(res))
    print(f"K {K}: last {L} batches at {ipc} images per cluster: {B * K / (best / 1e3):9.1f} img/s  ({best:.2f} ms total)  equal: {ok}", flush=True)
    del tail
