"""Feasibility probe: overlap the encoder of batch i+1 (low-priority stream, lands on the SMs the decode clusters leave idle)
with the decode loop of batch i (high-priority stream).  Prints images/s for K batches, serial vs pipelined."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
from mdcnet_b200.model import GenerationPlan
B, T, K = 64, 99, int(sys.argv[1]) if len(sys.argv) > 1 else 12
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
eng = m._engine(torch.device("cuda", 0))
xs = [cases.images(B, seed=100 + i).to("cuda") for i in range(4)]
# serial reference
for _ in range(3): m.generate_tokens(xs[0], T)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
outs = [m.generate_tokens(xs[i % 4], T)[0] for i in range(K)]
b.record(); torch.cuda.synchronize()
print(f"serial   : {B * K / (a.elapsed_time(b) / 1e3):9.1f} img/s  ({a.elapsed_time(b) / K:.3f} ms/batch)")
for depth in (2, 3):
    plans = [GenerationPlan(eng, B, T, 0, 1.0, False, False, True, split=True) for _ in range(depth)]
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    s_enc, s_dec = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
    def run(K):
        res = []
        for i in range(K):
            p = plans[i % depth]
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
    run(4); torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    a.record()
    s_enc.wait_stream(cur); s_dec.wait_stream(cur)
    res = run(K)
    cur.wait_stream(s_dec); cur.wait_stream(s_enc)
    b.record(); torch.cuda.synchronize()
    ok = all(torch.equal(r, o) for r, o in zip(res, outs))
    print(f"pipelined depth {depth}: {B * K / (a.elapsed_time(b) / 1e3):9.1f} img/s  ({a.elapsed_time(b) / K:.3f} ms/batch)  tokens equal to serial: {ok}")
