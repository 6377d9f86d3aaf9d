"""Config T (the geometry the reference was trained with: dim 1024, 8 heads x 128, 8 layers, vocab 332 -- trail_01.py:158-160) through
generate_tokens at B = 64, 99 greedy tokens, bf16: the per-operation decode kernels (the fused cluster kernel is specialised to dim 256)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases
import mdcnet_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 99
cases.product_cfg(100)
torch.manual_seed(4)
enc = M.Encoder(model_name=cases.VIT, pretrained=False, out_dim=1024)
dec = M.Decoder(332, 196, 1024, 8, 8)
model = M.EncoderDecoder(enc, dec).eval().to("cuda").set_precision("bf16")
x = cases.images(B, seed=5).to("cuda")
for _ in range(2): model.generate_tokens(x, T)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): model.generate_tokens(x, T)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"config T, B={B}, {T} greedy tokens, bf16: {ms:.2f} ms per batch = {B / ms * 1e3:.0f} img/s, {ms / T * 1e3:.0f} us per token-step (encoder included)")
# the batch pipeline on the same geometry: the per-operation decode graphs of several batches overlap (each kernel of the chain is
# a few microseconds on at most 64-192 CTAs, so one batch alone leaves most of the GPU idle)
K = 12
for ndec in ((4, 8, 12) if "--pipeline" in sys.argv else ()):
    pipe = M.GenerationPipeline(model, B, T, depth=ndec + 2, decode_streams=ndec)
    for _ in range(ndec + 2): pipe.submit(x)
    pipe.join(); torch.cuda.synchronize()
    a.record()
    ts = [pipe.submit(x) for _ in range(K)]
    pipe.join()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    ref = model.generate_tokens(x, T)[0]
    ok = all(torch.equal(t.tokens, ref) for t in ts)
    print(f"config T, batch pipeline, {ndec} decode streams: {ms:.2f} ms per batch = {B / ms * 1e3:.0f} img/s  (tokens equal to generate_tokens: {ok})")
    del pipe
