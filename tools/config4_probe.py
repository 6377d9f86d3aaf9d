"""BASELINE configs[3] shape: 512x512 inputs -> 1024 patches, strips of 1025 tokens, cross-attention over 1024 keys.  Times the
encoder and a short decode at batch B."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, mdc_oracle as O
import mdcnet_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cases.product_cfg(100)
torch.manual_seed(3)
enc = M.Encoder(model_name=cases.VIT, pretrained=False, out_dim=256, img_size=512)
dec = M.Decoder(305, 1024, 256, 8, 6)
m = M.EncoderDecoder(enc, dec).eval().to("cuda").set_precision("bf16")
x = torch.randn(B, 3, 512, 512, device="cuda")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print(f"config 4, B={B}: encoder {timeit(lambda: m.encoder(x)):.2f} ms; encode + {T} greedy tokens {timeit(lambda: m.generate_tokens(x, T)):.2f} ms")
