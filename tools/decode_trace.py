"""Developer aid: phase timestamps (clock64) of CTA 0 of the fused decode kernel at one step.  python tools/decode_trace.py [B] [T] [step]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the trace instantiations and the MDC_DECODE_* switches only exist in the developer build (build.py --devtools)
os.environ.setdefault("MDC_LIB_PATH", os.path.join(ROOT, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200", "libmdc_b200_dev.so"))
from oracle import cases
import mdcnet_b200 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 99
step = int(sys.argv[3]) if len(sys.argv) > 3 else 50
m = cases.build_product_model("P", seed=0, gamma_seed=5).to("cuda").set_precision("bf16")
x = cases.images(B, seed=5).to("cuda")
m.generate_tokens(x, T, use_graph=False)
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
os.environ["MDC_DECODE_TRACE_PTR"] = str(trace.data_ptr()); os.environ["MDC_DECODE_TRACE_T"] = str(step)
m.generate_tokens(x, T, use_graph=False)
torch.cuda.synchronize()
tr = trace.cpu().tolist()
names = ["in-proj", "append+own", "self-attn", "merge+gather o", "so proj+push", "LN1(wait y)", "cross q", "cross-attn", "merge+gather o", "co proj+push",
         "LN2(wait y)", "FFN1", "FFN2 issue", "f2 wait+reduce+push", "LN3(wait y)"]
n = len([v for v in tr[:100] if v]); L = (n - 2) // 16
print(f"B={B} T={T} step={step}: {n} stamps, {L} layers")
tot = [0] * 15
for l in range(L):
    s = tr[l * 16:(l + 1) * 16]
    d = [s[i + 1] - s[i] for i in range(15)]
    tot = [a + b for a, b in zip(tot, d)]
    print(f"layer {l}: total {s[15]-s[0]:6d} cyc | " + " ".join(f"{v:5d}" for v in d))
print("mean per layer:")
for nm, v in zip(names, tot):
    print(f"  {nm:22s} {v / L:8.0f} cyc")
print(f"  whole loop: consumer warp 0 blocked on ring stages {tr[200]} cyc, on exchanges {tr[201]} cyc; producer blocked on free slots {tr[202]} cyc")
wn = ["in-proj", "self K/V", "self out-proj", "cross q", "cross K/V", "cross out-proj", "FFN1", "FFN2", "head"]
print("  ring-stage waits of warp 0 by phase (whole loop, cycles): " + ", ".join(f"{n} {tr[210 + i]}" for i, n in enumerate(wn)))
print(f"  layer total {sum(tot)/L:.0f} cyc; head {tr[L*16]-tr[L*16-1]} cyc; select+token exchange {tr[L*16+1]-tr[L*16]} cyc; step {tr[L*16+1]-tr[0]} cyc")

f = tr[100:180]
print("fine stamps, layer 2 (cycles relative to FFN1 stage-0 start of warp 0): per FFN1 stage [enter, data ready, mma done, released]")
for s4 in range(2):
    v = f[s4 * 4:s4 * 4 + 4]
    if v[0]:
        print(f"  FFN1 stage {s4 * 4} (warp 0): enter +{v[0]-f[0]:6d}  wait {v[1]-v[0]:5d}  mma {v[2]-v[1]:5d}  epilogue+release {v[3]-v[2]:5d}")
print("cross stages (warp 0): [enter, data ready, released]")
for g in range(8):
    v = f[20 + g * 3:23 + g * 3]
    if v[0]:
        print(f"  cross image {g}: enter +{v[0]-f[20]:6d}  wait {v[1]-v[0]:5d}  attend+release {v[2]-v[1]:5d}")
for nm, i in (("LN2", 60), ("LN3", 64)):
    v = f[i:i + 4]
    if v[0]:
        print(f"  {nm} (warp 0, layer 2): wait for y {v[1]-v[0]:5d}  normalise own images {v[2]-v[1]:5d}  barrier {v[3]-v[2]:5d}")
import sys; sys.exit(0)
for g in range(8):
    v = tr[100 + g * 8:100 + g * 8 + 6]
    if v[0]:
        print(f"  stage {g}: start +{v[0]-tr[100]:6d} | " + " ".join(f"{v[i+1]-v[i]:6d}" for i in range(5)))
