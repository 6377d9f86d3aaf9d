#!/usr/bin/env python
"""bench.py -- images/sec of MDC-Net's batched inference hot path (encode + greedy decode) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one batch of 64 synthetic NEU-DET-shaped images (200x200 gray -> 3x224x224 normalised)
through encoder -> cross-K/V -> 99 greedy decode steps (BASELINE.json configs[1]: config P, B=64, bf16).
N>1 (torchrun): every rank owns its own 64 images (weak scaling; configs[2] at 8 ranks = B 512) and the
step ends with ONE all-gather of the packed results.  Prints ONE JSON line on rank 0.

  value  : K steps through the batch pipeline (mdcnet_b200.GenerationPipeline: the encoder of step i+1 overlaps the decode loops
           of earlier steps on other streams), device-resident inputs, CUDA events around the K steps, max over ranks
  e2e    : the public streaming API generate_stream(model, pinned host batches, tokenizer, max_len): H2D of every batch and D2H
           of its results inside; e2e_gray_u8: the same from raw 200x200 u8 images (fused preprocessing kernel)
  serial : the same numbers one batch at a time (generate_tokens / generate), 256 MiB L2 flush between steps
  roofline: the dominant kernel (fused decode loop) timed alone with CUDA events, algorithmic bytes / time against the measured
           HBM peak, `traffic` = DRAM bytes per launch from the newest committed ncu capture; roofline_gemm: mlp.fc1 against the
           measured bf16 tensor peak
  cpu_baseline: the oracle port of the reference's own loop (encoder recomputed every step, model.py:177-181)
              on the host cores, bounded sample (N = 1 only).   --impl reference prints only that arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 64
T_NEW = 99
METRIC = "images/sec end-to-end encode+greedy decode"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled every 10 ms through NVML in a
    background thread (nvidia-smi -lms needs ~0.5 s to start, longer than a short timed region); nvidia-smi is the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.samples, self.run, self.t, self.h, self.mx = index, [], False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        self.samples = []
        if self.h is not None:
            self.run = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if self.h is None:
            return self._smi_once()
        self.run = False
        self.t.join(timeout=1.0)
        sm = [c for c, _ in self.samples]
        mask = 0
        for _, r in self.samples:
            mask |= r
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": [n for b, n in self.REASONS.items() if mask & b],
                "samples": len(sm), "source": "nvml, 10 ms period"}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
            c = [x.strip() for x in r.stdout.strip().split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(c[0]), "sm_max_mhz": float(c[1]), "reasons": [n for i, n in enumerate(names) if c[2 + i].lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi, one sample right after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def cpu_reference_arm(steps, warmup, sample_steps=6):
    """The reference's own CPU path, restated (oracle port): generate() loop with the encoder recomputed every
    step (Q9) and the full padded decoder (here: its causal-equivalent prefix form), fp32, all host threads.
    Bounded sample: B=1 image, `sample_steps` of the 99 decode steps, extrapolated linearly to 99."""
    from oracle import cases, mdc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = cases.build_product_model("P", seed=0, gamma_seed=5)
    sd, cfg = cases.state_dict_of(model), cases.oracle_cfg("P")
    x = cases.images(1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.generate(sd, x, cfg, max_len=sample_steps, recompute_encoder=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    per_step = statistics.mean(times) / sample_steps
    img_s = 1.0 / (per_step * T_NEW)
    return {"value": img_s, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"B=1, {sample_steps} of {T_NEW} decode steps of the reference loop (encoder recomputed each step), "
                      f"extrapolated x{T_NEW}/{sample_steps}; {per_step * 1e3:.1f} ms/token"}, statistics.mean(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)      # 100 steps ~ 0.6 s pipelined: the fill and drain of the 6-deep pipeline (~8 ms) stay below 2 %
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + 1 step only (for an ncu launch list)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"workload": f"MDC-Net config P (deit3_medium 224 + 6-layer dim-256 decoder, V=305), batch {B_PER_GPU}/GPU, "
                       f"{T_NEW} greedy tokens, synthetic 200x200 gray -> 3x224x224", "global_batch": B_PER_GPU * world,
           "new_tokens": T_NEW, "parallelism": f"dp{world}",
           "pipeline": "batch pipeline (GenerationPipeline / generate_stream): 6 plans, 4 decode streams at 16 images per 8-SM cluster + 1 encoder stream; steps overlap, every step does all of its work inside the timed region",
           "l2": "no flush inside the pipelined region: 4 rotating input batches (154 MB) and a per-step working set of ~330 MB both exceed the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        base, step_s = cpu_reference_arm(max(1, min(args.steps, 3)), min(args.warmup, 1))
        out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
               "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch.distributed as dist
    import mdcnet_b200 as M
    from oracle import cases           # seeded weights / synthetic inputs only (no oracle compute on this arm)
    assert torch.cuda.is_available(), "bench.py needs a B200"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    model = cases.build_product_model("P", seed=0, gamma_seed=5).to(dev).set_precision("bf16")
    tok = M.Tokenizer()
    B = B_PER_GPU
    NROT = 4                                         # rotating input batches: 4 x 38.5 MB > the 126 MB L2
    xs_host = [cases.images(B, seed=1234 + 17 * rank + i).pin_memory() for i in range(NROT)]
    xs_dev = [x.to(dev) for x in xs_host]
    x_host, x_dev = xs_host[0], xs_dev[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T1, C = T_NEW + 1, (T_NEW + 3) // 4

    def step_device():
        toks, confs = model.generate_tokens(x_dev, T_NEW)
        packed = M.parallel.pack_results(toks, confs)
        return M.parallel.all_gather_results(packed, B * world)

    def step_e2e():
        bp, cf = M.generate(model, x_host, tok, max_len=T_NEW)       # H2D of x inside, D2H of tokens+confs inside
        return bp

    if args.profile:
        step_device(); torch.cuda.synchronize()
        n0 = M._lib.launch_count(dev)
        step_device(); torch.cuda.synchronize()
        print(json.dumps({"profile_step_launches": int(M._lib.launch_count(dev) - n0)}))
        return

    pipe = M.GenerationPipeline(model, B, T_NEW)

    def steps_pipelined(k):
        """k steps through the batch pipeline (encoder of step i+1 overlaps the decode loop of step i); every step ends with the
        all-gather of its packed results, stream-ordered behind its decode."""
        outs = []
        for i in range(k):
            t = pipe.submit(xs_dev[i % NROT])
            with torch.cuda.stream(t.stream):
                outs.append(M.parallel.all_gather_results(M.parallel.pack_results(t.tokens, t.confs), B * world))
        pipe.join()
        return outs

    def steps_e2e_pipelined(k):
        """the public streaming API with HOST buffers: H2D of every batch and D2H of its tokens + confs inside"""
        n = 0
        for bp, cf in M.generate_stream(model, (xs_host[i % NROT] for i in range(k)), tok, max_len=T_NEW):
            n += bp.shape[0]
        return n

    # warm-up (also builds the engine / tensor maps / graphs of both the serial plan and the pipeline)
    for _ in range(max(3, args.warmup)):
        out = step_device()
    steps_pipelined(max(8, args.warmup))
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    n0 = M._lib.launch_count(dev)
    for a, b in ev:
        flush.fill_(1)
        a.record()
        out = step_device()
        b.record()
    torch.cuda.synchronize()
    n1 = M._lib.launch_count(dev)
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    serial_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    serial_ms = t.item()
    value_serial = B * world * args.steps / (serial_ms / 1e3)

    # ---- headline: the same K steps through the batch pipeline (device-resident inputs, CUDA events, max over ranks) ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler2 = ClockSampler(local)
    sampler2.start()
    n2 = M._lib.launch_count(dev)
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record()
    out = steps_pipelined(args.steps)
    pb.record()
    torch.cuda.synchronize()
    n3 = M._lib.launch_count(dev)
    if world > 1:
        dist.barrier()
    clocks2 = sampler2.stop()
    clocks_serial = clocks
    if clocks2.get("sm_mhz"):
        clocks = clocks2                      # the headline (pipelined) region
    total_ms = pa.elapsed_time(pb)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    value = B * world * args.steps / (total_ms / 1e3)
    n0, n1 = n2, n3

    # end-to-end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_serial = B * world * args.steps / t.item()
    # end-to-end through the public streaming API (generate_stream), host buffers
    steps_e2e_pipelined(8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps_e2e_pipelined(args.steps)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / t.item()
    # the same with raw 200x200 u8 grayscale host batches (40 KB / image over PCIe, fused preprocessing kernel on the device)
    from oracle import mdc_oracle as _O         # synthetic-image generator only
    gs_host = [_O.synth_gray_u8(B, seed=4321 + 17 * rank + i).pin_memory() for i in range(NROT)]
    def steps_e2e_gray(k):
        for bp, cf in M.generate_stream(model, (gs_host[i % NROT] for i in range(k)), tok, max_len=T_NEW):
            pass
    steps_e2e_gray(8)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps_e2e_gray(args.steps)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_gray = B * world * args.steps / t.item()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (the fused decode loop, ~3/4 of the step), timed alone with CUDA events on this
    # stream; the encoder's largest GEMM against the tensor-pipe peak alongside
    dec_ms = decode_token_ms(M, model, x_dev, dev) * T_NEW
    roof = decode_roofline(dec_ms, peaks, B)
    # the same bytes against the pipelined step (four 16-images-per-cluster decode kernels and an encoder share the GPU): the rate the
    # decode loops sustain together inside the headline region
    roof["in_pipeline"] = {"achieved": roof["algorithmic_bytes_per_launch"] / (total_ms / args.steps / 1e3) / 1e9, "unit": "GB/s",
                           "note": "algorithmic decode bytes of one batch / pipelined ms_per_step (decode kernels overlap each other and the encoder)"}
    roof["in_pipeline"]["frac"] = roof["in_pipeline"]["achieved"] / peaks["hbm_gbs"]
    roof_gemm = roofline_probe(M, dev, peaks, B)
    out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "config": cfg, "clocks": clocks,
           "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": B * (T1 * 4 + C * 4)},
           "gpu_launches": int(n1 - n0), "ms_per_decode_token": dec_ms / T_NEW, "roofline": roof, "roofline_gemm": roof_gemm,
           "e2e_gray_u8": {"value": e2e_gray, "unit": "images/s", "h2d_bytes_per_step": B * 200 * 200, "d2h_bytes_per_step": B * (T1 * 4 + C * 4),
                           "note": "generate_stream over raw u8 200x200 host images; normalisation by mdc_preprocess_gray on the device"},
           "serial": {"value": value_serial, "ms_per_step": serial_ms / args.steps, "e2e": e2e_serial, "clocks": clocks_serial,
                      "note": "one batch at a time (generate_tokens / generate), 256 MiB L2 flush between steps (untimed)"}}
    if not args.no_cpu_baseline and world == 1:          # reported on rank 0 at N = 1 only; `--impl reference` is the arm for every N
        out["cpu_baseline"], _ = cpu_reference_arm(1, 1)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def decode_token_ms(M, model, x_dev, dev):
    eng = model._engine(dev)
    _, memory = eng.encode(x_dev, want_enc_out=False, want_memory=True)
    ckv = eng.cross_kv(memory)
    Bn = x_dev.shape[0]
    tokens = torch.full((Bn, T_NEW + 1), 302, dtype=torch.int32, device=dev); tokens[:, 0] = 300
    kv, scratch = eng.decode(ckv, tokens, 0, T_NEW, max_tokens=T_NEW, forced=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.decode(ckv, tokens, 0, T_NEW, max_tokens=T_NEW, forced=False, kv=kv, scratch=scratch)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / T_NEW


def decode_algorithmic_bytes(B, T=T_NEW, S=196, dim=256, layers=6, ffn=2048, vocab=305):
    """SURVEY 8(d) / DESIGN.md 3.3: bytes one decode launch (T steps) must move, bf16 storage:
    per step  B*(cross_KV + self_KV(t)) + W_step + B*(V*4 + 2*dim*2)."""
    cross = layers * 2 * S * dim * 2                                   # 1 204 224 B / image
    w_step = layers * (3 * dim * dim + 3 * dim * dim + 2 * dim * ffn) * 2 + vocab * dim * 2   # decode-touched weights
    per_step_fixed = B * cross + w_step + B * (vocab * 4 + 2 * dim * 2)
    self_kv = sum(B * layers * 2 * t * dim * 2 for t in range(T))
    return T * per_step_fixed + self_kv


def ncu_traffic(name):
    """DRAM bytes per launch of `name` from the newest committed ncu --set full capture (profiles/*_traffic.json), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    try:
        return json.load(open(files[-1])).get(name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def decode_roofline(dec_ms, peaks, B):
    nbytes = decode_algorithmic_bytes(B)
    ach = nbytes / (dec_ms / 1e3) / 1e9
    return {"kernel": f"decode_fused_kernel (one launch = {T_NEW} decode steps x 6 layers, B={B})", "bound": "hbm", "achieved": ach,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic("decode"),
            "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": dec_ms,
            "peak_src": peaks["src"] + " HBM copy bandwidth (kernel timed alone, CUDA events)"}


def roofline_probe(M, dev, peaks, B):
    """Times the encoder's largest GEMM (mlp.fc1: M=B*197, N=2048, K=512, bias+GELU epilogue) alone."""
    L = M._lib
    Mr, N, K = B * 197, 2048, 512
    A = (torch.randn(Mr, K, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    D = torch.empty(Mr, N, dtype=torch.bfloat16, device=dev)

    def run():
        L.check(L.lib().mdc_gemm(L.ctx(dev), L.MDC_BF16, L.EPI_BIAS_GELU, L.ptr(A), K, L.ptr(W), K, L.ptr(D), N, L.ptr(bias), None, 0,
                                 Mr, N, K, L.stream_ptr()))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    reps = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    flops = 2.0 * Mr * N * K
    ach = flops / (ms / 1e3) / 1e12
    return {"kernel": "gemm_tc_kernel (mlp.fc1 shape, bias+GELU)", "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"],
            "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"], "traffic": ncu_traffic("gemm_fc1"), "peak_src": peaks["src"] + " burst (kernel timed alone)",
            "ms_per_launch": ms}


if __name__ == "__main__":
    main()
