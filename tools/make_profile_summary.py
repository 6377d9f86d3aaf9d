"""Turns the raw artefacts of tools/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: python tools/make_profile_summary.py <tag> <round-name>      e.g.  p1 r1a"""
import collections, csv, io, json, os, re, subprocess, sys
tag, rnd = sys.argv[1], sys.argv[2]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
out = [f"# Profile summary {rnd} (raw artefacts: gpurun_out/{tag}_*, produced by tools/profile_round.sh on one B200)\n"]

# ---- launch list of one bench step (ncu --metrics gpu__time_duration.sum --clock-control none) ----
rows = list(csv.DictReader(l for l in open(f"{G}/{tag}_launches.csv") if l.startswith('"')))
starts = [i for i, r in enumerate(rows) if "im2col" in r["Kernel Name"]]
step = rows[starts[-1]:]
with open(f"{P}/{rnd}_launches.csv", "w") as f:
    w = csv.writer(f); w.writerow(["idx", "kernel", "grid", "block", "duration_ns"])
    for i, r in enumerate(step):
        w.writerow([i, re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""), r["Grid Size"], r["Block Size"], r["Metric Value"]])
agg = collections.OrderedDict()
for r in step:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
out.append(f"## Launch list of one bench step (B=64, T=99; last CUDA-graph replay of `bench.py --profile`; {len(step)} launches, {tot/1e3:.2f} ms serialised)\n")
out.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{n[:80]}` | {c} | {t:.1f} | {100*t/tot:.1f}% |")
out.append(f"\nFull list: `profiles/{rnd}_launches.csv`.\n")

# ---- ncu --set full captures ----
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__pcsamp_sample_count"]
traffic = {}
for name in ["gemm_fc1", "attn", "decode", "decode16", "ln", "iou"]:
    rep = f"{G}/{tag}_{name}.ncu-rep"
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rr[0], rr[1], rr[2]
    d = dict(zip(hdr, zip(vals, units)))
    out.append(f"## `{d['Kernel Name'][0][:110]}`  (ncu --set full --clock-control none; `{rep}`)\n")
    def _bytes(key):
        v, u = d.get(key, ("0", "byte"))
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    traffic[name] = {"kernel": d["Kernel Name"][0][:80], "dram_bytes_per_launch": _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum"),
                     "duration_us_under_ncu": float(d["gpu__time_duration.sum"][0]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1)}
    out.append("| metric | value |\n|---|---|")
    for k in WANT:
        if k in d:
            out.append(f"| {k} | {d[k][0]} {d[k][1]} |")
    stalls = sorted(((float(v[0]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)[:6]
    out.append("| top stall reasons (pc samples) | " + ", ".join(f"{n} {int(v)}" for v, n in stalls) + " |")
    out.append("")
    lines = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "14"], capture_output=True, text=True).stdout
    out.append("Hottest source lines (stall samples, instructions executed):\n\n```\n" + lines + "```\n")
for f in ["bench.log", "gemm.log", "cublas.log", "iou.log", "trace.log", "trace16.log", "pipeline.log"]:
    pth = f"{G}/{tag}_{f}"
    if os.path.exists(pth):
        out.append(f"## {f}\n\n```\n" + open(pth).read().strip() + "\n```\n")
open(f"{P}/{rnd}_summary.md", "w").write("\n".join(out))
json.dump(traffic, open(f"{P}/{rnd}_traffic.json", "w"), indent=1)
print("wrote", f"{P}/{rnd}_summary.md", f"{P}/{rnd}_launches.csv")
