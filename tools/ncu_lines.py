"""Per-source-line stall samples of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] [source_file_for_text]"""
import csv, subprocess, sys, io, collections
def I(v):
    try: return int(v)
    except Exception: return 0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        lines.append((cur_file, int(r[0]), r[1], I(d["# Samples"]), I(d["Instructions Executed"]), d))
tot = sum(l[3] for l in lines)
print(f"total samples {tot}")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for f, ln, src, smp, inst, d in sorted(lines, key=lambda x: -x[3])[:top]:
    st = sorted(((I(d[c]), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{100*smp/tot:5.1f}% {smp:7d} {inst:10d}  {f}:{ln:<4d} {src.strip()[:90]:90s} | " + " ".join(f"{n}={v}" for v, n in st if v))
