"""SASS evidence for profiles/: per hot kernel, the instruction-mnemonic histogram of the shipped cubin (cuobjdump -sass of
libmdc_b200.so) with the Blackwell-specific ones called out (UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG/UTMAREDG = TMA load / store /
reduce, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, SYNCS = mbarrier, HMMA = mma.sync, LDSM = ldmatrix, FFMA2/FMUL2 = packed fp32,
UBLKCP = bulk copy), and a short excerpt around the first tensor-core instruction.   usage: python tools/sass_evidence.py <round>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200", "libmdc_b200.so")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
KERNELS = [("gemm_tc_kernelILi256ELi1E", "gemm_tc_kernel<256, BIAS_GELU> (mlp.fc1)"), ("gemm_tc_kernelILi128ELi3E", "gemm_tc_kernel<128, LS_RESIDUAL> (attn.proj / mlp.fc2)"),
           ("decode_fused_kernelILb0E", "decode_fused_kernel<false>"), ("attn_umma_kernel", "attn_umma_kernel (ViT strip attention, tcgen05)"), ("prefill_attn_kernel", "prefill_attn_kernel"),
           ("layernorm_rows_kernelI13__nv_bfloat16Li4E", "layernorm_rows_kernel<bf16,4>"), ("iou_batch_kernel", "iou_batch_kernel"),
           ("decode_tokens_kernel", "decode_tokens_kernel")]
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "LDSM", "FFMA2", "FMUL2",
       "FADD2", "MUFU", "LDGSTS", "LDG", "STG", "RED", "ST.ASYNC", "STAS", "MAPA", "UCGABAR", "CCTL", "FENCE", "ACQBULK", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
out = [f"# SASS evidence {rnd}: cuobjdump -sass of libmdc_b200.so (sm_100a), mnemonic counts per kernel\n"]
for pat, title in KERNELS:
    body = next((f for f in funcs if f.split("\n", 1)[0].find(pat) >= 0), None)
    if body is None:
        out.append(f"## {title}: not found\n"); continue
    ins = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    hist = collections.Counter(i.split(".")[0] if not i.startswith("ST.ASYNC") else "ST.ASYNC" for i in ins)
    full = collections.Counter(ins)
    out.append(f"## {title}\n\n`{body.split(chr(10), 1)[0][:150]}`: {len(ins)} instructions\n")
    out.append("| mnemonic | count | variants |\n|---|---:|---|")
    for k in KEY:
        n = sum(v for i, v in hist.items() if i == k or i.startswith(k))
        if n:
            var = ", ".join(f"{i} x{v}" for i, v in sorted(full.items(), key=lambda x: -x[1]) if i.startswith(k))[:160]
            out.append(f"| {k} | {n} | {var} |")
    lines = body.split("\n")
    first = next((i for i, l in enumerate(lines) if re.search(r"UTCHMMA|HMMA|UTMALDG", l)), None)
    if first is not None:
        out.append("\nexcerpt around the first tensor / TMA instruction:\n\n```")
        out += [l.rstrip()[:150] for l in lines[max(0, first - 4):first + 10] if "/*" in l]
        out.append("```\n")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{rnd}_sass.md"), "w").write("\n".join(out))
print("wrote", f"profiles/{rnd}_sass.md")
