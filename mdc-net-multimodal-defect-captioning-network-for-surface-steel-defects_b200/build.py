"""Builds csrc/*.cu into libmdc_b200.so (sm_100a only) next to this file.  nvcc cross-compiles
without a GPU, so this runs in the CPU-only build container; the .so travels to the GPU box."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmdc_b200.so")
# developer build (-DMDC_DEVTOOLS): the same sources plus the decode kernel's phase-trace instantiations and the environment
# switches tools/*.py use for A/B runs (MDC_DECODE_TRACE_PTR, MDC_DECODE_IPC, MDC_DECODE_CPS, MDC_GEMM_BACKEND, MDC_ATTN_BACKEND).
# Never loaded by the package unless MDC_LIB_PATH points at it.
OBJ_DEV = os.path.join(HERE, "build_dev")
LIB_DEV = os.path.join(HERE, "libmdc_b200_dev.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "mdc_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False, devtools=False):
    global OBJ, LIB
    if devtools:
        saved = (OBJ, LIB)
        OBJ, LIB = OBJ_DEV, LIB_DEV
        try:
            return _build(force, verbose, ["-DMDC_DEVTOOLS"])
        finally:
            OBJ, LIB = saved
    return _build(force, verbose, [])


def _build(force, verbose, extra):
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        r = subprocess.run([NVCC, *FLAGS, *extra, "-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {os.path.basename(s)}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    objs = [os.path.join(OBJ, src[:-3] + ".o") for src in _sources()]
    if jobs or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-Xcompiler", "-fPIC", "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, devtools="--devtools" in sys.argv))
