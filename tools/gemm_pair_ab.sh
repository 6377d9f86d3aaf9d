#!/bin/bash
# Developer A/B (GPU box): CTA-pair (cta_group::2) GEMM against the single-CTA form, CUDA-graph timing, with the MDC_GEMM_DBG splits
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or gelu" 2>&1 | tail -3
export MDC_LIB_PATH=$PWD/mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/libmdc_b200_dev.so
echo "== pairs (default)"; timeout 120 python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2,crosskv
echo "== single CTA"; MDC_GEMM_2CTA=0 timeout 120 python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2,crosskv
for d in 1 12 16; do echo "== pairs dbg $d"; MDC_GEMM_DBG=$d timeout 120 python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2; done
