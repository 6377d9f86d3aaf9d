#!/bin/bash
# Developer experiment 3: are the epilogue's TMA stores expensive in the TMA unit or in the L2 / HBM write path?  64 = every store to tile (0,0)
export MDC_LIB_PATH=$PWD/mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/libmdc_b200_dev.so
for c in 0 1; do for d in 0 64 16; do echo "== MDC_GEMM_2CTA=$c MDC_GEMM_DBG=$d"; MDC_GEMM_2CTA=$c MDC_GEMM_DBG=$d timeout 120 python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2; done; done
