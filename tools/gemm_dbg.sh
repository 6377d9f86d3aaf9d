#!/bin/bash
# Developer experiment: which part of gemm_tc_kernel bounds the encoder shapes?  MDC_GEMM_DBG bits (dev build only):
# 1 = no epilogue body, 2 = every load from tile (0,0), 4 = no loads, 8 = no MMAs.   usage (GPU box): tools/gemm_dbg.sh <tag>
tag=${1:-g}
export MDC_LIB_PATH=$PWD/mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/libmdc_b200_dev.so
for d in 0 1 2 3 4 5 8 9 12 13; do
  echo "== MDC_GEMM_DBG=$d"
  MDC_GEMM_DBG=$d python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2,patch,crosskv
done > gpurun_out/${tag}_gemm_dbg.log 2>&1
