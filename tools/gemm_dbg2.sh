#!/bin/bash
# Developer experiment 2: what in the epilogue slows the main loop?  16 = staging but no TMA stores, 32 = no TMEM reads
export MDC_LIB_PATH=$PWD/mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/libmdc_b200_dev.so
export MDC_GEMM_2CTA=0
for d in 0 16 32 48 1; do echo "== single CTA, MDC_GEMM_DBG=$d"; MDC_GEMM_DBG=$d timeout 120 python tools/gemm_probe.py 64 20 qkv,fc1,proj,fc2; done
