export MDC_LIB_PATH=$PWD/mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/libmdc_b200_dev.so
for g in 148 112 96 80 64 48 32; do echo "== MDC_GEMM_GRID=$g"; MDC_GEMM_GRID=$g timeout 200 python tools/pipeline_probe.py 60 16,4,6,0 | tail -1; done
