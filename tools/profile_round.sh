#!/bin/bash
# One GPU-box pass that produces everything under profiles/: tests, bench line, traces, launch list, ncu --set full captures.
# usage: tools/profile_round.sh <tag>     (run through gpurun; writes gpurun_out/<tag>_*)
# Every ncu command is preceded by a plain run of the same command line (B200_PROFILING.md).
tag=${1:-p}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_tests.log 2>&1
python bench.py > $o/${tag}_bench.log 2> $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_ref.log 2> $o/${tag}_ref.err
python tools/gemm_probe.py 64 20 > $o/${tag}_gemm.log 2>&1
python tools/decode_trace.py 64 99 50 > $o/${tag}_trace.log 2>&1
MDC_DECODE_IPC=16 python tools/decode_trace.py 64 99 50 > $o/${tag}_trace16.log 2>&1
python tools/decode_variants.py 99 64,0,0 64,8,0 64,16,0 128,16,0 64,8,2 256,8,2 > $o/${tag}_variants.log 2>&1
python tools/pipeline_probe.py 48 16,4,6,0 8,4,6,2 > $o/${tag}_pipeline.log 2>&1
python bench.py --profile > $o/${tag}_plain_profile.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $o/${tag}_launches.csv python bench.py --profile > $o/${tag}_ncu_launch.log 2>&1
python tools/gemm_probe.py 64 1 fc1 > $o/${tag}_plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -f -o $o/${tag}_gemm_fc1 python tools/gemm_probe.py 64 1 fc1 > $o/${tag}_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_umma_kernel -c 1 -f -o $o/${tag}_attn python bench.py --profile > $o/${tag}_ncu_attn.log 2>&1
python tools/decode_variants.py 99 64,0,0 > $o/${tag}_plain_dec.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_fused_kernel -s 1 -c 1 -f -o $o/${tag}_decode python tools/decode_variants.py 99 64,0,0 > $o/${tag}_ncu_decode.log 2>&1
python tools/decode_variants.py 99 64,16,0 > $o/${tag}_plain_dec16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_fused_kernel -s 1 -c 1 -f -o $o/${tag}_decode16 python tools/decode_variants.py 99 64,16,0 > $o/${tag}_ncu_decode16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layernorm_rows_kernel -c 1 -f -o $o/${tag}_ln python bench.py --profile > $o/${tag}_ncu_ln.log 2>&1
python tools/iou_probe.py > $o/${tag}_iou.log 2>&1
# the trained geometry T on the per-operation chain: timing, pipeline, launch list
python tools/config_t_probe.py 64 99 --pipeline > $o/${tag}_cfgt.log 2>&1
python tools/config_t_probe.py 64 4 > $o/${tag}_plain_cfgt.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/${tag}_cfgt_launches.csv python tools/config_t_probe.py 64 4 > $o/${tag}_ncu_cfgt.log 2>&1
