#!/bin/bash
# Round-2 evidence for the flow decoder, one gpurun call: per-class launch table of one estimator evaluation
# (bf16 + tf32, B = 32 and B = 1), CTA-0 timelines of the three flow_blk modes, and ncu --set full of the four new kernels.
set -x
mkdir -p gpurun_out
python tools/flow_profile.py bf16 32 500 > gpurun_out/r02_flow_launch_table_bf16_B32_T500.csv 2>&1
python tools/flow_profile.py tf32 32 500 > gpurun_out/r02_flow_launch_table_tf32_B32_T500.csv 2>&1
python tools/flow_profile.py bf16 1 500 > gpurun_out/r02_flow_launch_table_bf16_B1_T500.csv 2>&1
python tools/flow_timing.py bf16 > gpurun_out/r02_flow_timing_bf16.txt 2>&1
for m in 0 1 2; do TRACE_LINES=400 python tools/flow_blk_trace.py $m > gpurun_out/r02_flow_blk_timeline_mode$m.txt 2>&1; done
python tools/flow_one.py bf16 32 500 1 > gpurun_out/flow_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'flow_blk|flow_attn_tc' -s 4 -c 4 -o gpurun_out/prof_flow_r02 \
    python tools/flow_one.py bf16 32 500 1 > gpurun_out/flow_one_ncu.log 2>&1
tail -2 gpurun_out/flow_one_ncu.log
