#!/bin/bash
# Round-end refresh, one gpurun call: GPU tests, smoke(), the bench line, and the flow decoder's per-class launch tables.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu.txt
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke.txt 2>&1; tail -3 gpurun_out/r02_smoke.txt
python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_stderr.txt; tail -c 600 gpurun_out/r02_bench_line.json
python tools/flow_profile.py bf16 32 500 > gpurun_out/r02_flow_launch_table_bf16_B32_T500.csv 2>&1
python tools/flow_profile.py tf32 32 500 > gpurun_out/r02_flow_launch_table_tf32_B32_T500.csv 2>&1
python tools/flow_profile.py bf16 1 500 > gpurun_out/r02_flow_launch_table_bf16_B1_T500.csv 2>&1
python tools/flow_timing.py bf16 > gpurun_out/r02_flow_timing_bf16.txt 2>&1
python tools/flow_front_profile.py bf16 32 250 > gpurun_out/r02_flow_front_launch_table_bf16_B32_L250.csv 2>&1
python tools/flow_front_profile.py bf16 1 250 > gpurun_out/r02_flow_front_launch_table_bf16_B1_L250.csv 2>&1
python tools/flow_front_profile.py tf32 32 250 > gpurun_out/r02_flow_front_launch_table_tf32_B32_L250.csv 2>&1
tail -5 gpurun_out/r02_flow_timing_bf16.txt
