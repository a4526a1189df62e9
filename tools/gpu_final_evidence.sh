#!/bin/bash
# Round-end evidence with the final build: tests, smoke, bench, ncu launch list + DRAM traffic of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -q --no-header -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?: $(tail -1 gpurun_out/final_tests.log)"
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --profile-table gpurun_out/launch_table_final.csv > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-first-chunk"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/launches_r01.csv
