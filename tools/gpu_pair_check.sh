#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_plan_cache.py -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/pair_tests.log 2>&1
echo "tests rc=$?: $(tail -1 gpurun_out/pair_tests.log)"
grep -E "FAILED|rror:|assert|DEAD" gpurun_out/pair_tests.log | head -20
bash tools/gpu_pair_ab.sh A=1 GONOVA_PAIR_DIRECT=0
