#!/bin/bash
# Runs the conv kernel tests under every A-slab / M-tile mode of conv_tc2, then the decode tests and
# a short bench under the best mode that passed.  Usage (on the GPU box): bash tools/gpu_v2_check.sh
mkdir -p gpurun_out
best=""
for slab in 1 0; do
  for mh in 0 2; do
    log=gpurun_out/v2_slab${slab}_mh${mh}.log
    GONOVA_TC2_SLAB=$slab GONOVA_TC2_MH=$mh timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" \
      --no-header -p no:cacheprovider -x > $log 2>&1
    rc=$?
    echo "slab=$slab mh=$mh rc=$rc: $(tail -1 $log)"
    if [ $rc -ne 0 ]; then grep -E "^E  |Error|error" $log | head -8; fi
    if [ $rc -eq 0 ] && [ -z "$best" ] && [ $mh -eq 2 ]; then
      # both mh modes of this slab mode must pass
      if tail -1 gpurun_out/v2_slab${slab}_mh0.log | grep -q passed && ! tail -1 gpurun_out/v2_slab${slab}_mh0.log | grep -q failed; then best=$slab; fi
    fi
  done
done
echo "best slab mode: '$best'"
if [ -z "$best" ]; then
  # pinpoint the faulting instruction of the first failing case
  first=$(grep -h -o "test_conv[a-z0-9_]*\[[a-z0-9_-]*\]" gpurun_out/v2_slab1_mh0.log | head -1)
  echo "sanitizing $first"
  GONOVA_TC2_SLAB=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_kernels.py::$first" -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/v2_sanitizer.log 2>&1
  grep -E "=========|Invalid|Illegal|at .*conv_tc" gpurun_out/v2_sanitizer.log | head -40
  exit 1
fi
export GONOVA_TC2_SLAB=$best
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_kernels.py -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/v2_decode.log 2>&1
echo "decode+kernels rc=$?: $(tail -1 gpurun_out/v2_decode.log)"
grep -E "^FAILED|^ERROR" gpurun_out/v2_decode.log | head
timeout 600 python bench.py --steps 5 --warmup 3 --profile-table gpurun_out/launch_table_v2.csv > gpurun_out/bench_v2.log 2>&1
echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_v2.log
