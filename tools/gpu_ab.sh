#!/bin/bash
# A/B of environment switches inside one box: usage: bash tools/gpu_ab.sh "tagA:ENV=1 ENV2=0" "tagB:..." ...
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-first-chunk --profile-table gpurun_out/launch_table_$tag.csv > gpurun_out/bench_$tag.log 2>&1
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1]);print('$tag',round(d['value']),round(d['ms_per_step'],2),round(d['roofline']['achieved']),d['clocks']['sm_mhz'])"
done
