#!/bin/bash
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.log 2>&1
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1]);fc=d['first_chunk'];print('$tag',round(d['value']),round(d['ms_per_step'],2),'first chunk graph p50',round(fc['p50_ms'],3),'eager',round(fc['eager_p50_ms'],3))"
done
