#!/bin/bash
# A/B timing of the whole-ResBlock kernel's knobs (per-launch table of the bench step).
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" python bench.py --steps 5 --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0 --profile-table gpurun_out/ab_$tag.csv > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || echo "$tag failed"
  echo "$tag: $(grep -E 'chain' gpurun_out/ab_$tag.csv | tr '\n' ' ') step $(python -c "import json;print(round(json.load(open('gpurun_out/ab_$tag.json'))['ms_per_step'],3))")"
}
for v in "$@"; do run "$v" $(echo $v | tr '+' ' '); done
