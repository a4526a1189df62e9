#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" python bench.py --steps 3 --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0 --profile-table gpurun_out/f0_$tag.csv > gpurun_out/f0_$tag.json 2> gpurun_out/f0_$tag.err || echo "$tag failed"
  echo "$tag: $(grep -E 'f0_predictor.condnet|conv_pre|conv_post|ups|source_downs' gpurun_out/f0_$tag.csv | cut -d, -f2,4 | tr '\n' ' ') step $(python -c "import json;print(round(json.load(open('gpurun_out/f0_$tag.json'))['ms_per_step'],3))")"
}
for v in "$@"; do run "$v" $(echo $v | tr '+' ' '); done
