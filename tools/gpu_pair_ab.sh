#!/bin/bash
# Ablation timing of the fused pair kernel's epilogues (per-launch table of the bench step).
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" python bench.py --steps 4 --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0 --profile-table gpurun_out/pab_$tag.csv > gpurun_out/pab_$tag.json 2> gpurun_out/pab_$tag.err || echo "$tag failed"
  echo "$tag: $(grep -E 'resblocks.2.pair0|resblocks.7.pair0|resblocks.8.pair0|resblocks.8.pair2|resblocks.6' gpurun_out/pab_$tag.csv | cut -d, -f2,4 | tr '\n' ' ') step $(python -c "import json;print(round(json.load(open('gpurun_out/pab_$tag.json'))['ms_per_step'],3))")"
}
for v in "$@"; do run "$v" $(echo $v | tr '+' ' '); done
