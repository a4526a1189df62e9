#!/bin/bash
mkdir -p gpurun_out
GONOVA_PAIR_CTA2=1 GONOVA_FUSE_MAX_C=128 timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pc2_decode.log 2>&1
echo "forced pair cta2 (C<=128) decode rc=$?: $(tail -1 gpurun_out/pc2_decode.log)"; grep -E "^E  |^FAILED" gpurun_out/pc2_decode.log | head -5
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-first-chunk --profile-table gpurun_out/launch_table_$tag.csv > gpurun_out/bench_$tag.log 2>&1
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1]);print('$tag',round(d['value']),round(d['ms_per_step'],2),round(d['roofline']['achieved']),d['clocks']['sm_mhz'])"; }
run base GONOVA_PAIR_CTA2=0
run pc2_c64 GONOVA_PAIR_CTA2=1
run pc2_c128 GONOVA_PAIR_CTA2=1 GONOVA_FUSE_MAX_C=128
run base2 GONOVA_PAIR_CTA2=0
