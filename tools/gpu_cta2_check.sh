#!/bin/bash
mkdir -p gpurun_out
GONOVA_TC2_CTA2=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" --no-header -p no:cacheprovider -x > gpurun_out/cta2_kernels.log 2>&1
echo "forced cta2 kernels rc=$?: $(tail -1 gpurun_out/cta2_kernels.log)"; grep -E "^E  " gpurun_out/cta2_kernels.log | head -5
GONOVA_TC2_CTA2=1 GONOVA_TC2_MH=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" --no-header -p no:cacheprovider -x > gpurun_out/cta2_kernels_mh2.log 2>&1
echo "forced cta2 mh2 kernels rc=$?: $(tail -1 gpurun_out/cta2_kernels_mh2.log)"; grep -E "^E  " gpurun_out/cta2_kernels_mh2.log | head -5
GONOVA_TC2_CTA2=1 timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/cta2_decode.log 2>&1
echo "forced cta2 decode rc=$?: $(tail -1 gpurun_out/cta2_decode.log)"; grep -E "^E  " gpurun_out/cta2_decode.log | head -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-table gpurun_out/launch_table_cta2.csv > gpurun_out/bench_cta2.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_cta2.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'])"
GONOVA_TC2_CTA2=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-first-chunk > gpurun_out/bench_nocta2.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_nocta2.log').read().strip().splitlines()[-1]);print('no cta2:',d['value'],d['ms_per_step'])"
