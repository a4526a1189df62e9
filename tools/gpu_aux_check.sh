#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_decode.py tests/test_gpu_incremental.py -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/aux_tests.log 2>&1
echo "tests rc=$?: $(tail -1 gpurun_out/aux_tests.log)"
grep -E "source: kernel|FAILED|rror:|assert" gpurun_out/aux_tests.log | head -20
timeout 600 python bench.py --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0 --profile-table gpurun_out/aux_table.csv > gpurun_out/aux_bench.json 2> gpurun_out/aux_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/aux_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"])
for r in d["roofline_kernels"]:
    print(r["kernel"], round(r["ms"], 4), round(r["frac"], 3))
PY
