#!/bin/bash
# Round-2 check: GPU tests (parity lines kept), smoke, one bench run.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?: $(tail -1 gpurun_out/r2_tests.log)"
grep -E "FAILED|Error|error" gpurun_out/r2_tests.log | head -20
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --profile-table gpurun_out/r2_launch_table.csv > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("tf32"), d.get("stock_torch_gpu"), d.get("streams512"), d["first_chunk"])
    for r in d["roofline_kernels"]:
        print(r["kernel"], round(r["ms"], 4), round(r["frac"], 3))
except Exception as e:
    print("bench parse failed", e)
PY
