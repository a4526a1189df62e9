#!/bin/bash
# Whole-ResBlock kernel: correctness first (variants + a few decode tests, both dtypes), then the per-launch table.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -s -x --no-header -p no:cacheprovider -k "variants or golden_plain or one_chunk or ragged_batch_equals or small_lengths" > gpurun_out/chain_tests.log 2>&1
echo "tests rc=$?: $(tail -1 gpurun_out/chain_tests.log)"
grep -E "variant|FAILED|Error|rror:|DEAD|assert" gpurun_out/chain_tests.log | head -40
timeout 600 python bench.py --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0 --profile-table gpurun_out/chain_table.csv > gpurun_out/chain_bench.json 2> gpurun_out/chain_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/chain_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/chain_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY
grep -E "chain|resblocks.6|resblocks.3|ups" gpurun_out/chain_table.csv
