#!/bin/bash
mkdir -p gpurun_out
GONOVA_CHAIN_MAX_K=7 timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x --no-header -p no:cacheprovider -k "golden_plain or one_chunk or ragged or small_lengths or ten_second" 2>&1 | tail -3
bash tools/gpu_chain_ab.sh A=1 GONOVA_CHAIN_MAX_K=7 B=1 GONOVA_CHAIN_MAX_K=7+C=1
grep -E "resblocks.7" gpurun_out/ab_A=1.csv gpurun_out/ab_GONOVA_CHAIN_MAX_K=7.csv | cut -d, -f2,4 | tr '\n' ' '
