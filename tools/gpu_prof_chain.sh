#!/bin/bash
# ncu --set full of one conv_chain_kernel launch (B = 16, T = 500), details + raw + per-instruction source page.
mkdir -p gpurun_out
python tools/profile_decode.py 16 500 > gpurun_out/prof_decode_plain.log 2>&1 || { echo decode plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_chain -s ${CHAIN_SKIP:-0} -c 1 -f -o /tmp/prof_chain python tools/profile_decode.py 16 500 > gpurun_out/prof_chain_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_chain.ncu-rep --page details > gpurun_out/prof_chain_details.txt 2>/dev/null
ncu -i /tmp/prof_chain.ncu-rep --page raw --csv > gpurun_out/prof_chain_raw.csv 2>/dev/null
ncu -i /tmp/prof_chain.ncu-rep --page source --csv > gpurun_out/prof_chain_source.csv 2>/dev/null
ls -la /tmp/prof_chain.ncu-rep
