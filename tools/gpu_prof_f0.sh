#!/bin/bash
mkdir -p gpurun_out
python tools/profile_decode.py 64 500 > gpurun_out/prof_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -f -o /tmp/prof_f0 python tools/profile_decode.py 64 500 > gpurun_out/prof_f0_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_f0.ncu-rep --page details > gpurun_out/prof_f0_details.txt 2>/dev/null
ncu -i /tmp/prof_f0.ncu-rep --page raw --csv > gpurun_out/prof_f0_raw.csv 2>/dev/null
GONOVA_TC2_DEBUG=1 python tools/profile_decode.py 64 500 2>&1 | head -5
