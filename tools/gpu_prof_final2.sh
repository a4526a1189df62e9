#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 0 -c 1 -f -o /tmp/prof_pair64 python tools/profile_decode.py 64 500 > gpurun_out/prof_pair64_ncu.log 2>&1
ncu -i /tmp/prof_pair64.ncu-rep --page source --csv > gpurun_out/prof_pair64_source.csv 2>/dev/null
ncu -i /tmp/prof_pair64.ncu-rep --page source --csv --print-source cuda > gpurun_out/prof_pair64_source_cuda.csv 2>/dev/null
ls -la gpurun_out/prof_pair64_source*
