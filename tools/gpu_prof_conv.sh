#!/bin/bash
# ncu --set full on one layer shape; exports small CSV summaries (the .ncu-rep itself stays on the box if large).
# usage: bash tools/gpu_prof_conv.sh <tag> <profile_conv.py args...>
tag=$1; shift
mkdir -p gpurun_out
python tools/profile_conv.py "$@" > gpurun_out/prof_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 1 -c 1 -f -o /tmp/prof_${tag} python tools/profile_conv.py "$@" > gpurun_out/prof_${tag}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_${tag}.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
ncu -i /tmp/prof_${tag}.ncu-rep --page details > gpurun_out/prof_${tag}_details.txt 2>/dev/null
ls -la /tmp/prof_${tag}.ncu-rep gpurun_out/prof_${tag}_*
size=$(stat -c %s /tmp/prof_${tag}.ncu-rep)
if [ $size -lt 20000000 ]; then cp /tmp/prof_${tag}.ncu-rep gpurun_out/; fi
