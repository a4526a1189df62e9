#!/bin/bash
mkdir -p gpurun_out
python tools/profile_decode.py 16 500 > gpurun_out/prof_decode_plain.log 2>&1 || { echo decode plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s ${PAIR_SKIP:-0} -c 1 -f -o /tmp/prof_pair python tools/profile_decode.py 16 500 > gpurun_out/prof_pair_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_pair.ncu-rep --page details > gpurun_out/prof_pair_${PAIR_TAG:-k11}_details.txt 2>/dev/null
ncu -i /tmp/prof_pair.ncu-rep --page raw --csv > gpurun_out/prof_pair_${PAIR_TAG:-k11}_raw.csv 2>/dev/null
ncu -i /tmp/prof_pair.ncu-rep --page source --csv > gpurun_out/prof_pair_${PAIR_TAG:-k11}_source.csv 2>/dev/null
ls -la /tmp/prof_pair.ncu-rep
