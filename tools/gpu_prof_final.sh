#!/bin/bash
# ncu --set full of ONE launch each, at the bench shape (B=64, T=500, bf16): the k = 11 fused pair (stage 2) and a
# C = 128 k = 11 conv (stage 1).  Summaries (details page + selected raw metrics) come back in gpurun_out/.
mkdir -p gpurun_out
python tools/profile_decode.py 64 500 > gpurun_out/prof_final_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/prof_final_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 0 -c 1 -f -o /tmp/prof_pair64 python tools/profile_decode.py 64 500 > gpurun_out/prof_pair64_ncu.log 2>&1
ncu -i /tmp/prof_pair64.ncu-rep --page details > gpurun_out/prof_pair64_details.txt 2>/dev/null
ncu -i /tmp/prof_pair64.ncu-rep --page raw --csv > gpurun_out/prof_pair64_raw.csv 2>/dev/null
# conv_tc2 launch order in a decode: conv_pre, source_downs.0, srb0 x6, ups.0, rb0-2 x18, source_downs.1, srb1 x6, ups.1, rb3 x6, rb4 x6, rb5 convs1.0 = index 47
ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 52 -c 1 -f -o /tmp/prof_tc2_64 python tools/profile_decode.py 64 500 > gpurun_out/prof_tc2_64_ncu.log 2>&1
ncu -i /tmp/prof_tc2_64.ncu-rep --page details > gpurun_out/prof_tc2_64_details.txt 2>/dev/null
ncu -i /tmp/prof_tc2_64.ncu-rep --page raw --csv > gpurun_out/prof_tc2_64_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/prof_*64*
