#!/bin/bash
# DRAM bytes + duration of every conv launch of one B=64 step (single-pass metrics: no replay, no memory restore)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-first-chunk"
$CMD > gpurun_out/traffic_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/traffic_plain.log; exit 1; }
N=$(python -c "import json;print(json.loads(open('gpurun_out/traffic_plain.log').read().strip().splitlines()[-1])['roofline']['kernel'].split(';')[1].split()[0])")
echo "conv launches per step: $N"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_tc2|conv_pair" -s $N -c $N --csv --log-file gpurun_out/traffic_r01.csv $CMD > gpurun_out/traffic_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/traffic_r01.csv
# one full-set capture of a fused pair (stage 2, k=11: source_resblocks.2 pair 0) and of a stage-1 conv, small batch
python tools/profile_decode.py 2 500 > gpurun_out/prof_decode_plain.log 2>&1 || { echo decode plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 0 -c 1 -f -o /tmp/prof_pair python tools/profile_decode.py 2 500 > gpurun_out/prof_pair_ncu.log 2>&1
ncu -i /tmp/prof_pair.ncu-rep --page details > gpurun_out/prof_pair_details.txt 2>/dev/null
ncu -i /tmp/prof_pair.ncu-rep --page raw --csv > gpurun_out/prof_pair_raw.csv 2>/dev/null
ncu -i /tmp/prof_pair.ncu-rep --page source --csv > gpurun_out/prof_pair_source.csv 2>/dev/null
ls -la /tmp/prof_pair.ncu-rep gpurun_out/prof_pair_*
