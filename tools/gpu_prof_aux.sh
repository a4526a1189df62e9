#!/bin/bash
mkdir -p gpurun_out
python tools/profile_decode.py 16 500 > gpurun_out/prof_decode_plain.log 2>&1 || { echo decode plain failed; exit 1; }
ncu --set full --clock-control none -k regex:"source_kernel|stft_kernel|istft_kernel" -c 3 -f -o /tmp/prof_aux python tools/profile_decode.py 16 500 > gpurun_out/prof_aux_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_aux.ncu-rep --page details > gpurun_out/prof_aux_details.txt 2>/dev/null
grep -E "^  [a-z_:]+.*\(|Duration|DRAM Throughput|Compute \(SM\) Throughput|Executed Ipc Active|Issued Warp Per|Registers Per|Achieved Occupancy|Theoretical Occupancy|Warp Cycles Per Issued|Block Limit|L1/TEX Hit|Shared Memory Config|Bank conflicts|Local" gpurun_out/prof_aux_details.txt | head -60
