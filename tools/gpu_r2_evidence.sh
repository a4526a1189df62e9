#!/bin/bash
# Round-2 evidence with the final build: tests, smoke, bench (+ launch table), ncu launch list with DRAM bytes of one bench
# step, ncu --set full of the byte-moving kernels and of the whole-ResBlock kernel at the bench shape.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?: $(tail -1 gpurun_out/r02_tests.log)"
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --profile-table gpurun_out/r02_launch_table.csv > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --streams 0"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launch_list.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/r02_ncu_launch_list.csv
python tools/profile_decode.py 64 500 > gpurun_out/r02_prof_plain.log 2>&1 || { echo decode plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"source_kernel|stft_kernel|istft_kernel|f0_head|nct_to_nlc" -c 5 -f -o /tmp/r02_aux python tools/profile_decode.py 64 500 > gpurun_out/r02_aux_ncu.log 2>&1; echo "aux ncu rc=$?"
ncu -i /tmp/r02_aux.ncu-rep --page details > gpurun_out/r02_aux_details.txt 2>/dev/null
ncu -i /tmp/r02_aux.ncu-rep --page raw --csv > gpurun_out/r02_aux_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:conv_chain -c 1 -f -o /tmp/r02_chain python tools/profile_decode.py 64 500 > gpurun_out/r02_chain_ncu.log 2>&1; echo "chain ncu rc=$?"
ncu -i /tmp/r02_chain.ncu-rep --page details > gpurun_out/r02_chain_details.txt 2>/dev/null
ncu -i /tmp/r02_chain.ncu-rep --page raw --csv > gpurun_out/r02_chain_raw.csv 2>/dev/null
cat > /tmp/pcm_only.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from gonova_tts_b200 import pcm_tail
x = (torch.rand(64, 240000, device="cuda") * 2 - 1)
out = torch.empty(64, 240000, dtype=torch.int16, device="cuda")
for _ in range(3):
    pcm_tail(x, None, None, 0.99, want_i16=True, out_i16=out)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none -k regex:pcm_tail -s 2 -c 1 -f -o /tmp/r02_pcm python /tmp/pcm_only.py > gpurun_out/r02_pcm_ncu.log 2>&1; echo "pcm ncu rc=$?"
ncu -i /tmp/r02_pcm.ncu-rep --page details > gpurun_out/r02_pcm_details.txt 2>/dev/null
ncu -i /tmp/r02_pcm.ncu-rep --page raw --csv > gpurun_out/r02_pcm_raw.csv 2>/dev/null
python tools/chain_trace.py 64 500 > gpurun_out/r02_chain_trace.txt 2>&1
ls -la gpurun_out/r02_*
