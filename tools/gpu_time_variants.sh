#!/bin/bash
# kernel-only durations (ncu gpu__time_duration) of epilogue variants of one stage-2 layer shape
mkdir -p gpurun_out
run() { tag=$1; shift; flags=$1; shift
  HOOK_FLAGS=$flags python tools/profile_conv.py "$@" > /dev/null 2>&1 || { echo "$tag plain failed"; return; }
  HOOK_FLAGS=$flags ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc2 --csv python tools/profile_conv.py "$@" 2>/dev/null | grep conv_tc2 | awk -F'","' -v t=$tag '{gsub(/"/,"",$NF); printf "%s %s us\n", t, $NF}' | tail -1
}
run raw_only_k3        0 64 60001 3 1 16 bf16 none 0
run raw_res_k3         0 64 60001 3 1 16 bf16 none 1
run lrelu_noraw_k3     8 64 60001 3 1 16 bf16 lrelu 0
run snake_noraw_k3     8 64 60001 3 1 16 bf16 snake_fast 0
run snake_raw_k3       0 64 60001 3 1 16 bf16 snake_fast 0
run snake_noraw_k11    8 64 60001 11 1 16 bf16 snake_fast 0
run snake_noraw_k11_tf32 8 64 60001 11 1 16 tf32 snake_fast 0
run c128_snake_noraw_k3 8 128 20000 3 1 16 bf16 snake_fast 0
run c128_snake_noraw_k11 8 128 20000 11 1 16 bf16 snake_fast 0
