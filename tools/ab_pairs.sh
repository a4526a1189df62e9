# Three interleaved bench runs per environment switch (edit the `run` lines): the step time moves +-3 % with the box and its power state.
F="--steps 20 --warmup 3 --no-cpu-baseline --no-first-chunk --no-tf32 --no-stock-torch --no-flow --streams 0"
run() { name=$1; shift; env "$@" python bench.py $F 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"; }
for i in 1 2 3; do
run base A=1
run fuse_k3_c128 GONOVA_FUSE_K3_MAX_C=128
run chain_c128 GONOVA_CHAIN_MAX_C=128
done
