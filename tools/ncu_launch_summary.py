"""Turn the ncu launch list of one bench command (tools/gpu_r2_evidence.sh: gpu__time_duration + dram bytes per launch) into
the two files profiles/ keeps: the compact list of ONE complete vocoder step (nct_to_nlc ... pcm_tail) and the traffic json.
python tools/ncu_launch_summary.py gpurun_out/r02_ncu_launch_list.csv profiles/r02_ncu_launch_list_one_step.csv profiles/r02_conv_traffic.json"""
import csv
import json
import re
import sys

src, out_csv, out_json = sys.argv[1:4]
rows = {}
with open(src) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    i = int(r["ID"])
    e = rows.setdefault(i, {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
    e[r["Metric Name"]] = float(r["Metric Value"])
launches = [rows[i] for i in sorted(rows)]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("gnv::", "")


# the LAST complete step: from the last nct_to_nlc_kernel that is followed by a pcm_tail_kernel
starts = [i for i, l in enumerate(launches) if "nct_to_nlc" in l["kernel"]]
step = None
for s in reversed(starts):
    for e in range(s, min(s + 200, len(launches))):
        if "pcm_tail" in launches[e]["kernel"]:
            step = launches[s:e + 1]
            break
    if step:
        break
assert step, "no complete step in the launch list"
with open(out_csv, "w") as f:
    f.write("# one complete bench step (B=64, T=500, bf16) under ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
            "dram__bytes_write.sum --clock-control none (cold-cache, serialised: compare shares)\n")
    f.write("idx,kernel,grid,duration_us,dram_read_MB,dram_write_MB\n")
    for i, l in enumerate(step):
        f.write(f'{i},{short(l["kernel"])},"{l["grid"]}",{l["gpu__time_duration.sum"] / 1e3:.1f},'
                f'{l["dram__bytes_read.sum"] / 1e6:.1f},{l["dram__bytes_write.sum"] / 1e6:.1f}\n')
by = {}
for l in step:
    k = short(l["kernel"]).split("<")[0]
    e = by.setdefault(k, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
    e["launches"] += 1
    e["ms"] += l["gpu__time_duration.sum"] / 1e6
    e["dram_bytes"] += l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
conv = [k for k in by if k.startswith("conv_")]
tot_ms = sum(e["ms"] for e in by.values())
conv_ms = sum(by[k]["ms"] for k in conv)
json.dump({"traffic_bytes": sum(by[k]["dram_bytes"] for k in conv), "conv_launches": sum(by[k]["launches"] for k in conv),
           "conv_ms_under_ncu": conv_ms, "step_ms_under_ncu": tot_ms, "share_of_step": conv_ms / tot_ms, "by_kernel": by,
           "source": "tools/gpu_r2_evidence.sh -> tools/ncu_launch_summary.py (final round-2 build)"}, open(out_json, "w"), indent=1)
print(len(step), "launches in the step;", f"conv family {conv_ms:.2f} of {tot_ms:.2f} ms under ncu, "
      f"{sum(by[k]['dram_bytes'] for k in conv) / 1e9:.1f} GB")
