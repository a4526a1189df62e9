"""Selected raw metrics per captured launch from `ncu -i x.ncu-rep --page raw --csv` output -> the text summaries in profiles/.
python tools/ncu_raw_summary.py raw.csv "header line" > profiles/xyz_summary.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
print("# " + sys.argv[2])
print("# ncu --set full --clock-control none (kernel runs cold and serialised under ncu: compare fractions)\n#")
print("# selected raw metrics per captured launch")
kn = hdr.index("Kernel Name")
for d in data:
    print("## " + d[kn][:100])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:<90} {d[i]} {units[i]}")
