"""Summarise `ncu --page raw --csv` exports (gpurun_out/prof_*_raw.csv) into a markdown table for profiles/.
  python scripts/summarize_ncu.py TAG > profiles/TAG_kernels_ncu_summary.md"""
import csv, glob, os, sys
tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
rows = {}
for path in sorted(glob.glob(f"gpurun_out/prof_*_{tag}_raw.csv")):
    name = os.path.basename(path)[5:-len(f"_{tag}_raw.csv")]
    with open(path) as f:
        r = list(csv.reader(f))
    hdr = None
    for i, line in enumerate(r):
        if line and line[0] == "ID":
            hdr = i
            break
    if hdr is None:
        continue
    names, units, vals = r[hdr], r[hdr + 1], r[hdr + 2]
    d = {n: (v, u) for n, u, v in zip(names, units, vals)}
    rows[name] = d
print(f"# ncu `--set full --clock-control none` captures, tag `{tag}` (one steady-state launch per kernel)\n")
print("Driver: `scripts/prof_driver.py <what>`; capture: `scripts/gpu_profile.sh` (`-k regex:<kernel> -s 2 -c 1`).\n")
cols = list(rows)
print("| metric | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
print("| kernel | " + " | ".join(rows[c].get("Kernel Name", ("?", ""))[0][:60] for c in cols) + " |")
for k in KEYS:
    print(f"| `{k}` | " + " | ".join((rows[c].get(k, ("-", ""))[0] + " " + rows[c].get(k, ("", ""))[1]).strip() for c in cols) + " |")
