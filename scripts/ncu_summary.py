"""Summarise an .ncu-rep (read here, no GPU): key metrics per kernel launch + top stall reasons."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
stall = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio", h)]
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    print("=====", r[idx["ID"]], name, r[idx["Grid Size"]], r[idx["Block Size"]])
    for k in KEYS:
        if k in idx:
            print(f"  {k:62s} {r[idx[k]]:>16s} {units[idx[k]]}")
    st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:40s} {v:8.2f}")
