"""Summarise an .ncu-rep (raw page) into a compact table: python scripts/ncu_summary.py file.ncu-rep [regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    print("==", name[:110])
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:70s} {r[i]:>18s} {units[i]}")
