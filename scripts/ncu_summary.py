"""Key metrics of every kernel in an .ncu-rep (ncu --set full) as the text summaries kept under profiles/:
python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("=====", d["Kernel Name"][:90])
    for k in KEYS:
        if k in d:
            print("  ", k, d[k], u[k])
    st = []
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]:
            st.append((float(d[k]), k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    for v, n in sorted(st, reverse=True)[:8]:
        print("   stall", n, round(v, 3))
