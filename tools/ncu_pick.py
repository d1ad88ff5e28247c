"""Print a compact selection of metrics from an `ncu --page raw --csv` export: tools/ncu_pick.py raw.csv [name-regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if pat and not pat.search(r[ki]):
        continue
    print("=====", r[ki].split("(")[0][-60:], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]:>14s} {units[i]}")
    st = [(float(r[i]), hdr[i].split("issue_stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(hdr)
          if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio") and r[i]]
    print("  stalls:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
