"""Prints the key metrics of every kernel in an `ncu --page raw --csv` dump.
usage: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_keys.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
STALL = "smsp__average_warps_issue_stalled_"
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("=" * 100)
    for k in KEYS:
        if k in idx:
            print("%-70s %s %s" % (k, r[idx[k]][:90], rows[1][idx[k]]))
    st = [(float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for h, i in idx.items()
          if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
    print("stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
