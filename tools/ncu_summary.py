"""Prints the handful of ncu raw-page metrics we track.  usage: python tools/ncu_summary.py raw.csv"""
import csv
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    for k in KEEP:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k:75s} {vals[i]:>18s} {units[i]}")
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
            v = float(vals[i]) if vals[i] not in ('', '-') else 0
            if v >= 0.05:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.3f}")
    print()
