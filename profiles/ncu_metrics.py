"""Print the headline metrics of every kernel in an ncu report.  usage: python profiles/ncu_metrics.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            k = hdr.index(w)
            print("%-88s %s %s" % (w, r[k], units[k]))
    for k, h in enumerate(hdr):                      # tensor-core / TMEM / clock metrics, whatever this ncu version calls them
        hl = h.lower()
        if h not in want and any(t in hl for t in ("tmem", "tensor", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.max", "cycles_elapsed.avg.per_second", "pipe_tc", "pipe_uniform")):
            print("%-88s %s %s" % (h, r[k], units[k]))
    print("-" * 60)
