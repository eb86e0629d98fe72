"""Summarise an `ncu --page raw --csv` export: a few decisive metrics per kernel (first launches)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
have = [w for w in want if w in idx]
seen = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > per:
        continue
    print('---', name, '#%d' % seen[name])
    for w in have:
        print(f"   {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}")
