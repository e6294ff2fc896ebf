"""Condense `ncu -i X.ncu-rep --page raw --csv` into a small per-kernel table (markdown).
usage: ncu -i rep --page raw --csv | python tools/ncu_kernel_summary.py > profiles/xyz.md"""
import csv, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]; units = rows[1]
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
ki = h.index('Kernel Name')
print('| metric | ' + ' | '.join(r[ki][:48] for r in rows[2:]) + ' |')
print('|---|' + '---|' * (len(rows) - 2))
for k in KEYS:
    if k in h:
        i = h.index(k)
        print(f'| {k} [{units[i]}] | ' + ' | '.join(r[i] for r in rows[2:]) + ' |')
