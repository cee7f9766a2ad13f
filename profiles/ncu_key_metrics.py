"""Dump the key metrics of an .ncu-rep (ncu -i ... --page raw --csv) as text."""
import csv, subprocess, sys
PATS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__issue_active.avg.pct', 'sm__pipe_tensor_cycles_active.avg', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'launch__waves', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, units = r[0], r[1]
for row in r[2:]:
    print('=' * 100)
    print(row[hdr.index('Kernel Name')][:160])
    for i, h in enumerate(hdr):
        if any(p in h for p in PATS) and row[i] not in ('', 'n/a'):
            print(f'  {h} = {row[i]} {units[i]}')
