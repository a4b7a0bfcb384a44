#!/usr/bin/env python
"""Print the key metrics of an `ncu --page raw --csv` dump and, optionally, the share of executed instructions /
stall samples between barriers from the matching `--page source --csv` dump (gzip)."""
import csv, gzip, sys
raw = sys.argv[1]
r = list(csv.reader(open(raw)))
hdr, units = r[0], r[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for d in r[2:]:
    print('==', d[hdr.index('Kernel Name')][:80])
    for i, h in enumerate(hdr):
        if h in want:
            print(f'  {h} [{units[i]}] {d[i]}')
    st = [(h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), float(d[i].replace(',', '') or 0)) for i, h in enumerate(hdr)
          if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    print('  stalls:', ' '.join(f'{a}={b:.2f}' for a, b in sorted(st, key=lambda x: -x[1])[:7]))
if len(sys.argv) > 2:
    rows = list(csv.reader(gzip.open(sys.argv[2], 'rt')))
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
    tot = sum(int(x[idx['Instructions Executed']]) for x in data); samp = sum(int(x[idx['# Samples']]) for x in data)
    print('total warp inst', tot, 'samples', samp)
    seg = acc = sacc = start = 0
    for k, x in enumerate(data):
        ins = x[idx['Source']].strip(); acc += int(x[idx['Instructions Executed']]); sacc += int(x[idx['# Samples']])
        if ins.startswith('BAR.SYNC') or 'EXIT' in ins or k == len(data) - 1:
            if acc > 0.002 * tot or sacc > 0.002 * samp:
                print(f'  seg {seg} sass[{start}:{k}] inst {acc/tot*100:5.1f}%  samples {sacc/samp*100:5.1f}%  ends: {ins[:40]}')
            seg += 1; acc = sacc = 0; start = k + 1
