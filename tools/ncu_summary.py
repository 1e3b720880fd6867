"""Key metrics + stall ratios of every kernel in an .ncu-rep.   python tools/ncu_summary.py <rep>"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('----', d['Kernel Name'][:70])
    for w in KEYS:
        if w in d:
            print('  ', w, d[w], units[hdr.index(w)])
    st = []
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('_per_issue_active.ratio'):
            try:
                st.append((float(d[k].replace(',', '')), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('   stalls:', ', '.join(f'{k} {v:.2f}' for v, k in sorted(st, reverse=True)[:7]))
