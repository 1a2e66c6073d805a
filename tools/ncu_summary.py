"""Key counters of every launch in an ncu report (development aid). usage: ncu_summary.py <report.ncu-rep>"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:60], r[idx['Grid Size']] if 'Grid Size' in idx else '')
    for w in want:
        if w in idx:
            print(f'  {w}: {r[idx[w]]}')
    st = sorted(((float(r[idx[h]]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]) for h in stalls), reverse=True)
    print('  stalls per issue:', ', '.join(f'{n} {v:.2f}' for v, n in st[:8]))
