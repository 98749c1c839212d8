"""Summarise an .ncu-rep: key raw metrics + stall reasons per kernel (reads via `ncu -i`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('---')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('  %-70s %s %s' % (w, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            v = float(r[i].replace(',', '') or 0)
            if v > 3: st.append((v, h.split('issue_stalled_')[1].split('_per')[0]))
    print('  stalls:', ', '.join('%s %.0f' % (n, v) for v, n in sorted(st, reverse=True)))
