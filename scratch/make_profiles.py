"""Turn gpurun_out/<rep>.ncu-rep (+ launch list CSV) into the tracked summaries under profiles/.

    python scratch/make_profiles.py gpurun_out/prof_final.ncu-rep gpurun_out/launches_final.csv r01_final 64

Writes profiles/<tag>_ncu_full_summary.csv (one row per captured kernel), profiles/<tag>_launches.csv (per-kernel totals
and shares of one step, from the gpu__time_duration launch list), profiles/<tag>_stalls.txt (warp-state samples per
kernel) and profiles/traffic.json (DRAM bytes per launch of the three dominant kernels; read by bench.py).
"""
import csv, io, json, subprocess, sys
from collections import OrderedDict, Counter

rep, launches, tag, batch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
           'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
           # the LSU's L1 data pipe (what bounds every kernel of the step) and what the TEX pipe carries beside it
           'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
           'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg',
           'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
           'l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum', 'l1tex__tex_writeback_active.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [m for m in METRICS if m in hdr]
with open('profiles/%s_ncu_full_summary.csv' % tag, 'w', newline='') as f:
    wr = csv.writer(f)
    wr.writerow(['Kernel Name'] + cols)
    wr.writerow([''] + [units[hdr.index(m)] for m in cols])
    traffic = OrderedDict()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        wr.writerow([name] + [r[hdr.index(m)] for m in cols])
        short = name.split('(')[0].split('<')[0].replace('void ', '').replace('silent::', '').strip()
        def mb(metric):
            v = float(r[hdr.index(metric)].replace(',', '')); u = units[hdr.index(metric)]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        traffic[short] = traffic.get(short, 0) + int(mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'))   # (variants add up)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name': blocks.append([r[1], None, []]); continue
    if r and r[0] == 'Address': blocks[-1][1] = r; continue
    if blocks and blocks[-1][1] and len(r) >= len(blocks[-1][1]) - 2: blocks[-1][2].append(r)
with open('profiles/%s_stalls.txt' % tag, 'w') as f:
    f.write('warp-state samples per kernel (ncu source page, all samples), share of the kernel\'s samples\n')
    seen = set()
    for kern, h3, data in blocks:
        if kern in seen: continue
        seen.add(kern)
        sc = [i for i, h in enumerate(h3) if h.startswith('stall_') and 'Not Issued' not in h]
        c = Counter()
        for r in data:
            for i in sc: c[h3[i][6:]] += int(r[i] or 0)
        total = max(sum(c.values()), 1)
        iE = h3.index('Instructions Executed')
        f.write('%s\n  warp instructions %d; ' % (kern[:110], sum(int(r[iE]) for r in data)) +
                ', '.join('%s %.1f%%' % (k, 100.0 * v / total) for k, v in c.most_common(9)) + '\n')
# launch list -> per-kernel totals of ONE step (the capture covers `steps` identical steps)
lines = [l for l in open(launches) if l.startswith('"')]
rd = list(csv.reader(lines))
h2 = rd[0]
ik, iv = h2.index('Kernel Name'), h2.index('Metric Value')
tot = Counter(); cnt = Counter()
for r in rd[1:]:
    if 'silent::' not in r[ik]: continue
    short = r[ik].split('(')[0].split('<')[0].replace('void ', '').replace('silent::', '').strip()
    tot[short] += float(r[iv].replace(',', '')); cnt[short] += 1
steps = min(cnt.values())
with open('profiles/%s_launches.csv' % tag, 'w', newline='') as f:
    wr = csv.writer(f)
    wr.writerow(['kernel', 'launches_per_step', 'ns_per_step', 'share_of_step'])
    total = sum(tot.values())
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        wr.writerow([k, cnt[k] // steps, '%.0f' % (v / steps), '%.4f' % (v / total)])
json.dump({'source': 'profiles/%s_ncu_full_summary.csv (ncu --set full, batch %d per launch)' % (tag, batch),
           'batch': batch, 'kernels': traffic, 'dram_bytes_per_step': sum(traffic.values())},
          open('profiles/traffic.json', 'w'), indent=1)
print(open('profiles/%s_launches.csv' % tag).read())
print(json.dumps(traffic))
