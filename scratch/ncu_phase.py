"""Stall-reason totals per SASS index range for one kernel: ncu_phase.py rep kernel b0,b1,b2..."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; pat = sys.argv[2]; bounds = [int(x) for x in sys.argv[3].split(',')]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name': blocks.append([r[1], None, []]); continue
    if r and r[0] == 'Address': blocks[-1][1] = r; continue
    if blocks and blocks[-1][1] and len(r) >= len(blocks[-1][1]) - 2: blocks[-1][2].append(r)
for kern, hdr, data in blocks:
    if pat not in kern: continue
    iE = hdr.index('Instructions Executed'); iP = hdr.index('# Samples')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    bounds = bounds + [len(data)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        c = Counter(); ins = 0
        for r in data[a:b]:
            ins += int(r[iE])
            for i in stall_cols: c[hdr[i][6:]] += int(r[i] or 0)
        tot = sum(c.values())
        print('[%d,%d) instr %.2fM samples %d: ' % (a, b, ins / 1e6, tot) + ' '.join('%s:%d' % kv for kv in c.most_common(7)))
    break
