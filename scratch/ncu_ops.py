"""Per-kernel dynamic opcode histogram + hottest source lines from an .ncu-rep source page."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 22
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--print-source', sys.argv[4]] if len(sys.argv) > 4 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = None; hdr = None; blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name': kern = r[1]; blocks.append([kern, None, []]); continue
    if r and r[0] in ('Address', '#'): blocks[-1][1] = r; continue
    if blocks and blocks[-1][1] and len(r) >= len(blocks[-1][1]) - 2: blocks[-1][2].append(r)
for kern, hdr, data in blocks:
    if pat not in kern: continue
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iP = hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iP]) for r in data)
    print('==', kern[:90], 'warp-instr', tot, 'samples', totS)
    c = Counter(); s = Counter()
    for r in data:
        op = r[iS].split()
        if op[0].startswith('@'): op = op[1:]
        o = '.'.join(op[0].split('.')[:2]) if op[0].startswith(('LDS','STS','LDG','STG','IMAD')) else op[0].split('.')[0]
        c[o] += int(r[iE]); s[o] += int(r[iP])
    for o, v in c.most_common(top): print('  %-14s %6.2f%% instr  %6.2f%% samples' % (o, 100 * v / tot, 100 * s[o] / max(totS,1)))
