import csv, subprocess, sys, io
from collections import Counter, OrderedDict
rep, pat = sys.argv[1], sys.argv[2]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name': blocks.append([r[1], None, []]); continue
    if r and r[0] == 'Address': blocks[-1][1] = r; continue
    if blocks and blocks[-1][1] and len(r) >= len(blocks[-1][1]) - 2: blocks[-1][2].append(r)
for kern, hdr, data in blocks:
    if pat not in kern: continue
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed')
    iW = hdr.index('L1 Wavefronts Shared'); iI = hdr.index('L1 Wavefronts Shared Ideal')
    groups = OrderedDict()
    for n, r in enumerate(data):
        w = int(r[iW] or 0)
        if not w: continue
        op = r[iS].split(); op = op[1] if op[0].startswith('@') else op[0]
        key = (op, r[iE], n // 150)
        g = groups.setdefault(key, [0, 0, 0, n])
        g[0] += 1; g[1] += w; g[2] += int(r[iI] or 0)
    for k, g in groups.items():
        print('%-8s exec %8s block@%4d sites %3d wavefronts %9d ideal %9d ratio %.2f' % (k[0], k[1], g[3], g[0], g[1], g[2], g[1] / max(g[2], 1)))
    break
