import csv, subprocess, sys, io
from collections import Counter
rep, pat, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
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
    c = Counter(); n = Counter()
    for r in data[a:b]:
        op = r[iS].split()
        if op[0].startswith('@'): op = op[1:]
        c[op[0]] += int(r[iE]); n[op[0]] += 1
    tot = sum(c.values())
    print('range', a, b, 'instr %.2fM' % (tot / 1e6))
    for o, v in c.most_common(25): print('  %-22s %7.3fM  sites %d' % (o, v / 1e6, n[o]))
    if len(sys.argv) > 5:
        for i, r in enumerate(data[a:b]): print(a + i, r[iS].strip()[:80], r[iE])
    break
