"""Shared-memory wavefronts per SASS instruction (ideal vs excessive) for one kernel of an .ncu-rep."""
import csv, subprocess, sys, io
rep = sys.argv[1]; pat = sys.argv[2]
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
    iW = hdr.index('L1 Wavefronts Shared'); iX = hdr.index('L1 Wavefronts Shared Excessive'); iI = hdr.index('L1 Wavefronts Shared Ideal')
    tot = sum(int(r[iW] or 0) for r in data); ex = sum(int(r[iX] or 0) for r in data)
    print(kern[:60], 'wavefronts', tot, 'excessive', ex)
    agg = {}
    for n, r in enumerate(data):
        w = int(r[iW] or 0)
        if w == 0: continue
        op = r[iS].split()
        op = op[1] if op[0].startswith('@') else op[0]
        key = op
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += int(r[iE]); a[1] += w; a[2] += int(r[iX] or 0); a[3] += 1
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('  %-28s sites %4d exec %9d wavefronts %9d (%.2f/instr) excessive %9d' % (k, a[3], a[0], a[1], a[1] / max(a[0], 1), a[2]))
    # worst sites
    worst = sorted(data, key=lambda r: -int(r[iX] or 0))[:12]
    for r in worst:
        print('   worst: %s exec %s wf %s ideal %s' % (r[iS].strip()[:60], r[iE], r[iW], r[iI]))
    break
