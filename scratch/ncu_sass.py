"""Dump SASS with dynamic counts for one kernel of an .ncu-rep."""
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
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iP = hdr.index('# Samples')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    for n, r in enumerate(data):
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        sts = ' '.join('%s:%d' % (b, a) for a, b in st if a > 0)
        print('%4d %9s %5s  %-80s %s' % (n, r[iE], r[iP], r[iS].strip()[:80], sts))
    break
