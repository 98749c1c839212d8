import csv, subprocess, sys, io
rep = sys.argv[1]; pat = sys.argv[2]; marks = sys.argv[3].split(',')
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
    print(kern[:80], len(data))
    for n, r in enumerate(data):
        if any(m in r[iS] for m in marks): print(n, r[iS].strip()[:70], 'exec', r[iE], 'samples', r[iP])
    break
