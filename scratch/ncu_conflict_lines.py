"""Shared-memory wavefronts per CUDA source line of one kernel (excess over the ideal count = bank conflicts): joins the
SASS rows of an ncu report with nvdisasm's line table of the SAME build.
python scratch/ncu_conflict_lines.py rep.ncu-rep 'stack_b_kernel<(int)16' pysilent_b200/build/stack_fused.o _ZN6silent14stack_b_kernelILi16E"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, pat, obj, mangled = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
lines = []; inside = False; cur = ('?', 0)
for l in sass:
    if l.startswith('.text.'):
        inside = l.startswith('.text.' + mangled); continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): lines.append(cur)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name': blocks.append([r[1], None, []]); continue
    if r and r[0] == 'Address': blocks[-1][1] = r; continue
    if blocks and blocks[-1][1] and len(r) >= len(blocks[-1][1]) - 2: blocks[-1][2].append(r)
for kern, hdr, data in blocks:
    if pat not in kern: continue
    iS = hdr.index('Source'); iW = hdr.index('L1 Wavefronts Shared'); iI = hdr.index('L1 Wavefronts Shared Ideal')
    if len(data) != len(lines): print('WARNING: report and object are different builds', len(data), len(lines))
    agg = defaultdict(lambda: [0, 0, 0, set()])
    for r, ln in zip(data, lines):
        w = int(r[iW] or 0)
        if not w: continue
        op = r[iS].split(); op = op[1] if op[0].startswith('@') else op[0]
        a = agg[ln]; a[0] += 1; a[1] += w; a[2] += int(r[iI] or 0); a[3].add(op)
    cache = {}
    for ln, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] - kv[1][2])):
        f = ln[0]
        if f not in cache:
            path = os.path.join('pysilent_b200/csrc', f)
            cache[f] = open(path).read().split('\n') if os.path.exists(path) else []
        text = cache[f][ln[1] - 1].strip()[:80] if 0 < ln[1] <= len(cache[f]) else ''
        print('%-16s:%-5d %-16s sites %3d wavefronts %8d excess %8d | %s' % (f, ln[1], ','.join(sorted(a[3])), a[0], a[1], a[1] - a[2], text))
    break
