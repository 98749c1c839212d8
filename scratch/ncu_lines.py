"""Instructions executed / warp-state samples of one kernel per CUDA source line: joins the SASS rows of an ncu report
(source page) with the line table of the SAME build (nvdisasm --print-line-info on the object's cubin), by instruction
order.   python scratch/ncu_lines.py rep.ncu-rep 'stack_b_kernel<(int)16' pysilent_b200/build/stack_fused.o _ZN6silent14stack_b_kernelILi16E [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, pat, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
lines = []   # (file, line) per instruction of the function
inside = False; cur = ('?', 0)
for l in sass:
    if l.startswith('.text.'):
        inside = l.startswith('.text.' + mangled)
        continue
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
    iE = hdr.index('Instructions Executed'); iP = hdr.index('# Samples')
    print(kern[:90], 'ncu rows', len(data), 'nvdisasm instructions', len(lines))
    if len(data) != len(lines): print('WARNING: instruction counts differ -- report and object are different builds')
    ins = defaultdict(int); smp = defaultdict(int)
    for r, ln in zip(data, lines):
        ins[ln] += int(r[iE]); smp[ln] += int(r[iP] or 0)
    ti, ts = sum(ins.values()), sum(smp.values())
    print('total instr %.2fM samples %d' % (ti / 1e6, ts))
    src_cache = {}
    for ln, v in sorted(ins.items(), key=lambda kv: -smp[kv[0]])[:top]:
        f = ln[0]
        if f not in src_cache:
            path = os.path.join('pysilent_b200/csrc', f)
            src_cache[f] = open(path).read().split('\n') if os.path.exists(path) else []
        text = src_cache[f][ln[1] - 1].strip()[:90] if 0 < ln[1] <= len(src_cache[f]) else ''
        print('%-16s:%-5d instr %5.2f%% samples %5.2f%%  %s' % (f, ln[1], 100 * v / ti, 100 * smp[ln] / ts, text))
    break
