"""Phase H with the column sums of two tile rows interleaved per slot (float4 = row 2p, row 2p+1, both frame lanes):
one LDS.128 fetches a tap for two rows, a wavefront is a quarter-warp (8 lanes x 16 B).  Counts wavefronts per load
over the library's real tap tables, for several group paddings, against today's LDS.64 layout (half-warps, 8-byte
slots, 18 slots per 16 byte-columns)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysilent_b200.util.zoom.from_image import PyramidPlan
shape = (1080, 1920, 3) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split('x'))
center = (288, 192) if len(sys.argv) < 3 else tuple(int(v) for v in sys.argv[2].split('x'))
plan = PyramidPlan(shape, torch.uint8, 3, center, 2 ** .5)
L, w = plan.levels, plan.w
IX = [plan.level_tables(s)[2] for s in range(L)]
FC = 3; TW = 72; NT = 128

def wf(addrs, lanes, nbank):
    tot = 0
    for g in range(0, 32, lanes):
        banks = {}
        for a in addrs[g:g + lanes]:
            if a is None: continue
            banks.setdefault(a % nbank, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot

def run(G, lanes, nbank, mapping):
    out = []
    for s in range(L):
        ix = IX[s]; ok = ix[:, 0] >= 0
        total = 0; ideal = 0
        for bx in range((w + TW - 1) // TW):
            cols = [ox for ox in range(bx * TW, min(w, (bx + 1) * TW)) if ok[ox]]
            if not cols: continue
            lo = min(ix[ox].min() for ox in cols); wlo4 = (lo * FC // 16) * 16
            for base in range(0, 3 * TW, NT):
                for wi in range(base, min(base + NT, 3 * TW), 32):
                    for i in range(6):
                        addrs = []
                        for l in range(32):
                            item = wi + l
                            if mapping == 'cc':  c = item % 3; col = item // 3
                            else: c = item // TW; col = item % TW
                            ox = bx * TW + col
                            if item >= 3 * TW or ox >= w: addrs.append(None); continue
                            b = ix[ox][i] * FC - wlo4 + c if ok[ox] else 0
                            addrs.append(b + (G - 16) * (b >> 4))
                        total += wf(addrs, lanes, nbank)
                        ideal += sum(1 for g in range(0, 32, lanes) if any(a is not None for a in addrs[g:g + lanes]))
        out.append(round(total / max(ideal, 1), 2))
    return out

print('today  LDS.64  G=18 half-warps :', run(18, 16, 16, 'cc'))
for G in (16, 17, 19, 21, 23, 25):
    print('rowpair LDS.128 G=%d quarter-warps:' % G, run(G, 8, 8, 'cc'), ' planar:', run(G, 8, 8, 'pl'))
