"""Would a channel-planar column-sum buffer make phase H of pyramid_pair_kernel conflict-free? Simulates, over the real
tap tables, (a) the phase-V stores as STS.64 into plane[c][pos(px)] and (b) the phase-H LDS.64 with lanes = consecutive
output columns of ONE channel. pos(px) = px + px // Q (one padding slot per Q pixels), plane c starts at c * PS."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysilent_b200.util.zoom.from_image import PyramidPlan
plan = PyramidPlan((1080, 1920, 3), torch.uint8, 3, (288, 192), 2 ** .5)
L, w = plan.levels, plan.w
FC = 3; TW = 72; NT = 128

def wf64(addrs):   # wavefronts of one 64-bit shared access: two half-warps, each max multiplicity over 16 bank pairs
    tot = 0
    for half in (addrs[:16], addrs[16:]):
        a = [x % 16 for x in half if x is not None]
        tot += max(np.bincount(a, minlength=16)) if a else 0
    return tot

def run(Q, PS_mod):
    out = []
    for s in range(L):
        iy, wy, ix, wx = plan.level_tables(s)
        ok = ix[:, 0] >= 0
        h_wf = h_n = v_wf = v_n = 0
        for bx in range((w + TW - 1) // TW):
            cols = [ox for ox in range(bx * TW, min(w, (bx + 1) * TW)) if ok[ox]]
            if not cols: continue
            lo = min(ix[ox].min() for ox in cols); hi = max(ix[ox].max() for ox in cols)
            b0 = (lo * FC // 16) * 16; b1 = -(-((hi + 1) * FC) // 16) * 16
            nq = (b1 - b0) // 16
            px0 = b0 // 3
            npx = (b1 - 1) // 3 - px0 + 1
            pos = lambda px: (px - px0) + ((px - px0) // Q if Q else 0)
            plane_len = pos(px0 + npx - 1) + 1
            PS = plane_len + ((PS_mod - plane_len) % 16)          # plane stride with the requested residue
            # phase H: a warp = 32 consecutive ox of one channel
            for c in range(3):
                for g in range(0, TW, 32):
                    oxs = [bx * TW + g + l for l in range(32)]
                    for i in range(6):
                        addrs = [c * PS + pos(ix[ox][i]) if (ox < min(w, (bx + 1) * TW) and ox - bx * TW < TW and ok[ox]) else None for ox in oxs]
                        addrs = [a if (o - bx * TW) < TW else None for a, o in zip(addrs, oxs)]
                        if all(a is None for a in addrs): continue
                        h_wf += wf64(addrs); h_n += (any(a is not None for a in addrs[:16]) + any(a is not None for a in addrs[16:]))
            # phase V: lanes = consecutive groups of one row (row-crossing ignored), 16 STS.64 per task
            for g in range(0, nq, 32):
                for k in range(16):
                    addrs = []
                    for l in range(32):
                        qi = g + l
                        if qi >= nq: addrs.append(None); continue
                        B = b0 + 16 * qi + k
                        addrs.append((B % 3) * PS + pos(B // 3))
                    v_wf += wf64(addrs); v_n += (any(a is not None for a in addrs[:16]) + any(a is not None for a in addrs[16:]))
        out.append((h_wf / h_n * 2, v_wf / v_n * 2))
    return out

for Q in (0, 8, 16, 32):
    for PS_mod in (0, 1, 5, 6, 11):
        r = run(Q, PS_mod)
        print("Q=%2d PSmod=%2d  H:" % (Q, PS_mod), " ".join("%.2f" % a for a, b in r), "| V:", " ".join("%.2f" % b for a, b in r),
              "| mean H %.2f V %.2f" % (np.mean([a for a, b in r]), np.mean([b for a, b in r])))
