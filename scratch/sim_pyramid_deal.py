"""Phase H of pyramid_pair_kernel: can a host-built permutation of the 32 items of a warp between its two half-warps
remove the LDS.64 bank conflicts? (a 64-bit shared load is served one half-warp per wavefront; a half-warp is
conflict-free when its 16 lanes hit 16 different 8-byte bank pairs.) Prints wavefronts per LDS.64 for the current lane
order and for the best split found, per level."""
import numpy as np, sys, os, itertools, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysilent_b200.util.zoom.from_image import PyramidPlan
shape = tuple(int(v) for v in os.environ.get("SHAPE", "1080,1920").split(","))
plan = PyramidPlan(shape + (3,), torch.uint8, 3, (288, 192), 2 ** .5)
L, w = plan.levels, plan.w
FC = 3; TW = int(os.environ.get("TW", "72")); NT = 128; VG = int(os.environ.get("VG", "18"))

def cost(items):   # items: list of 6-vectors of f2 addresses; wavefronts summed over the 6 taps for one half-warp
    if not items: return 0
    a = np.array(items) % 16
    return sum(np.bincount(a[:, i], minlength=16).max() for i in range(6))

tot_cur = tot_new = tot_ideal = 0
for s in range(L):
    iy, wy, ix, wx = plan.level_tables(s)
    ok = ix[:, 0] >= 0
    cur = new = ideal = 0
    for bx in range((w + TW - 1) // TW):
        cols = [ox for ox in range(bx * TW, min(w, (bx + 1) * TW)) if ok[ox]]
        if not cols: continue
        lo = min(ix[ox].min() for ox in cols); wlo4 = (lo * FC // 16) * 16
        items = []
        for it in range(3 * TW):
            c = it % 3; ox = bx * TW + it // 3
            if ox >= w: break
            if ok[ox]:
                b = ix[ox] * FC - wlo4 + c
                items.append(list(b + (VG - 16) * (b >> 4)))
            else:
                items.append([0] * 6)
        for g in range(0, len(items), 32):
            grp = items[g:g + 32]
            cur += cost(grp[:16]) + cost(grp[16:])
            ideal += 6 * (1 + (len(grp) > 16))
            # greedy split: by tap-0 residue alternate halves, then pairwise swap improvement
            order = sorted(range(len(grp)), key=lambda i: grp[i][0] % 16)
            A = [grp[i] for i in order[0::2]]; B = [grp[i] for i in order[1::2]]
            if len(grp) <= 16: A, B = grp, []
            best = cost(A) + cost(B)
            improved = True
            while improved and B:
                improved = False
                for i in range(len(A)):
                    for j in range(len(B)):
                        A[i], B[j] = B[j], A[i]
                        c2 = cost(A) + cost(B)
                        if c2 < best: best = c2; improved = True
                        else: A[i], B[j] = B[j], A[i]
            new += best
    print("level %d: current %.2f wavefronts/LDS.64, dealt %.2f (ideal 2.00 -> %.2f)" % (s, 2 * cur / ideal, 2 * new / ideal, 2.0))
    tot_cur += cur; tot_new += new; tot_ideal += ideal
print("all levels: current %.2f, dealt %.2f" % (2 * tot_cur / tot_ideal, 2 * tot_new / tot_ideal))
