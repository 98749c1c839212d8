"""Phase H deal over groups of G items (G = 32 ... whole tile): LDS.64 wavefronts per load after a greedy residue-balanced
assignment of items to half-warps, and the 128-byte lines one warp's STG.64 then touches (3 planes x 576 B per tile row)."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysilent_b200.util.zoom.from_image import PyramidPlan
plan = PyramidPlan((1080, 1920, 3), torch.uint8, 3, (288, 192), 2 ** .5)
L, w = plan.levels, plan.w
FC = 3; TW = 72; VG = 18

def half_cost(vecs):
    if not vecs: return 0
    a = np.array(vecs) % 16
    return sum(np.bincount(a[:, i], minlength=16).max() for i in range(6))

def soft_cost(vecs):   # every collision counts (sum of squared multiplicities): the greedy's objective
    if not vecs: return 0
    a = np.array(vecs) % 16
    return sum(int((np.bincount(a[:, i], minlength=16) ** 2).sum()) for i in range(6))

def deal(items, G):
    """items: list of (id, 6-vector). returns list of warps (each a list of <=32 items) after dealing inside groups of G."""
    warps = []
    for g in range(0, len(items), G):
        grp = items[g:g + G]
        nh = -(-len(grp) // 16)
        halves = [[] for _ in range(nh)]
        # greedy: items sorted by tap-0 residue, each goes to the non-full half where it adds the least cost
        for it in sorted(grp, key=lambda x: x[1][0] % 16):
            best, bc = None, None
            for hi, hf in enumerate(halves):
                if len(hf) >= 16: continue
                c = soft_cost([v for _, v in hf] + [it[1]]) - soft_cost([v for _, v in hf])
                if bc is None or c < bc or (c == bc and len(hf) < len(halves[best])): best, bc = hi, c
            halves[best].append(it)
        # pair halves into warps: keep order
        for i in range(0, nh, 2):
            warps.append(halves[i] + (halves[i + 1] if i + 1 < nh else []))
    return warps

for G in (32, 64, 96, 216):
    tot_wf = tot_n = tot_lines = tot_warps = 0
    per_level = []
    for s in range(L):
        iy, wy, ix, wx = plan.level_tables(s)
        ok = ix[:, 0] >= 0
        wf = n = lines = nw = 0
        for bx in range((w + TW - 1) // TW):
            cols = [ox for ox in range(bx * TW, min(w, (bx + 1) * TW)) if ok[ox]]
            if not cols: continue
            lo = min(ix[ox].min() for ox in cols); wlo4 = (lo * FC // 16) * 16
            items = []
            for it in range(3 * TW):
                c = it % 3; ox = bx * TW + it // 3
                if ox >= w: break
                b = ix[ox] * FC - wlo4 + c if ok[ox] else np.zeros(6, int)
                items.append((it, list(b + (VG - 16) * (b >> 4))))
            for warp in deal(items, G):
                wf += half_cost([v for _, v in warp[:16]]) + half_cost([v for _, v in warp[16:]])
                n += 6 * (1 + (len(warp) > 16))
                # output address of item (ox_local, c): plane c, byte (bx*TW + ox_local) * 8 in a 288*8-byte row
                ln = set((it % 3, ((bx * TW + it // 3) * 8) // 128) for it, _ in warp)
                lines += len(ln); nw += 1
        per_level.append((2 * wf / n, lines / nw))
        tot_wf += wf; tot_n += n; tot_lines += lines; tot_warps += nw
    print("G=%3d  LDS wavefronts/instr %.2f  STG lines/warp %.1f  -> L1 wavefronts per warp-row %.1f | per level" % (
        G, 2 * tot_wf / tot_n, tot_lines / tot_warps, 6 * 2 * tot_wf / tot_n + tot_lines / tot_warps),
        " ".join("%.2f/%.1f" % p for p in per_level))
