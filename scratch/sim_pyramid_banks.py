import numpy as np, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from pysilent_b200.util.zoom.from_image import PyramidPlan
# tap tables from the library's own plan (a geometry-only plan needs no GPU)
plan = PyramidPlan((1080, 1920, 3), torch.uint8, 3, (288, 192), 2 ** .5)
L, w = plan.levels, plan.w
t = {'ix': [], 'okx': []}
for s_ in range(L):
    iy, wy, ix, wx = plan.level_tables(s_)
    t['ix'].append(ix)
    t['okx'].append(ix[:, 0] >= 0)
FC=3; TW=72; NT=256
def wavefronts(addrs):  # addrs: f2 indices of up to 32 lanes (None = inactive); LDS.64 -> two half-warps
    tot=0
    for half in (addrs[:16], addrs[16:]):
        banks={}
        for a in half:
            if a is None: continue
            banks.setdefault((2*a)%32, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot
for s in range(L):
    ix=t['ix'][s]; ok=t['okx'][s]
    res={}
    for vg in (16,18,20,22,24,26,28,30,32):
        total=0; ideal=0
        for bx in range((w+TW-1)//TW):
            cols=[ox for ox in range(bx*TW,min(w,(bx+1)*TW)) if ok[ox]]
            if not cols: continue
            lo=min(ix[ox].min() for ox in cols); byte_lo=lo*FC; wlo4=(byte_lo//16)*16
            for wi in range(0,NT,32):
                for i in range(6):
                    addrs=[]
                    for l in range(32):
                        tid=wi+l; c=tid%3; col=tid//3; ox=bx*TW+col
                        if tid>=3*TW or ox>=w: addrs.append(None); continue
                        b = ix[ox][i]*FC - wlo4 + c if ok[ox] else 0
                        addrs.append(b + (vg-16)*(b>>4))
                    wf=wavefronts(addrs); total+=wf
                    ideal += (1 if any(a is not None for a in addrs[:16]) else 0)+(1 if any(a is not None for a in addrs[16:]) else 0)
        res[vg]=(total,ideal)
    print(s, {k:round(v[0]/v[1],2) for k,v in res.items()})
