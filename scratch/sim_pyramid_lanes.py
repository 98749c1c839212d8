import numpy as np, sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from pysilent_b200.util.zoom.from_image import PyramidPlan
plan = PyramidPlan((1080, 1920, 3), torch.uint8, 3, (288, 192), 2 ** .5)
L, w = plan.levels, plan.w
FC=3; TW=72; VG=18
def wf(addrs):
    tot=0
    for half in (addrs[:16], addrs[16:]):
        banks={}
        for a in half:
            if a is None: continue
            banks.setdefault(a%16, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot
T_old=T_new=0
for s in range(L):
    iy,wy,ix,wx = plan.level_tables(s)
    ok = ix[:,0] >= 0
    o_=n_=0
    for bx in range((w+TW-1)//TW):
        def off(col,c,i):
            ox=bx*TW+col
            if ox>=w or not ok[ox]: return None
            B=ix[ox][i]*FC+c
            return B+(VG-16)*(B>>4)
        # old: lanes = (col,c)
        items=[(col,c) for col in range(TW) for c in range(3)]
        for base in range(0,len(items),32):
            grp=items[base:base+32]
            for i in range(6):
                a=[off(col,c,i) for (col,c) in grp]+[None]*(32-len(grp)); o_+=wf(a)
        # new: lanes = columns, one channel per instruction
        for base in range(0,TW,32):
            cols=list(range(base,min(TW,base+32)))
            for i in range(6):
                for c in range(3):
                    a=[off(col,c,i) for col in cols]+[None]*(32-len(cols)); n_+=wf(a)
    print(s,'old wavefronts',o_,'new',n_)
    T_old+=o_; T_new+=n_
print('total', T_old, T_new, T_new/T_old)
