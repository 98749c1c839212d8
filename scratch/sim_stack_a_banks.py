import itertools
TRIPLES, RUNS, CH = 6, 7, 3
def cost(pitch, rows, order):
    Pp = pitch//2; plane = rows*Pp
    # tasks enumerated with 'order' = tuple naming fastest..slowest among ('rt','c','k')
    dims = {'rt':TRIPLES,'c':CH,'k':RUNS}
    tasks=[]
    for t in range(TRIPLES*RUNS*CH):
        idx={}; r=t
        for name in order:
            idx[name]=r%dims[name]; r//=dims[name]
        tasks.append((idx['rt'],idx['c'],idx['k']))
    tot=0; ideal=0
    for wbase in range(0,len(tasks),32):
        warp=tasks[wbase:wbase+32]
        for i in range(5):
            for q in range(5):
                for qb in range(0,len(warp),8):
                    lanes=warp[qb:qb+8]
                    units={}
                    for (rt,c,k) in lanes:
                        a=c*plane+(3*rt+i)*Pp+4*k+q
                        units.setdefault(a%8,set()).add(a)
                    tot+=max(len(v) for v in units.values()); ideal+=1
    return tot/ideal
best=[]
for pitch in range(58,72,2):
    for rows in (20,21,22):
        for order in itertools.permutations(('rt','c','k')):
            best.append((cost(pitch,rows,order),pitch,rows,order))
best.sort()
for b in best[:12]: print(b)
print('current', cost(58,20,('rt','c','k')), cost(58,20,('rt','k','c')))
