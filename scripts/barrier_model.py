"""Randomized model of conv_tc3's mbarrier protocol (all roles: raw-halo TMA producer, transform group, weight TMA producer, one or
two MMA issuers with asynchronous commits, two epilogue groups): picks a runnable role / pending asynchronous arrival at random
and reports schedules that deadlock (a waiter lapped by its barrier, a wrong arrival count, a stage ring out of step).
usage: python scripts/barrier_model.py      (no GPU; prints deadlock counts per shape, expected 0 everywhere)"""
import random, sys
class Bar:
    def __init__(s,count): s.count=count; s.pending=count; s.phase=0; s.completions=0
    def arrive(s):
        s.pending-=1
        assert s.pending>=0
        if s.pending==0: s.phase^=1; s.pending=s.count; s.completions+=1
    def ready(s,parity): return s.phase!=parity

def run(seed, T=5, kblocks=(9,9), SA=3, SB=3, dual=True, verbose=False):
    rnd=random.Random(seed)
    NI=2 if dual else 1
    a_full=[Bar(1) for _ in range(SA)]; a_empty=[Bar(NI) for _ in range(SA)]; raw=[Bar(1) for _ in range(SA)]
    b_full=[Bar(1) for _ in range(SB)]; b_empty=[Bar(1) for _ in range(SB)]
    acc_full=[Bar(NI) for _ in range(2)]; acc_empty=[Bar(1) for _ in range(2)]
    asyncq=[]   # list of per-source FIFOs of pending arrivals: (bar)
    fifos={'tma_raw':[], 'tma_b':[], 'mma0':[], 'mma1':[]}
    log=[]
    def W(bar,par): return ('wait',bar,par)
    def rawprod():
        s=0;ph=0
        for t in range(T):
            for it in range(len(kblocks)):
                yield W(a_empty[s],ph^1)
                fifos['tma_raw'].append(raw[s])
                s+=1
                if s==SA: s=0;ph^=1
    def transform():
        s=0;ph=0
        for t in range(T):
            for it in range(len(kblocks)):
                yield W(raw[s],ph)
                a_full[s].arrive()
                s+=1
                if s==SA: s=0;ph^=1
    def bprod():
        s=0;ph=0
        for t in range(T):
            for it,nt in enumerate(kblocks):
                for tap in range(nt):
                    yield W(b_empty[s],ph^1)
                    fifos['tma_b'].append(b_full[s])
                    s+=1
                    if s==SB: s=0;ph^=1
    def issuer(wi):
        sa=0;sb=0;pha=0;phb=0;g=0
        q=fifos['mma%d'%wi]
        for tl in range(T):
            ab=tl&1
            yield W(acc_empty[ab],((tl>>1)&1)^1)
            for it,nt in enumerate(kblocks):
                yield W(a_full[sa],pha)
                for tap in range(nt):
                    if dual and (g&1)!=wi:
                        sb+=1
                        if sb==SB: sb=0;phb^=1
                        g+=1
                        continue
                    yield W(b_full[sb],phb)
                    q.append(b_empty[sb])      # commit after MMAs
                    sb+=1
                    if sb==SB: sb=0;phb^=1
                    g+=1
                q.append(a_empty[sa])
                sa+=1
                if sa==SA: sa=0;pha^=1
            q.append(acc_full[ab])
    def epi(eg):
        for tl in range(eg,T,2):
            ab=tl&1
            yield W(acc_full[ab],(tl>>1)&1)
            yield ('work',)
            acc_empty[ab].arrive()
    procs={'rawprod':rawprod(),'xform':transform(),'bprod':bprod(),'iss0':issuer(0),'epi0':epi(0),'epi1':epi(1)}
    if dual: procs['iss1']=issuer(1)
    pending={k:None for k in procs}
    done=set()
    steps=0
    while True:
        steps+=1
        runnable=[]
        for k,gen in procs.items():
            if k in done: continue
            p=pending[k]
            if p is None or p[0]=='work' or (p[0]=='wait' and p[1].ready(p[2])): runnable.append(('proc',k))
        for k,f in fifos.items():
            if f: runnable.append(('async',k))
        if not runnable:
            if len(done)==len(procs): return True
            if verbose:
                for k in procs:
                    if k not in done: print('blocked',k,pending[k][0], [n for n,lst in [('a_full',a_full),('a_empty',a_empty),('raw',raw),('b_full',b_full),('b_empty',b_empty),('acc_full',acc_full),('acc_empty',acc_empty)] if pending[k][1] in lst], pending[k][2] if pending[k][0]=='wait' else '')
            return False
        kind,k=rnd.choice(runnable)
        if kind=='async':
            fifos[k].pop(0).arrive()
        else:
            try:
                pending[k]=next(procs[k])
            except StopIteration:
                done.add(k); pending[k]=None
        if steps>10_000_000: return False

if __name__=='__main__':
    bad=0
    for cfg in [dict(kblocks=(9,),SA=3,SB=3),dict(kblocks=(9,9),SA=3,SB=3),dict(kblocks=(9,1),SA=3,SB=3),dict(kblocks=(9,1,1,1),SA=3,SB=3),dict(kblocks=(9,9,9),SA=2,SB=5),dict(kblocks=(9,),SA=2,SB=5), dict(kblocks=(9,1,1),SA=2,SB=5)]:
        for dual in (False,True):
            fails=0
            for seed in range(300):
                if not run(seed,T=6,dual=dual,**cfg): fails+=1
            print(cfg,'dual',dual,'deadlocks',fails)
