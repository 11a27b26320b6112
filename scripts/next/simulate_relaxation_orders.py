import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, helpers
from helpers import Oracle
from swarm_b200 import HostDb
n=int(sys.argv[1]) if len(sys.argv)>1 else 300000
fa=f"/tmp/sim_{n}.fa"
helpers.make_fasta(fa, n, 150, 42)
db=HostDb(fa); orc=Oracle(db); t=time.time(); orc.network(); print("oracle network s", round(time.time()-t,1))
L=orc.links(); src=L[:,0].astype(np.int64); dst=L[:,1].astype(np.int64); m=len(src)
print("n",n,"m",m, "back-links", int((src>dst).sum()))
INF=np.iinfo(np.int64).max
def jacobi_rounds(order_ranges):
    """order_ranges: list of index arrays (sub-steps per round). keys updated after each sub-step."""
    key=(np.arange(n,dtype=np.int64)<<32)
    lowered=np.ones(n,bool)   # round 0: everything active
    rounds=0; offers=0; succ=0
    while True:
        new_low=np.zeros(n,bool)
        late=np.zeros(n,bool)
        for sub,idx in enumerate(order_ranges):
            act=idx[lowered[src[idx]] | new_low_sub_mask(new_low, src[idx], sub)] if False else idx[lowered[src[idx]] | new_low[src[idx]]]
            offers+=len(act)
            cand=key[src[act]]+1
            better=cand<key[dst[act]]
            a=act[better]; c=cand[better]
            if len(a):
                # scatter-min
                o=np.lexsort((c,dst[a])); da=dst[a][o]; ca=c[o]
                first=np.concatenate(([True],da[1:]!=da[:-1]))
                key[da[first]]=np.minimum(key[da[first]],ca[first])
                new_low[da[first]]=True
                succ+=int(first.sum())
        rounds+=1
        if not new_low.any(): break
        lowered=new_low
    return rounds, offers, succ, key
def new_low_sub_mask(*a): return None
allidx=np.arange(m)
r,o,s,k0=jacobi_rounds([allidx])
print("flat rounds",r,"offers/link",round(o/m,2),"lowerings/vertex",round(s/n,2))
for R in (4,16,64):
    rng=(src*R//n)
    parts=[allidx[rng==b] for b in range(R)]
    r,o,s,k=jacobi_rounds(parts)
    assert np.array_equal(k,k0)
    print("sweep R",R,"rounds",r,"offers/link",round(o/m,2),"lowerings/vertex",round(s/n,2))
