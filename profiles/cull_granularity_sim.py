"""Analysis behind DESIGN.md section 5 ("finer cull granularity"): uniform points at the bench's O density (32 768 in a 99.328 A box, 9^3 cells,
cutoff 10 A). The reference points of a home cell are split into G spatial chunks (recursive median split along the longest axis), each with its own
bounding box; a candidate is listed for a chunk when its gap to that box is within the cutoff. Prints, per home cell: executed pair tests (reference
points padded to 2, candidates to 64, as the kernel does), list entries, hits, and box tests of a non-hierarchical cull. Run: python profiles/cull_granularity_sim.py"""
import numpy as np
rng=np.random.default_rng(3)
L=99.328; nc=9; a=L/nc; r=10.0; N=32768
P=rng.random((N,3))*L
cell=(P//a).astype(int)
def cellpts(c): return P[np.all(cell==np.array(c),axis=1)]
res={}
for h in range(8):
    hc=np.array([2+h%2,3+(h//2)%2,4+h//4])
    R=cellpts(hc)
    T=np.concatenate([cellpts(hc+np.array(o)) for o in [(dx,dy,dz) for dx in(-1,0,1) for dy in(-1,0,1) for dz in(-1,0,1)]])
    hits=(((T[:,None,:]-R[None,:,:])**2).sum(2)<=r*r).sum()
    for G in (1,2,4,8,16):
        # split refs: recursive median split along the longest axis
        chunks=[R]
        while len(chunks)<G:
            new=[]
            for c in chunks:
                ext=c.max(0)-c.min(0); ax=np.argmax(ext); o=np.argsort(c[:,ax]); m=len(c)//2
                new+= [c[o[:m]], c[o[m:]]]
            chunks=new
        tests=0; kept=0
        for c in chunks:
            if len(c)==0: continue
            lo=c.min(0); hi=c.max(0)
            gap=np.maximum(np.maximum(lo-T,T-hi),0); keep=((gap**2).sum(1)<=r*r).sum()
            # pad: refs to multiple of 2, candidates to multiples of 64
            tests+= (-(-len(c)//2)*2) * (-(-keep//64)*64); kept+=keep
        d=res.setdefault(G,[0,0,0,0]); d[0]+=tests; d[1]+=kept; d[2]+=hits; d[3]+=len(T)*G
for G,(t,k,hh,boxtests) in res.items(): print(G,'tests',t/8,'list entries',k/8,'hits',hh/8,'tests/hit',t/hh,'box tests',boxtests/8)
