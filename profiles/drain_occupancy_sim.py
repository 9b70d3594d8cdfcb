"""Analysis behind DESIGN.md section 5 ("drain balance by ordering the candidate lists"): the hit queue of k_rdf_pairs_v2 replayed on uniform points
(one column of 40 slots per lane, drain when a lane passes 32, rounds of 4 rows): fraction of drained slots that hold an entry for the list order the
cull produces today ('none'), for candidates bucketed by their gap to the home cell's bounding box ('bucket'), by distance to its centre ('center') and
for the exact order of hit probability ('exactp'). Run: python profiles/drain_occupancy_sim.py"""
import numpy as np
rng=np.random.default_rng(1)
L=99.328; nc=9; a=L/nc; r=10.0; N=32768
P=rng.random((N,3))*L
cell=(P//a).astype(int)
def cellpts(c): 
    m=np.all(cell==np.array(c)%nc,axis=1); return P[m]
def sim(order, nb=8, QCAP=40, QTRIG=32, homes=12):
    tot_useful=0; tot_slots=0; tests=0
    for h in range(homes):
        hc=np.array([2+h%3,3+(h//3)%3,4])
        R=cellpts(hc)
        lo=R.min(0); hi=R.max(0)
        # class 0: 13 'larger' neighbour cells (no wrap needed in the interior)
        T=[]
        offs=[(dx,dy,dz) for dx in(-1,0,1) for dy in(-1,0,1) for dz in(-1,0,1)]
        offs=[o for o in offs if (o[2],o[1],o[0])>(0,0,0)]
        for o in offs: T.append(cellpts(hc+np.array(o)))
        T=np.concatenate(T)
        gap=np.maximum(np.maximum(lo-T,T-hi),0); lb2=(gap**2).sum(1)
        keep=lb2<=r*r; T=T[keep]; lb2=lb2[keep]
        if order=='bucket':
            b=np.minimum((np.sqrt(lb2)/r*nb).astype(int),nb-1)
            idx=np.argsort(b,kind='stable'); T=T[idx]
        elif order=='center':
            c=(lo+hi)/2; dc=((T-c)**2).sum(1); 
            b=np.minimum(((np.sqrt(dc)-0)/ (r+a*0.87) *nb).astype(int),nb-1)
            idx=np.argsort(b,kind='stable'); T=T[idx]
        elif order=='exactp':
            d2=((T[:,None,:]-R[None,:,:])**2).sum(2); p=(d2<=r*r).mean(1); idx=np.argsort(p); T=T[idx]
        q=np.zeros(32,int)
        n=len(T)
        for j0 in range(0,n,128):
            ch=T[j0:j0+128]; m=len(ch)
            nslots=128 if m>64 else 64
            pad=np.full((nslots,3),1e9); pad[:m]=ch
            hits=((pad[:,None,:]-R[None,:,:])**2).sum(2)<=r*r   # [slot, ref]
            lanehits=hits.reshape(nslots//32,32,-1).sum(0)       # [lane, ref]
            nref=R.shape[0]
            for g in range(0,nref,2):
                q+=lanehits[:,g:g+2].sum(1)
                tests+=nslots*2
                if q.max()>QTRIG:
                    rows=-(-q.max()//4)*4
                    tot_useful+=q.sum(); tot_slots+=32*rows; q[:]=0
        rows=-(-q.max()//4)*4; tot_useful+=q.sum(); tot_slots+=32*rows
    return tot_useful/tot_slots, tot_useful/tests
for o in ['none','bucket','center','exactp']:
    for nb in ([8] if o in('none','exactp') else [4,8,16]):
        print(o,nb,sim(o,nb))
