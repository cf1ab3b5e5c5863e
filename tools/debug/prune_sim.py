"""Debug / analysis: how much work an EXACT prune of the difference sums could save on config-4-shaped data.
Rows of a 128 px cell play the part of the 128-pixel chunks. Default metric: CIE76 (numpy, fast); `--ciede` evaluates the real
CIEDE2000 per pixel with the f64 oracle (test infrastructure; ~15 s of CPU per cell)."""
import numpy as np, cv2, time, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mosaicmagnifique_b200 import synthetic
N=10000; S=128; K=145
CIEDE = "--ciede" in sys.argv
if CIEDE:
    from oracle import oracle
    oracle.build()
def cell_rows(cell):
    """[N][S] row sums of the per-pixel difference between `cell` and every library image"""
    if not CIEDE:
        return np.sqrt(((libL-cell[None])**2).sum(-1)).sum(2)
    out=np.empty((N,S))
    c=np.ascontiguousarray(np.broadcast_to(cell[None],(250,S,S,3))).reshape(-1,3)
    for i in range(0,N,250):
        out[i:i+250]=oracle.diff_batch(2,c,libL[i:i+250].reshape(-1,3)).reshape(250,S,S).sum(2)
    return out
lib=synthetic.make_library(N,S,1004)
main=synthetic.make_main_image(4320,7680,2004)
t=time.time()
libL=np.empty((N,S,S,3),np.float32)
for i in range(0,N,500):
    libL[i:i+500]=cv2.cvtColor((lib[i:i+500].reshape(-1,S,3).astype(np.float32)/255),cv2.COLOR_BGR2Lab).reshape(-1,S,S,3)
mainL=cv2.cvtColor(main.astype(np.float32)/255,cv2.COLOR_BGR2Lab)
print('prep',time.time()-t)
means=libL.mean((1,2))
# colour sort by morton code of quantised mean
q=np.clip(((means-means.min(0))/(np.ptp(means,0))*31).astype(np.int64),0,31)
def morton(q):
    code=np.zeros(len(q),np.int64)
    for b in range(5):
        for c in range(3):
            code|=((q[:,c]>>b)&1)<<(3*b+c)
    return code
order=np.argsort(morton(q),kind='stable')
rng=np.random.default_rng(0)
res=[]
for (cy,cx) in [(3,5),(10,20),(20,40),(30,55),(15,33),(8,48)]:
    cell=mainL[cy*S:(cy+1)*S,cx*S:(cx+1)*S]
    rows=cell_rows(cell)                         # N,S row sums
    pre=np.cumsum(rows,1)                        # prefix over rows (each row = 1 chunk of 128 px)
    full=pre[:,-1]
    T=np.sort(full)[K-1]
    # proxy threshold: mean colour distance candidates
    cm=cell.mean((0,1)); prox=np.sqrt(((means-cm)**2).sum(1))
    cand=np.argsort(prox)[:int(1.5*K)+8]
    Tp=np.sort(full[cand])[K-1]
    # per-pair prune chunk (first chunk index where prefix > T), check every 4 chunks
    def work(Tv, group):
        alive=pre<=Tv                      # N,S bool: still alive after chunk k
        # pair stops at first check point (multiple of 4) where not alive
        kstop=np.full(N,S)
        for k in range(3,S,4):
            dead=(~alive[:,k])&(kstop==S)
            kstop[dead]=k+1
        if group:
            ks=kstop[order].reshape(-1,8).max(1)   # warp = 8 colour-sorted images: stops when all dead
            return ks.sum()*8/(N*S)
        return kstop.sum()/(N*S)
    res.append((T/np.median(full), Tp/T, work(T,False), work(Tp,False), work(Tp,True), work(Tp*1.0,True)))
    print((cy,cx),'T/median %.3f Tproxy/T %.3f  work per-pair(T) %.3f per-pair(Tp) %.3f per-warp8 sorted(Tp) %.3f'%res[-1][:5])
    # random (unsorted) groups
    ks=np.full(N,S)
    alive=pre<=Tp
    for k in range(3,S,4):
        dead=(~alive[:,k])&(ks==S); ks[dead]=k+1
    print('   unsorted groups of 8: %.3f'%(ks.reshape(-1,8).max(1).sum()*8/(N*S)))

print("---- segment-pass pruning (math fraction incl. candidate completion), rows = (cell, 8 colour-sorted images)")
for S_ in (4, 8, 16):
    fr=[]
    for (cy,cx) in [(3,5),(10,20),(20,40),(30,55),(15,33),(8,48)]:
        cell=mainL[cy*S:(cy+1)*S,cx*S:(cx+1)*S]
        rows=cell_rows(cell); pre=np.cumsum(rows,1); full=pre[:,-1]
        seg=S//S_
        R=[pre[:,(s+1)*seg-1] for s in range(S_)]          # running totals after each segment
        M=int(1.25*K)+8
        cand=np.argsort(R[0])[:M]                            # proxy = first segment
        Tp=np.sort(full[cand])[K-1]
        tiles=order.reshape(-1,8)                           # lib tiles of 8 sorted images
        inv=np.empty(N,np.int64); inv[order]=np.arange(N)
        cand_tiles=np.unique(inv[cand]//8)
        work=np.zeros(len(tiles))                           # segments computed per (row=cell, tile)
        work[:]=1                                            # seg 0 for all
        complete=np.zeros(len(tiles),bool); complete[cand_tiles]=True
        work[cand_tiles]=S_
        alive=~complete
        for s in range(1,S_):
            a=(R[s-1][tiles]<=Tp).any(1)&alive
            work[a]+=1
            alive=a
        fr.append(work.sum()/(len(tiles)*S_))
    print('S=%d'%S_, ['%.3f'%f for f in fr], 'mean %.3f'%np.mean(fr))
