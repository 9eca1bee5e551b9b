import sys, numpy as np
sys.path.insert(0,'/root/repo')
from miluphcuda_b200 import scenarios
from scipy.spatial import cKDTree
name=sys.argv[1]; n=int(sys.argv[2]); variable=int(sys.argv[3])
sc=scenarios.make(name,n,stirred=False)
x=np.asarray(sc.x,dtype=np.float64).reshape(sc.n,-1)[:, :sc.dim]; h=np.asarray(sc.h,dtype=np.float64)
if h.ndim==0: h=np.full(sc.n,float(h))
n=sc.n
hmax=h.max(); hmin=h.min(); hmean=h.mean()
def morton3(c):
    k=np.zeros(len(c),dtype=np.int64)
    for b in range(20):
        for a in range(c.shape[1]):
            k|=((c[:,a]>>b)&1)<<(c.shape[1]*b+a)
    return k
def order_for(kind):
    cell = min(hmax, max(hmin, 0.5*hmean))*1.0001 if variable else hmax*1.0001/2
    if kind=='fine': cell=cell/8
    lo=x.min(0)-1e-9
    c=np.floor((x-lo)/cell).astype(np.int64)
    nc=c.max(0)+1
    if kind=='raster':
        key=c[:,0].copy()
        if sc.dim>1: key+=nc[0]*c[:,1]
        if sc.dim>2: key+=nc[0]*nc[1]*c[:,2]
    else:
        key=morton3(c)
    return np.argsort(key,kind='stable')
def run(kind):
    order=order_for(kind)
    xs=x[order]; hs=h[order]
    tree=cKDTree(xs)
    lists=tree.query_ball_point(xs, hs, return_sorted=True)
    lists=[np.array([j for j in l if j!=i],dtype=np.int64) for i,l in enumerate(lists)]
    noi=np.array([len(l) for l in lists])
    rng=np.random.default_rng(0)
    nw=n//32
    sel=rng.choice(nw, size=min(nw,400), replace=False)
    out=[]
    for G in (1,4,8):
        ppw=32//G
        groups=[g for w in sel for g in np.arange(w*32, w*32+32).reshape(-1,ppw)]
        tot_req=tot_lines=tot_sect=slots=0
        for w in groups:
            L=[lists[p] for p in w]
            iters=max((len(l)+G-1)//G for l in L)
            for q in range(iters):
                js=np.concatenate([l[q*G:(q+1)*G] for l in L])
                if len(js)==0: continue
                tot_req+=1; tot_sect+=len(np.unique(js)); tot_lines+=len(np.unique(js//4))
            slots+=iters*32
        pairs=noi[np.concatenate(groups)].sum()
        out.append("G%d lines/pair %.3f sect/pair %.3f util %.2f"%(G,tot_lines/pairs,tot_sect/pairs,pairs/slots))
    print(name,kind,"noi %.1f |"%noi.mean()," | ".join(out))
for kind in ('raster','morton','fine'):
    run(kind)
