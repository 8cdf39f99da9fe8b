import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import gp_oracle as go
import scipy.linalg as sla
def digits(q):
    out=[]
    q=q.copy()
    for s in range(7):
        b=q&255
        d=np.where(b>=128,b-256,b)
        out.append(d)
        q=(q+128)>>8
    assert np.all(q==0), (q.min(), q.max())
    return out[::-1]   # top first
def scale_exp(m):
    f,e=np.frexp(m)
    return np.where(f<0.996,e+1,e+2)
def run(n,d,sn2,ls,seed=0,sf2=1.3):
    rng=np.random.default_rng(seed)
    x=rng.uniform(size=(n,d)); xs=rng.uniform(size=(64,d))
    ls2=np.full(d,ls) if ls else 0.5*(1+np.arange(d)/d)
    p=go.GPParams(sf2,ls2,sn2,None)
    K=go.kernel(p,x,x)+sn2*np.eye(n)
    L=np.linalg.cholesky(K); Linv=sla.solve_triangular(L,np.eye(n),lower=True)
    Ks=go.kernel(p,xs,x)
    Wref=(Ks.astype(np.longdouble)@Linv.T.astype(np.longdouble))
    W64=Ks@Linv.T
    ea=int(scale_exp(np.array(sf2)))
    qa=np.rint(Ks*2.0**(56-ea)).astype(np.int64)
    A=digits(qa)
    m=np.abs(Linv).max(axis=1); eb=scale_exp(m)
    qb=np.rint(Linv*(2.0**(56-eb))[:,None]).astype(np.int64)
    B=digits(qb)
    assert all(np.abs(a).max()<=128 for a in A+B)
    G=[np.zeros((64,n),np.int64) for _ in range(7)]
    for s in range(7):
        for t in range(7-s):
            G[s+t]+=A[s]@B[t].T
    assert max(np.abs(g).max() for g in G)<2**31
    hi=(G[0]<<16)+(G[1]<<8)+G[2]; lo=(G[3]<<24)+(G[4]<<16)+(G[5]<<8)+G[6]
    W=(2.0**(ea+eb-32))[None,:]*(hi.astype(np.float64)+lo.astype(np.float64)*2.0**-32)
    e_i8=np.abs(W-Wref).max(); e_64=np.abs(W64-Wref).max()
    s2ref=(Wref**2).sum(1); s2_i8=(W**2).sum(1); s2_64=(W64**2).sum(1)
    print(f"n={n} sn2={sn2}: maxLinv {np.abs(Linv).max():.3g}  errW i8 {float(e_i8):.3g} fp64 {float(e_64):.3g}; err sumW2 i8 {float(np.abs(s2_i8-s2ref).max()):.3g} fp64 {float(np.abs(s2_64-s2ref).max()):.3g}")
run(256,6,1e-3,None); run(512,6,1e-8,0.05); run(384,20,1e-6,None); run(300,4,1e-8,0.5,sf2=0.9999999); run(1000,20,1e-3,None)
