"""How early can the rejection of a slice candidate be proven?  (CPU oracle; the bound of csrc/lna_async.cuh)  usage: python tools/early_exit_probe.py"""
import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from conftest import Golden
from oracle import oracle as orc
g=Golden("changepoint_sim")
par=g.final["par"].ravel().copy(); eps=g.final["OriginTraj"].copy()
ini=g.init; C,D,y,gridrep=ini["C"],ini["D"],ini["y"],ini["gridrep"].astype(int); ng=int(np.ravel(ini["ng"])[0]); n0=ng-1
start=np.concatenate([[0],np.cumsum(gridrep)])
cellCD=np.array([np.sum(C[start[i]:start[i+1]]*D[start[i]:start[i+1]]) for i in range(n0)])
cellY=np.array([np.sum(y[start[i]:start[i+1]]) for i in range(n0)])
# per-cell upper bound: max over lam of -2 a lam + y log lam
ub=np.where(cellY>0, cellY*np.log(np.maximum(cellY,1e-300)/(2*cellCD))-cellY, 0.0)
PRIORS=[(1.0,1.0),(0.7,0.5),(-1.7,0.1),(3.0,0.2)]; ESS=[0,1,0,0,1,1,0]
def cells(parv):
    r=orc.New_Param_List(parv[2:],parv[:2],g.gridsize,g.times,g.x_r,g.x_i)
    lat=orc.TransformTraj(r["Ode"],eps,r["FT"]); bet=r["betaN"]
    L=np.argmax(lat[:,0]>=g.t_correct)
    t=np.zeros(n0)
    for i in range(n0):
        row=L-i; S,I=lat[row,1],lat[row,2]
        if S<=0 or I<=0: return None,L
        lam=bet[row]*S/I
        t[i]=-2*cellCD[i]*lam+cellY[i]*np.log(lam)
    return t,L
cur,L=cells(par); cl=cur.sum(); print("coalLog",cl, "n0",n0,"L",L)
rng=np.random.default_rng(0)
fr=[];acc_idx=[]
for rep in range(300):
    nrm=rng.standard_normal(43); u=rng.uniform(size=22)
    logy=cl+np.log(u[0])
    th=2*np.pi*u[1]; lo,hi=th-2*np.pi,th
    for j in range(1,21):
        newpar=orc.Param_Slice_update_all2(par,g.x_r,g.x_i,th,nrm,ESS,PRIORS)
        try: t,_=cells(newpar)
        except Exception: t=None
        if t is None: fr.append((j,0.0)); 
        else:
            ll=t.sum()
            if ll>logy: acc_idx.append(j); break
            # forward order = cells n0-1 down to 0; find earliest forward position where partial+bound<logy
            order=np.arange(n0-1,-1,-1)
            part=np.cumsum(t[order]); rem=np.concatenate([np.cumsum(ub[order][::-1])[::-1][1:],[0.0]])
            ok=np.nonzero(part+rem<logy-1e-6)[0]
            pos=(ok[0]+1)/n0 if len(ok) else 1.0
            # forward-time fraction: cell index order[pos] ↔ row L - i → fraction of rows
            row=(L-order[ok[0]]) if len(ok) else L
            fr.append((j,row/ (len(g.times)//g.gridsize)))
        if th<0: lo=th
        else: hi=th
        th=lo+(hi-lo)*u[j+1]
fr=np.array(fr)
print("mean accepted index",np.mean(acc_idx))
for j in range(1,13):
    m=fr[fr[:,0]==j,1]
    if len(m): print("cand",j,"n",len(m),"mean fraction of the trajectory needed to prove rejection",m.mean().round(3), "P(frac<0.5)",np.mean(m<0.5).round(2))
