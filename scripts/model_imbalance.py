"""List-scheduling model of the load balance of the Gaussian-major kernels (CPU only): per-warp work of the synthetic
   init scene in Morton order, CTAs scheduled in index order / heaviest first onto the resident slots.  python scripts/model_imbalance.py"""
import numpy as np, sys, heapq
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgegaussians_b200 import synth
N,W,H=500_000,1600,1200
m,q,s,o=synth.make_gaussians(N,'init',0)
vms,Ks=synth.make_cameras(8,W,H)
vm,K=vms[0],Ks[0]
# Morton order like the model (10 bits per axis over bbox)
def morton(m,bits=10):
    lo,hi=m.min(0),m.max(0)
    qv=np.clip(((m-lo)/(hi-lo+1e-12)*(2**bits-1)).astype(np.int64),0,2**bits-1)
    code=np.zeros(len(m),np.int64)
    for b in range(bits):
        for a in range(3):
            code|=((qv[:,a]>>b)&1)<<(3*b+a)
    return np.argsort(code,kind='stable')
perm=morton(m)
m=m[perm]; s=s[perm]; o=o[perm]
R,t=vm[:3,:3],vm[:3,3]
pc=m@R.T+t
z=pc[:,2]
fx=K[0,0]
scale=np.exp(s).max(1) if s.min()<0 else s.max(1)
# opacity
op=1/(1+np.exp(-o)) if o.min()<0 else o
sig=scale*fx/z   # pixel sigma (isotropic approx)
sig2=sig**2+0.3
tau=np.log(255*np.clip(op.reshape(-1),1e-9,1))
rad=np.sqrt(np.maximum(2*tau*sig2,0))
x=fx*pc[:,0]/z+K[0,2]; y=K[1,1]*pc[:,1]/z+K[1,2]
vis=(z>0.01)&(x>-rad)&(x<W+rad)&(y>-rad)&(y<H+rad)
area=np.where(vis,np.pi*rad**2,0)
print('visible',vis.mean(),'mean px',area.mean(), 'z range',z.min(),z.max())
rows=np.where(vis,2*rad+1,0)
# per-warp work model for lane=Gaussian path: sum over rows of max over lanes... approximate: warp time = max over lanes (rows*chunks) 
chunks_per_row=np.ceil((2*rad+4)/4)
lane_work=rows*chunks_per_row
for unit in (128,64,32):
    nb=(N+unit-1)//unit
    pad=np.zeros(nb*unit); pad[:N]=lane_work
    wl=pad.reshape(-1,32).max(1)           # per-warp duration ~ max lane work
    cta=wl.reshape(nb,-1).max(1)          # CTA duration = slowest warp (all warps run concurrently)
    tot_slots=148*6*(128//unit)
    # list scheduling in index order
    h=[0.0]*tot_slots; heapq.heapify(h)
    for d in cta:
        t0=heapq.heappop(h); heapq.heappush(h,t0+d)
    mk=max(h); ideal=cta.sum()/tot_slots
    # heavy-first
    h=[0.0]*tot_slots; heapq.heapify(h)
    for d in np.sort(cta)[::-1]:
        t0=heapq.heappop(h); heapq.heappush(h,t0+d)
    mk2=max(h)
    print(f'unit {unit}: CTAs {nb} mean {cta.mean():.1f} max {cta.max():.1f} p99 {np.percentile(cta,99):.1f}; makespan/ideal in-order {mk/ideal:.3f} heavy-first {mk2/ideal:.3f}; lane-util {lane_work.sum()/ (wl.sum()*32):.3f}')
