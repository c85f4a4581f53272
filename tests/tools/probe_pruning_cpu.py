"""CPU probe behind DESIGN.md section 8: how many (frame, component) exponentials of the config-4 scoring are negligible at
FP32 resolution, and how many 32 x 32 epilogue blocks become skippable after reordering components and frames.
Oracle features (checker code) on bench-style audio; numpy only.  python tests/tools/probe_pruning_cpu.py"""
import numpy as np, sys, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import frontend as ofe
from speech_signal_processing_b200 import synth
K,D=1024,39
w,mu,var=synth.synth_ubm(K,D,seed=0)
g=torch.Generator(); g.manual_seed(1000)
N=160; S=48000
tt=torch.arange(S,dtype=torch.float32)/16000.0
f0=80+170*torch.rand((N,1),generator=g)
sig=torch.zeros((N,S))
for h in range(1,12):
    sig+=torch.sin(2*np.pi*h*f0*tt[None])/h*torch.rand((N,1),generator=g)
sig+=0.3*torch.randn((N,S),generator=g)
sig*=3000.0/sig.pow(2).mean(dim=1,keepdim=True).sqrt()
pcm=sig.round().clamp(-32768,32767).to(torch.int16).numpy()
t=time.time()
feats=np.vstack([ofe.features(p,preset="sidekit",delta_order=2,cmvn=True) for p in pcm]).astype(np.float64)
print(feats.shape, time.time()-t)
LOG2E=1/np.log(2)
a=mu/var; b=-0.5/var
c=np.log(w)-0.5*(D*np.log(2*np.pi)+(mu*mu/var).sum(1))-0.5*np.log(var).sum(1)
L=(feats@a.T+(feats*feats)@b.T+c)*LOG2E
m=L.max(axis=1,keepdims=True)
d=L-m
best=d.argmax(axis=1)
print("significant per frame (-30):",(d>-30).sum(1).mean(),"(-54):",(d>-54).sum(1).mean())
def skipfrac(d, thr, fb=32, cb=32):
    T=(d.shape[0]//fb)*fb
    blk=d[:T].reshape(T//fb,fb,K//cb,cb).max(axis=(1,3))
    return (blk<thr).mean()
print("baseline order: skip(-30) %.3f skip(-54) %.3f"%(skipfrac(d,-30),skipfrac(d,-54)))
# reorder components by co-activation: spectral/1-D ordering via leading eigenvector of co-significance
sigm=(d>-40).astype(np.float32)
co=sigm.T@sigm
# hierarchical: order by recursive bisection with leading eigenvector of normalized co-activation
def order(idx):
    if len(idx)<=32: return list(idx)
    sub=co[np.ix_(idx,idx)]; dg=np.sqrt(np.maximum(sub.diagonal(),1)); nm=sub/dg[:,None]/dg[None,:]
    nm=nm-nm.mean(0,keepdims=True)
    vals,vecs=np.linalg.eigh(nm@nm.T if False else (nm+nm.T)/2)
    v=vecs[:,-1]; o=np.argsort(v); h=len(idx)//2
    return order(idx[o[:h]])+order(idx[o[h:]])
corder=np.array(order(np.arange(K)))
d2=d[:,corder]
best2=d2.argmax(axis=1)
forder=np.argsort(best2,kind='stable')
d3=d2[forder]
for thr in (-30,-40,-54):
    print("reordered comps+frames: thr",thr,"32x32 skip %.3f"%skipfrac(d3,thr),"32x64 skip %.3f"%skipfrac(d3,thr,32,64), "128x64 %.3f"%skipfrac(d3,thr,128,64))
# frames sorted only
d4=d[np.argsort(best,kind='stable')]
print("frames sorted only: 32x32 skip(-30) %.3f (-54) %.3f"%(skipfrac(d4,-30),skipfrac(d4,-54)))
