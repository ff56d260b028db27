import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from learning_embeddings_b200 import sharding, ops, hierarchy
from learning_embeddings_b200.engine import ConeStep, pack_index_block
h = hierarchy.ethec(); D=10; world=2; Nn=5; B=1024
rng=np.random.default_rng(0); edges=h.closure_edges()
sel=rng.integers(0,len(edges),size=B); u,v=edges[sel,0],edges[sel,1]; nt,nf=h.sample_negatives(u,v,Nn,rng)
g=torch.Generator().manual_seed(0); w=torch.randn(h.n,D,generator=g); W0=(0.1+0.05*torch.rand(h.n,1,generator=g))*w/w.norm(dim=1,keepdim=True)
xs=sharding.LocalExchange.make(world,h.n,12,torch.device('cuda'),timeout_ms=500,mode=1)
engs=[ConeStep(W0.cuda().clone(),'hyp',Nn,B,K=0.1,alpha=0.05,lr=0.01,exchange=xs[r]) for r in range(world)]
parts=[]
for r in range(world):
    lo,hi=sharding.shard_bounds(B,r,world); parts.append((pack_index_block(u[lo:hi],v[lo:hi],nt[lo:hi],nf[lo:hi]).cuda(),hi-lo))
def flags(r):
    b=xs[r].bufs[r].view(torch.int32); rsf=2*(2*384*12)*2  # floats before flags: 2*(rs+ag)
    return b[rsf:rsf+48].cpu().numpy().reshape(4,12)
for step in range(2):
    for r in range(world): engs[r].forward_backward(*engs[r]._split(*parts[r]))
    for ph in (1,2,4):
        for r in range(world):
            engs[r].reduce_and_update(phases=ph)
            torch.cuda.synchronize()
            print('step',step,'phase',ph,'rank',r,'err',[int(x.error.item()) for x in xs], 'px.step', xs[r].step)
        for r in range(world): print('  flags rank',r, flags(r).tolist())
print('tables equal', torch.equal(engs[0].table, engs[1].table))
