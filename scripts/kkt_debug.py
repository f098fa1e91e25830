import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
import test_gpu_fullsize as T
inp,w=T._inputs(1024)
gen=torch.Generator(device="cuda").manual_seed(3)
X0=w.X+0.01*torch.randn(w.X.shape,device="cuda",generator=gen); U0=w.U+0.1*torch.randn(w.U.shape,device="cuda",generator=gen); V0=w.V.clone()
s,X,U,V=T._solve(1024,inp,(X0,U0,V0))
A,B,Q,R,M=(s.view(k).double() for k in ("A","B","Q","R","M")); q,r,c,dX,dU,dV=(s.view(k).double() for k in ("q","r","c","dX","dU","dV"))
mv=lambda Mx,v: torch.einsum("etij,etj->eti",Mx,v); mtv=lambda Mx,v: torch.einsum("etji,etj->eti",Mx,v)
sx=mv(Q[:,:-1],dX[:,:-1])+mv(M,dU)+mtv(A,dV[:,1:])-dV[:,:-1]+q[:,:-1]
terms=[mv(Q[:,:-1],dX[:,:-1]),mv(M,dU),mtv(A,dV[:,1:]),dV[:,:-1],q[:,:-1]]
print("term maxes",[float(t.abs().max()) for t in terms])
a=sx.abs(); idx=np.unravel_index(int(a.argmax()),a.shape); print("argmax (e,t,i)",idx,float(a.max()))
e,t,i=idx
print("terms at argmax",[float(x[e,t,i]) for x in terms])
print("per-component max residual", a.amax(dim=(0,1)).cpu().numpy().round(4))
print("per-knot max residual", a.amax(dim=(0,2)).cpu().numpy().round(4))
# row-relative residual: relative to the largest term in the same row
rowscale=torch.stack([x.abs() for x in terms]).amax(0).clamp_min(1.0)
print("max row-relative residual", float((a/rowscale).max()))
print("P max", float(s.view("P").abs().max()), "dX max", float(dX.abs().max()))
