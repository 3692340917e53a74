import sys; sys.path.insert(0,'tests')
import numpy as np, b200
from conftest import nerr
L=b200.lib(); f32=np.float32
rng=np.random.default_rng(0)
m,n,l=128,64,256
x=rng.standard_normal((m,n)).astype(f32); w=(rng.standard_normal((l,n))/8).astype(f32); b=rng.standard_normal(l).astype(f32)
dx,dw,db,dy=b200.dev(x),b200.dev(w),b200.dev(b),b200.dev_empty((m,l))
for act,fn in ((0,lambda z:z),(1,lambda z:np.maximum(z,0)),(2,np.tanh)):
    L.ppo_b200_tc_linear(0,dy.ptr,dx.ptr,dw.ptr,db.ptr,m,n,l,act,1)
    ref=fn(x.astype(np.float64)@w.astype(np.float64).T+b)
    print("fwd act",act,nerr(dy.numpy(),ref))
# dX: structured test: g = one-hot rows, W = index pattern
g=np.zeros((m,l),f32); 
for i in range(m): g[i, i%l]=1.0
W=(np.arange(l)[:,None]*1000+np.arange(n)[None,:]).astype(f32)   # W[k][j]=1000k+j
dg,dW,dgx=b200.dev(g),b200.dev(W),b200.dev_empty((m,n))
L.ppo_b200_tc_linear(1,dgx.ptr,dg.ptr,dW.ptr,None,m,n,l,0,1)
got=dgx.numpy(); ref=g@W
print("dx nerr",nerr(got,ref))
print(got[:3,:8]); print(ref[:3,:8])
print(got[5,:40])
# dW structured: g[k][a] one-hot in a, x[k][j] pattern
m2=32
g2=np.zeros((m2,128),f32); x2=np.zeros((m2,64),f32)
for k in range(m2): g2[k,k]=1.0; x2[k,:]=k*100+np.arange(64)
dg2,dx2,do=b200.dev(g2),b200.dev(x2),b200.dev_empty((128,64))
L.ppo_b200_tc_linear(2,do.ptr,dg2.ptr,dx2.ptr,None,m2,64,128,0,1)
got=do.numpy(); ref=g2.T@x2
print("dw nerr",nerr(got,ref)); print(got[:4,:8]); print(ref[:4,:8]); print(got[9,:40])
