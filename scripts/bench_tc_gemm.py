import sys, ctypes as C; sys.path.insert(0,'tests')
import numpy as np, b200
L=b200.lib(); f32=np.float32
rng=np.random.default_rng(0)
m,n,l=65536,1024,1024
x=b200.dev(rng.standard_normal((m,n)).astype(f32)); w=b200.dev((rng.standard_normal((l,n))/32).astype(f32)); b=b200.dev(np.zeros(l,f32))
g=b200.dev(rng.standard_normal((m,l)).astype(f32)); y=b200.dev_empty((m,l)); gx=b200.dev_empty((m,n)); gw=b200.dev_empty((8,l,n))
def run(mode, reps=10, splits=4):
    for _ in range(3):
        if mode==0: L.ppo_b200_tc_linear(0,y.ptr,x.ptr,w.ptr,b.ptr,m,n,l,1,1)
        elif mode==1: L.ppo_b200_tc_linear(1,gx.ptr,g.ptr,w.ptr,x.ptr,m,n,l,1,1)
        else: L.ppo_b200_tc_linear(2,gw.ptr,g.ptr,x.ptr,None,m,n,l,0,splits)
    L.ppo_b200_sync(); L.ppo_b200_profile_begin()
    for _ in range(reps):
        if mode==0: L.ppo_b200_tc_linear(0,y.ptr,x.ptr,w.ptr,b.ptr,m,n,l,1,1)
        elif mode==1: L.ppo_b200_tc_linear(1,gx.ptr,g.ptr,w.ptr,x.ptr,m,n,l,1,1)
        else: L.ppo_b200_tc_linear(2,gw.ptr,g.ptr,x.ptr,None,m,n,l,0,splits)
    buf=C.create_string_buffer(4096); L.ppo_b200_profile_end(buf,4096)
    for ln in buf.value.decode().splitlines():
        name,cnt,tot=ln.rsplit(" ",2); ms=float(tot)/int(cnt)
        print("mode",mode,"splits",splits,"%.3f ms  %.1f TFLOP/s"%(ms, 2.0*m*n*l/ms/1e9), name[:50])
run(0); run(1); run(2,splits=4); run(2,splits=8); run(2,splits=2)
