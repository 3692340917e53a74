import sys, ctypes as C, time; sys.path.insert(0,'tests')
import numpy as np, b200, oracle
L=b200.lib(); f32=np.float32; u8=np.uint8
T=2048; N=int(sys.argv[1]) if len(sys.argv)>1 else 65536
rng=np.random.default_rng(0)
NB=4096
def block():
    n=T*NB
    r,v,vn=(rng.standard_normal(n).astype(f32) for _ in range(3))
    term=(rng.random(n)<1e-3).astype(u8)
    t=np.tile(np.arange(T),NB); trunc=(((t+1)%1000)==0).astype(u8); trunc[t==T-1]=1
    return r,v,vn,term,trunc
blk=block(); reps=N//NB; n=T*N
bufs=[]
for a in blk:
    d=b200.dev_empty(n, a.dtype)
    for i in range(reps): L.ppo_b200_h2d(d.ptr + i*a.nbytes, a.ctypes.data, a.nbytes)
    bufs.append(d)
adv=b200.dev_empty(n); tgt=b200.dev_empty(n); st=b200.dev_empty(2)
def run(norm, reps=10):
    for _ in range(3): L.ppo_b200_gae(*[b.ptr for b in bufs], n, 0.99, 0.95, adv.ptr, tgt.ptr, norm, st.ptr)
    L.ppo_b200_sync(); L.ppo_b200_profile_begin()
    for _ in range(reps): L.ppo_b200_gae(*[b.ptr for b in bufs], n, 0.99, 0.95, adv.ptr, tgt.ptr, norm, st.ptr)
    buf=C.create_string_buffer(4096); L.ppo_b200_profile_end(buf,4096)
    for ln in buf.value.decode().splitlines():
        name,cnt,tot=ln.rsplit(" ",2); ms=float(tot)/int(cnt)
        bpe={"gae_scan_kernel":22,"gae_normalize_kernel":8}.get(name,0)
        print("%-24s %.3f ms  %.0f GB/s (%.1f%% of 6452.8)"%(name, ms, bpe*n/ms/1e6, 100*bpe*n/ms/1e6/6452.8))
run(1)
# correctness: sampled envs
a=np.empty(T*NB,f32); L.ppo_b200_gae(*[b.ptr for b in bufs], n, 0.99, 0.95, adv.ptr, tgt.ptr, 0, st.ptr)
L.ppo_b200_d2h(a.ctypes.data, adv.ptr + (reps-1)*a.nbytes, a.nbytes)
e=17; s=slice(e*T,(e+1)*T)
raw,_,_,_,_=oracle.gae(blk[0][s],blk[1][s],blk[2][s],blk[3][s],blk[4][s],0.99,0.95)
print("sampled env max err", np.abs(a[s]-raw).max()/np.abs(raw).max())
