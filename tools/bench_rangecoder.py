"""Host range coder micro-benchmark (CPU only): ns per symbol for aivc_rc_encode_bounds and the windowed Laplace decoder on a
synthetic 1080p-latent-sized stream (bench-leg use of oracle/ for the CDF tables)."""
import numpy as np, time, ctypes as C, sys
sys.path.insert(0,'/root/repo')
from aivc_b200 import _lib
from oracle import codec_ref as O
L=_lib.lib()
rng=np.random.default_rng(0)
n=522240
sig=np.exp(rng.uniform(-1,1.5,n)).astype(np.float32)
q=np.clip(np.rint(rng.laplace(0,sig/np.sqrt(2))),-256,255).astype(np.int16)
table=O.laplace_table_spec(sig)
ar=np.arange(n)
lo=table[ar,q.astype(int)+256].astype(np.uint32); hi=table[ar,q.astype(int)+257].astype(np.uint32)
bounds=(lo|(hi<<16)).astype(np.uint32)
out=np.empty(L.aivc_rc_bound(n),np.uint8); ln=C.c_size_t()
t=time.time(); _lib.check(L.aivc_rc_encode_bounds(bounds.ctypes.data,n,out.ctypes.data,out.size,C.byref(ln))); te=time.time()-t
b=(sig/np.float32(1.41421354)).astype(np.float32)
win=np.ascontiguousarray(table[:,253:261]).astype(np.uint16)
dec=np.empty(n,np.int16)
best=1e9
for _ in range(5):
    t=time.time(); L.aivc_rc_decode_laplace_win(b.ctypes.data,win.ctypes.data,out.ctypes.data,ln.value,n,dec.ctypes.data); best=min(best,time.time()-t)
print('encode %.1f ns/sym; win decode %.1f ns/symbol'%(te/n*1e9,best/n*1e9), np.array_equal(dec,q), 'inside window: %.3f'%np.mean(np.abs(q)<=3))
