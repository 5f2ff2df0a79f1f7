import sys; sys.path.insert(0,".")
import numpy as np, ctypes as C
import numericals_b200 as nu
from numericals_b200 import _lib
nu.require_device(0)
lib=_lib.lib()
import torch
d=torch.rand(16384,16384,dtype=torch.float64,device="cuda")
from numericals_b200.array import from_pointer
D=from_pointer(d.data_ptr(),"double-float",(16384,16384),owner=d)
o=nu.empty((16384,),"double-float")
lib.nuk_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
def t(fn,reps=10):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
for name in ("sum","maximum","variance"):
    f=getattr(nu,name)
    lib.nuk_set_option(b"seq_cols",1); ms1=t(lambda: f(D,axes=0,out=o))
    lib.nuk_set_option(b"seq_cols",0); ms0=t(lambda: f(D,axes=0,out=o))
    print(name,"seq(TMA, reference order) %.3f ms %.0f GB/s | split %.3f ms %.0f GB/s"%(ms1,2**31/ms1/1e6,ms0,2**31/ms0/1e6))
