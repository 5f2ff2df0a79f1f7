#!/usr/bin/env python
"""tools/prof_c2.py -- the C2 step (nu:exp nu:sin nu:log nu:tanh, f32, 2^28, device-resident) run a
few times with nothing else around it: the command profiled by ncu for profiles/ (launch list and
the --set full capture of the dominant kernel).  Usage: prof_c2.py [steps] [ufunc,ufunc,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ufuncs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["exp", "sin", "log", "tanh"]
nu.require_device(0)
n = 1 << 28
m = 1 << 22
chunk = nu.asarray((np.random.default_rng(0).random(m) * 0.999 + 1e-3).astype(np.float32))
x = nu.empty((n,), "single-float")
nu.copy(chunk.broadcast_to((n // m, m)), out=x.reshape(n // m, m))
out = nu.empty((n,), "single-float")
for _ in range(steps):
    for f in ufuncs:
        getattr(nu, f)(x, out=out, broadcast=False)
from numericals_b200 import _lib  # noqa: E402
_lib.check(_lib.lib().nuk_device_sync())
print("ok", nu.launch_count(), "launches; tanh sample", out.aref(slice(0, 4)).to_numpy())
