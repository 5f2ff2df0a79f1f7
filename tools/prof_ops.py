#!/usr/bin/env python
"""tools/prof_ops.py -- run a list of BMAS_<symbol> calls once each on device-resident arrays
(the command ncu profiles for per-kernel evidence).  usage: prof_ops.py LOG2N sym[,sym...] [reps]
e.g. prof_ops.py 26 dsin,dtanh,dlog,spow"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
syms = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dsin"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
nu.require_device(0)
n = 1 << log2n
rng = np.random.default_rng(0)
bufs = {}
for p, dt in (("s", np.float32), ("d", np.float64)):
    bufs[p] = [nu.asarray((rng.random(n) * 0.9 + 0.05).astype(dt)) for _ in range(2)] + [nu.empty((n,), dt)]
    bufs[p + "+1"] = nu.asarray((rng.random(n) * 0.9 + 1.05).astype(dt))     # acosh: U[1.05, 1.95) as tools/sweep.py
L, P = C.c_long, C.c_void_p
BIN = {"pow", "atan2", "add", "sub", "mul", "div", "max", "min"}
for _ in range(reps):
    for s in syms:
        p = s[0]
        x, y, o = bufs[p]
        if s[1:] == "acosh":
            x = bufs[p + "+1"]
        if s[1:] in BIN:
            _lib.bmas(s)(L(n), P(x.ptr), L(1), P(y.ptr), L(1), P(o.ptr), L(1))
        else:
            _lib.bmas(s)(L(n), P(x.ptr), L(1), P(o.ptr), L(1))
        assert _lib.last_status() == 0, _lib.lib().nuk_last_error()
_lib.check(_lib.lib().nuk_device_sync())
print("ok", nu.launch_count())
