#!/usr/bin/env python
"""tools/time_ops.py -- CUDA-event timing of a list of BMAS symbols (device-resident, 2^LOG2N
elements).  usage: time_ops.py LOG2N sym[,sym...]   (NUK_LIB=path selects an A/B build of libnuk)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

log2n = int(sys.argv[1])
syms = sys.argv[2].split(",")
nu.require_device(0)
lib = _lib.lib()
n = 1 << log2n
rng = np.random.default_rng(0)
bufs = {}
for p, dt in (("s", np.float32), ("d", np.float64)):
    bufs[p] = [nu.asarray((rng.random(n) * 0.9 + 0.05).astype(dt)) for _ in range(2)] + [nu.empty((n,), dt)]
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0))
lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
L, P = C.c_long, C.c_void_p
BIN = {"pow", "atan2", "add", "sub", "mul", "div", "max", "min"}
for s in syms:
    x, y, o = bufs[s[0]]
    sz = 4 if s[0] == "s" else 8
    binary = s[1:] in BIN
    args = (L(n), P(x.ptr), L(1), P(y.ptr), L(1), P(o.ptr), L(1)) if binary else (L(n), P(x.ptr), L(1), P(o.ptr), L(1))
    f = _lib.bmas(s)
    f(*args)
    lib.nuk_stream_sync(st)
    lib.nuk_event_record(e0, st)
    for _ in range(10):
        f(*args)
    lib.nuk_event_record(e1, st)
    lib.nuk_event_sync(e1)
    ms = C.c_float()
    lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
    ms = ms.value / 10
    nb = (3 if binary else 2) * sz * n
    print("%-8s %.4f ms %7.0f GB/s" % (s, ms, nb / ms / 1e6))
