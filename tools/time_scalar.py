#!/usr/bin/env python
"""tools/time_scalar.py -- the scalar polymorphs of the split functors: nu:expt x 2.5, nu:expt 2.5 x, nu:atan2 x 0.7
on 2^28 (f32) / 2^27 (f64) elements, CUDA events.  NUK_LIB selects an A/B build."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
lib = _lib.lib()
rng = np.random.default_rng(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0))
lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
for name, dt, n in (("single-float", np.float32, 1 << 28), ("double-float", np.float64, 1 << 27)):
    base = (rng.random(1 << 22) * 4 + 0.05).astype(dt)
    x, o = nu.empty((n,), name), nu.empty((n,), name)
    x.copy_from_numpy(np.tile(base, n >> 22))
    for label, fn in (("expt x 2.5", lambda: nu.expt(x, 2.5, out=o)), ("expt 2.5 x", lambda: nu.expt(2.5, x, out=o)),
                      ("atan2 x 0.7", lambda: nu.atan2(x, 0.7, out=o))):
        fn(); lib.nuk_stream_sync(st)
        lib.nuk_event_record(e0, st)
        for _ in range(10):
            fn()
        lib.nuk_event_record(e1, st)
        lib.nuk_event_sync(e1)
        ms = C.c_float()
        lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
        sz = np.dtype(dt).itemsize
        print("%s %-12s %.3f ms  %.0f GB/s (2 x %d B/elem)" % (name, label, ms.value / 10, 2 * sz * n / (ms.value / 10) / 1e6, sz), flush=True)
