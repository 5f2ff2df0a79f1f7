#!/usr/bin/env python
"""tools/time_rows.py -- the rows kernel (matrix op row-vector, the C3b shape) with the table / split functors:
(8192 x 8192) ^ (8192) and tanh/atan2 on the same shapes, CUDA events.  NUK_LIB selects an A/B build."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
lib = _lib.lib()
r = c = 8192
rng = np.random.default_rng(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0))
lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
for name, dt in (("single-float", np.float32), ("double-float", np.float64)):
    a = nu.asarray((rng.random((r, c)) * 4 + 0.05).astype(dt), type=dt)
    v = nu.asarray((rng.random(c) * 6 - 3).astype(dt), type=dt)
    o = nu.empty((r, c), name)
    for label, fn in (("expt", lambda: nu.expt(a, v, out=o)), ("atan2", lambda: nu.atan2(a, v, out=o))):
        fn(); lib.nuk_stream_sync(st)
        lib.nuk_event_record(e0, st)
        for _ in range(10):
            fn()
        lib.nuk_event_record(e1, st)
        lib.nuk_event_sync(e1)
        ms = C.c_float()
        lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
        sz = np.dtype(dt).itemsize
        print("%s %-6s %.3f ms  %.0f GB/s (2 x %d B/elem)" % (name, label, ms.value / 10, 2 * sz * r * c / (ms.value / 10) / 1e6, sz), flush=True)
