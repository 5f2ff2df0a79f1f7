#!/usr/bin/env python
"""tools/time_logb.py -- nu:log x y on two 2^28-element arrays: the fused one-pass map (NUK_LOGB) against the
reference's composition (log, log, divide: src/transcendental/two-arg-fn-float.lisp:531-547), CUDA events."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
lib = _lib.lib()
n = 1 << 28
rng = np.random.default_rng(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0))
lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
for name, dt in (("single-float", np.float32), ("double-float", np.float64)):
    base = (rng.random(1 << 22) * 9 + 1.5).astype(dt)
    x, b, o, t = (nu.empty((n,), name) for _ in range(4))
    for a in (x, b):
        a.copy_from_numpy(np.tile(base, n >> 22))
    def timed(fn):
        fn(); lib.nuk_stream_sync(st)
        lib.nuk_event_record(e0, st)
        for _ in range(5):
            fn()
        lib.nuk_event_record(e1, st)
        lib.nuk_event_sync(e1)
        ms = C.c_float()
        lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value / 5
    fused = timed(lambda: nu.log(x, b, out=o, broadcast=False))
    def three():
        nu.log(b, out=o, broadcast=False); nu.log(x, out=t, broadcast=False); nu.divide(t, o, out=o, broadcast=False)
    comp = timed(three)
    sz = np.dtype(dt).itemsize
    print("%s: fused %.3f ms (%.0f GB/s of 3*%d B/elem), three passes %.3f ms  -> %.2fx" %
          (name, fused, 3 * sz * n / fused / 1e6, sz, comp, comp / fused), flush=True)
