#!/usr/bin/env python
"""tools/time_c3.py -- BASELINE config C3 shapes through the rows kernel, by rows-per-CTA (option rows_rpc; 0 = the
built-in rule) and against the flat kernel on the same number of bytes: (8192 x 1) + (1 x 8192), (4096 x 4096) * (4096),
(16384 x 16384) * (16384), single-float, OUT preallocated.  Warm = back-to-back launches; cold = L2 flushed (clean)
before every launch."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
lib = _lib.lib()
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0))
lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
PEAK = 6552.0


def timed(fn, reps, cold):
    fn()
    lib.nuk_stream_sync(st)
    ms = C.c_float()
    if not cold:
        lib.nuk_event_record(e0, st)
        for _ in range(reps):
            fn()
        lib.nuk_event_record(e1, st)
        lib.nuk_event_sync(e1)
        lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
        return ms.value / reps
    tot = 0.0
    for _ in range(reps):
        lib.nuk_flush_l2(st)
        lib.nuk_event_record(e0, st)
        fn()
        lib.nuk_event_record(e1, st)
        lib.nuk_event_sync(e1)
        lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
        tot += ms.value
    return tot / reps


rng = np.random.default_rng(0)
col = nu.asarray(rng.random((8192, 1), dtype=np.float32))
row = nu.asarray(rng.random((1, 8192), dtype=np.float32))
o8 = nu.empty((8192, 8192), "single-float")
m4 = nu.asarray(rng.random((4096, 4096), dtype=np.float32))
v4 = nu.asarray(rng.random(4096, dtype=np.float32))
o4 = nu.empty((4096, 4096), "single-float")
m16 = nu.rand(16384, 16384, type="single-float")
v16 = nu.asarray(rng.random(16384, dtype=np.float32))
o16 = nu.empty((16384, 16384), "single-float")
f4a, f4b, f4o = m4.reshape(4096 * 4096), nu.rand(4096 * 4096, type="single-float"), o4.reshape(4096 * 4096)
cases = (("C3a (8192x1)+(1x8192)", lambda: nu.add(col, row, out=o8), 4 * 8192 * 8192 + 65536),
         ("C3b (4096^2)*(4096)", lambda: nu.multiply(m4, v4, out=o4), 8 * 4096 * 4096 + 16384),
         ("C3b (16384^2)*(16384)", lambda: nu.multiply(m16, v16, out=o16), 8 * 16384 * 16384))
for rpc in (0, 8, 32):
    lib.nuk_set_option(b"rows_rpc", rpc)
    for name, fn, by in cases:
        w, c = timed(fn, 20, False), timed(fn, 10, True)
        print("rows_rpc=%-2d %-24s warm %8.2f us %5.3f   cold %8.2f us %5.3f of measured peak" % (
            rpc, name, w * 1e3, by / (w / 1e3) / 1e9 / PEAK, c * 1e3, by / (c / 1e3) / 1e9 / PEAK), flush=True)
lib.nuk_set_option(b"rows_rpc", 0)
# what the write side alone can do: fill (flat kernel, scalar operand), row vector * scalar, column vector + scalar
wby = 4 * 8192 * 8192
for rpc in (0, 16):
    lib.nuk_set_option(b"rows_rpc", rpc)
    for name, fn in (("fill 8192^2 (flat, write only)", lambda: nu.fill(o8, 1.5)),
                     ("row * scalar -> 8192^2", lambda: nu.multiply(row, 2.0, out=o8)),
                     ("col + scalar -> 8192^2", lambda: nu.add(col, 2.0, out=o8)),
                     ("col + row -> 8192^2", lambda: nu.add(col, row, out=o8))):
        w, c = timed(fn, 20, False), timed(fn, 10, True)
        print("rows_rpc=%-2d %-32s warm %8.2f us %5.3f   cold %8.2f us %5.3f" % (
            rpc, name, w * 1e3, wby / (w / 1e3) / 1e9 / PEAK, c * 1e3, wby / (c / 1e3) / 1e9 / PEAK), flush=True)
lib.nuk_set_option(b"rows_rpc", 0)
by = 12 * 4096 * 4096
w, c = timed(lambda: nu.multiply(f4a, f4b, out=f4o, broadcast=False), 20, False), timed(lambda: nu.multiply(f4a, f4b, out=f4o, broadcast=False), 10, True)
print("flat kernel, 2^24 x 2^24 -> 2^24 (192 MiB)  warm %8.2f us %5.3f   cold %8.2f us %5.3f" % (
    w * 1e3, by / (w / 1e3) / 1e9 / PEAK, c * 1e3, by / (c / 1e3) / 1e9 / PEAK))
for label, n in (("copy 2^24 f32", 1 << 24), ("copy 2^26 f32", 1 << 26)):
    a, b = nu.rand(n, type="single-float"), nu.empty((n,), "single-float")
    w, c = timed(lambda: nu.copy(a, out=b), 20, False), timed(lambda: nu.copy(a, out=b), 10, True)
    print("%-16s warm %8.2f us %5.3f   cold %8.2f us %5.3f" % (label, w * 1e3, 8 * n / (w / 1e3) / 1e9 / PEAK, c * 1e3, 8 * n / (c / 1e3) / 1e9 / PEAK))
