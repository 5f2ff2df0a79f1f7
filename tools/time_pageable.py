#!/usr/bin/env python
"""tools/time_pageable.py -- BMAS_sexp(2^28) on PAGEABLE host buffers (a Lisp heap vector) through the staged
path; the mover-thread count and the bounce switch come from the environment (NUK_STAGE_THREADS,
NUK_STAGE_BOUNCE=0 = the driver's own pageable copy), one configuration per process."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numericals_b200 import _lib  # noqa: E402

lib = _lib.lib()
_lib.require_device(0)
n = 1 << 28
x = np.random.default_rng(0).random(n, dtype=np.float32)
o = np.zeros(n, dtype=np.float32)
f = _lib.bmas("sexp")
args = (C.c_long(n), C.c_void_p(x.ctypes.data), C.c_long(1), C.c_void_p(o.ctypes.data), C.c_long(1))
f(*args)
best = 1e9
for _ in range(4):
    t0 = time.perf_counter()
    f(*args)
    best = min(best, time.perf_counter() - t0)
assert _lib.last_status() == 0
print("threads=%s bounce=%s: %.2f ms  %.2f Gelem/s  %.1f GB/s each way" % (
    lib.nuk_host_mover_threads() if os.environ.get("NUK_STAGE_BOUNCE", "1") != "0" else "-", os.environ.get("NUK_STAGE_BOUNCE", "1"),
    best * 1e3, n / best / 1e9, 4 * n / best / 1e9), flush=True)
