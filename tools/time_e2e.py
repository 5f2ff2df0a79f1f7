#!/usr/bin/env python
"""tools/time_e2e.py -- BMAS_sexp(n, host x, 1, host out, 1) with pinned host buffers (the staged path of
csrc/bmas.cu:run_map_1d), wall clock, 2^28 single-floats; NUK_STAGE_MB / NUK_STAGE_SLOTS select the pipeline shape."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
n = 1 << 28
x = torch.rand(n, dtype=torch.float32).pin_memory()
o = torch.empty(n, dtype=torch.float32).pin_memory()
f = _lib.bmas("sexp")
args = (C.c_long(n), C.c_void_p(x.data_ptr()), C.c_long(1), C.c_void_p(o.data_ptr()), C.c_long(1))
f(*args)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    f(*args)
    best = min(best, time.perf_counter() - t0)
assert _lib.last_status() == 0
print("stage_mb=%s slots=%s: %.2f ms  %.2f Gelem/s  %.1f GB/s each way" % (
    os.environ.get("NUK_STAGE_MB", "32"), os.environ.get("NUK_STAGE_SLOTS", "3"), best * 1e3, n / best / 1e9, 4 * n / best / 1e9))
