#!/usr/bin/env python
"""tools/prof_misc.py -- one launch each of the non-C2 kernels that carry BASELINE configs, for an ncu --set full
capture: k_map_flat2 (C3a column + row, C3b matrix * row vector at 4096^2 and 16384^2), k_reduce_flat (full sum
of 2^28 doubles: single launch, the last CTA folds), k_reduce_cols_tma / k_reduce_rows (C4 axis 0 / axis 1),
k_dot (sdot 2^28), k_reduce_flat on int8 extrema."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
rng = np.random.default_rng(0)
col = nu.asarray(rng.random((8192, 1), dtype=np.float32))
row = nu.asarray(rng.random((1, 8192), dtype=np.float32))
o8 = nu.empty((8192, 8192), "single-float")
m4, v4, o4 = nu.rand(4096, 4096, type="single-float"), nu.asarray(rng.random(4096, dtype=np.float32)), nu.empty((4096, 4096), "single-float")
m16, v16, o16 = nu.rand(16384, 16384, type="single-float"), nu.asarray(rng.random(16384, dtype=np.float32)), nu.empty((16384, 16384), "single-float")
d = nu.rand(16384, 16384, type="double-float")
oc = nu.empty((16384,), "double-float")
b8 = nu.DenseArray(m16.storage, "(signed-byte 8)", (1 << 28,))
for _ in range(2):
    nu.add(col, row, out=o8)
    nu.multiply(m4, v4, out=o4)
    nu.multiply(m16, v16, out=o16)
    s = nu.sum(d)
    nu.sum(d, axes=0, out=oc)
    nu.sum(d, axes=1, out=oc)
    nu.maximum(b8)
    nu.vdot(m16, o16)
_lib.check(_lib.lib().nuk_device_sync())
print("ok", nu.launch_count(), "launches; sum", s)
