#!/usr/bin/env python
"""tools/time_regions.py -- asinh / acosh (both precisions) on arguments INSIDE their polynomial regions (|x| <= 1,
(1, 3]), on a MIX (every warp holds both kinds) and entirely OUTSIDE: the price of the region split for data the
per-warp vote cannot keep on the polynomial path.  One JSON line per (symbol, distribution)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

nu.require_device(0)
lib = _lib.lib()
peak = 6552.0
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
n = 1 << 28
rng = np.random.default_rng(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.nuk_event_create(C.byref(e0)); lib.nuk_event_create(C.byref(e1))
st = lib.nuk_get_stream()
L, P = C.c_long, C.c_void_p
DIST = {"asinh": {"inside": (0.05, 0.95), "mixed": (0.05, 4.0), "outside": (1.5, 40.0)},
        "acosh": {"inside": (1.05, 2.95), "mixed": (1.05, 12.0), "outside": (3.5, 40.0)}}
for p, dt in (("s", np.float32), ("d", np.float64)):
    x, o = nu.empty((n,), dt), nu.empty((n,), dt)
    m = 1 << 22
    for fn, dists in DIST.items():
        for name, (lo, hi) in dists.items():
            chunk = nu.asarray((rng.random(m) * (hi - lo) + lo).astype(dt))
            nu.copy(chunk.broadcast_to((n // m, m)), out=x.reshape(n // m, m))
            f = _lib.bmas(p + fn)
            args = (L(n), P(x.ptr), L(1), P(o.ptr), L(1))
            f(*args); lib.nuk_stream_sync(st)
            best = None
            for _ in range(3):
                lib.nuk_event_record(e0, st)
                for _ in range(10):
                    f(*args)
                lib.nuk_event_record(e1, st); lib.nuk_event_sync(e1)
                ms = C.c_float(); lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
                best = ms.value / 10 if best is None else min(best, ms.value / 10)
            gbs = 2 * np.dtype(dt).itemsize * n / (best / 1e3) / 1e9
            print(json.dumps({"symbol": "BMAS_" + p + fn, "arguments": "%s U[%g, %g)" % (name, lo, hi), "ms": round(best, 4),
                              "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
    del x, o
