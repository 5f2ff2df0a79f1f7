#!/usr/bin/env python
"""tools/sweep.py -- achieved HBM GB/s of EVERY contiguous one-/two-arg ufunc and full reduction
of the hot path on one B200, device-resident (the north-star's per-ufunc metric).

Calls the BMAS_<name> symbols of libnuk.so with device pointers, times them with CUDA events on
the launching stream (nuk_event_*), arrays of 2^28 elements (>= 256 MiB each: larger than L2).
Writes one JSON line per kernel; `--out` collects them in a file for profiles/.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import numericals_b200 as nu  # noqa: E402
from numericals_b200 import _lib  # noqa: E402

SIZES = {"s": 4, "d": 8, "i64": 8, "i32": 4, "i16": 2, "i8": 1, "u64": 8, "u32": 4, "u16": 2, "u8": 1}
FLOATS, SINTS, ALL = ["s", "d"], ["i64", "i32", "i16", "i8"], list(SIZES)
HELPER_ETYPE = {"s": "single-float", "d": "double-float", "i64": "(signed-byte 64)", "i32": "(signed-byte 32)",
                "i16": "(signed-byte 16)", "i8": "(signed-byte 8)", "u64": "(unsigned-byte 64)",
                "u32": "(unsigned-byte 32)", "u16": "(unsigned-byte 16)", "u8": "(unsigned-byte 8)"}
TRANS1 = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "log", "sqrt"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None, help="comma-separated substrings of the symbol name (dexp,spow); ^d = prefix")
    ap.add_argument("--cooldown", type=float, default=0.0,
                    help="idle seconds before each symbol: the default sweep runs ~150 kernels back to back and the FP64-heavy "
                         "maps then sit under the board's power cap (SM clock 1965 -> 1500-1900 MHz, profiles/clockwatch_r01.txt); "
                         "with a pause each symbol is timed like a kernel launched on its own")
    args = ap.parse_args()
    nu.require_device(0)
    lib = _lib.lib()
    peak = 6552.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    n = 1 << args.log2n
    rng = np.random.default_rng(0)
    # three 8-byte-per-element buffers reused for every dtype
    base = rng.random(n // 4).astype(np.float64)
    bufs = []
    for k in range(3):
        a = nu.empty((n,), "double-float")
        h = np.tile(base + k, 4)
        a.copy_from_numpy(h)
        bufs.append(a)
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.nuk_event_create(C.byref(e0))
    lib.nuk_event_create(C.byref(e1))
    st = lib.nuk_get_stream()
    rows = []

    def fill_float(p, lo, hi):
        """put U[lo,hi) of the right float type into buffers 0 and 1"""
        dt = np.float32 if p == "s" else np.float64
        m = 1 << 22
        for k in (0, 1):
            chunk = (rng.random(m) * (hi - lo) + lo).astype(dt)
            view = nu.DenseArray(bufs[k].storage, "single-float" if p == "s" else "double-float", (n,))
            t = nu.asarray(chunk, type=dt)
            nu.copy(t.broadcast_to((n // m, m)), out=view.reshape(n // m, m))

    def time_call(fn, call_args):
        if args.cooldown > 0:
            import time
            time.sleep(args.cooldown)
        fn(*call_args)
        lib.nuk_stream_sync(st)
        best = None
        for _ in range(3):          # best of 3 batches of `reps` back-to-back launches (one outlier batch seen in r01)
            lib.nuk_event_record(e0, st)
            for _ in range(args.reps):
                fn(*call_args)
            lib.nuk_event_record(e1, st)
            lib.nuk_event_sync(e1)
            ms = C.c_float()
            lib.nuk_event_elapsed_ms(e0, e1, C.byref(ms))
            best = ms.value if best is None else min(best, ms.value)
        return best / args.reps

    def record(sym, kind, bytes_per_elem, ms, sync_included=False):
        gbs = bytes_per_elem * n / (ms / 1e3) / 1e9
        row = {"symbol": "BMAS_" + sym, "kind": kind, "n": n, "ms": round(ms, 5), "gelem_s": round(n / (ms / 1e3) / 1e9, 2),
               "algorithmic_bytes_per_elem": bytes_per_elem, "hbm_gbs": round(gbs, 1),
               "frac_of_measured_peak": round(gbs / peak, 4), "frac_of_8tbs": round(gbs / 8000.0, 4)}
        if args.cooldown > 0:
            row["cooldown_s"] = args.cooldown
        if sync_included:
            row["note"] = "returns by value: includes a D2H of the scalar and a stream sync"
        rows.append(row)
        print(json.dumps(row), flush=True)

    L, P = C.c_long, C.c_void_p
    x, y, o = (P(b.ptr) for b in bufs)

    def want(sym):
        if args.only is None:
            return True
        return any(sym.startswith(t[1:]) if t.startswith("^") else t in sym for t in args.only.split(","))

    for p in ALL:
        sz = SIZES[p]
        if p in FLOATS:
            fill_float(p, 0.05, 0.95)
        for op in (["add", "sub"] if p in FLOATS + SINTS else []) + ["mul", "max", "min"] + (["div"] if p in FLOATS else []):
            sym = p + op
            if want(sym):
                record(sym, "binary", 3 * sz, time_call(_lib.bmas(sym), (L(n), x, L(1), y, L(1), o, L(1))))
        for op in ["lt", "eq"]:
            sym = p + op
            if want(sym):
                record(sym, "compare", 2 * sz + 1, time_call(_lib.bmas(sym), (L(n), x, L(1), y, L(1), o, L(1))))
        if p in FLOATS:
            for op in TRANS1 + ["fabs", "round", "floor"]:
                sym = p + op
                if op == "acosh":
                    fill_float(p, 1.05, 1.95)
                if want(sym):
                    record(sym, "unary", 2 * sz, time_call(_lib.bmas(sym), (L(n), x, L(1), o, L(1))))
                if op == "acosh":
                    fill_float(p, 0.05, 0.95)
            for op in ["pow", "atan2"]:
                sym = p + op
                if want(sym):
                    record(sym, "binary", 3 * sz, time_call(_lib.bmas(sym), (L(n), x, L(1), y, L(1), o, L(1))))
        if p in FLOATS + SINTS:
            sym = p + "copy"
            if want(sym):
                record(sym, "unary", 2 * sz, time_call(_lib.bmas(sym), (L(n), x, L(1), o, L(1))))
            sym = p + "dot"
            if want(sym):
                record(sym, "reduce", 2 * sz, time_call(_lib.bmas(sym), (L(n), x, L(1), y, L(1))), True)
        # full reductions through the async nd entry point (result stays on the device: no D2H, no
        # sync inside the timed region); the by-value BMAS symbol adds ~15 us of sync per call
        dims, st1 = (C.c_int64 * 1)(n), (C.c_int64 * 1)(1)
        for red, code in (("sum", 0), ("hmax", 1)):
            if red == "sum" and p not in FLOATS + SINTS:
                continue
            sym = p + red
            if want(sym):
                dcode = nu.dtypes.code(HELPER_ETYPE[p])
                record(sym, "reduce", sz, time_call(lib.nuk_reduce, (code, dcode, 1, dims, st1, -1, None, x, o, 0, st)))
        for q in FLOATS:
            if p != q:
                sym = "cast_" + p + q
                if want(sym):
                    record(sym, "cast", sz + SIZES[q], time_call(_lib.bmas(sym), (L(n), x, L(1), o, L(1))))
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
