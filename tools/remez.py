#!/usr/bin/env python
"""tools/remez.py -- weighted minimax polynomial fitting (Remez exchange) in mpmath.

Generates the coefficients used by numericals_b200/csrc/nukmath_f32.cuh and
nukmath_f64.cuh.  Nothing here is copied from SLEEF, fdlibm or CUDA's libdevice: every
coefficient set in the math headers is reproducible with the command recorded next to it.

    minimise over c   max_{t in [a,b]}  | w(t) * (g(t) - sum_k c_k t^k) |

g and w are given as Python expressions in `t` using mpmath names (sin, cos, exp, log, sqrt,
atan, ...).  Coefficients are optionally rounded to binary32/binary64 one at a time, highest
degree... lowest first ("--round f32|f64"), re-fitting the remaining ones after each
rounding so the rounded set is itself near-minimax.
"""
import argparse
import struct
import sys

import mpmath as mp

mp.mp.dps = 40


def _solve(g, w, pts, nco, fixed):
    """Solve for free coeffs + levelled error E at reference points."""
    free = [k for k in range(nco) if k not in fixed]
    n = len(free)
    assert len(pts) == n + 1
    A = mp.matrix(n + 1, n + 1)
    rhs = mp.matrix(n + 1, 1)
    # the weight may vanish at an interior point (e.g. t^2 at 0): keep reference points off it
    wref = max(abs(w(pts[0])), abs(w(pts[-1])))
    span = pts[-1] - pts[0]
    pts = [t if abs(w(t)) > wref * mp.mpf("1e-6") else t + span / 200 for t in pts]
    for i, t in enumerate(pts):
        for j, k in enumerate(free):
            A[i, j] = t ** k
        A[i, n] = (-1) ** i / w(t)
        rhs[i] = g(t) - sum(c * t ** k for k, c in fixed.items())
    sol = mp.lu_solve(A, rhs)
    co = [mp.mpf(0)] * nco
    for j, k in enumerate(free):
        co[k] = sol[j]
    for k, c in fixed.items():
        co[k] = c
    return co, sol[n]


def _err(g, w, co, t):
    return w(t) * (g(t) - mp.polyval(co[::-1], t))


def remez(g, w, a, b, nco, fixed=None, iters=40, grid=1500):
    fixed = dict(fixed or {})
    nfree = nco - len(fixed)
    a, b = mp.mpf(a), mp.mpf(b)
    # Chebyshev nodes as the initial reference
    pts = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (nfree - i) / nfree) for i in range(nfree + 1)]
    co, E = None, None
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (grid - i) / grid) for i in range(grid + 1)]
    for _ in range(iters):
        co, E = _solve(g, w, pts, nco, fixed)
        es = [_err(g, w, co, t) for t in xs]
        # local extrema of the error on the grid (ends included)
        ext = []
        for i in range(len(xs)):
            l = es[i - 1] if i > 0 else None
            r = es[i + 1] if i + 1 < len(xs) else None
            e = es[i]
            if (l is None or abs(e) >= abs(l)) and (r is None or abs(e) >= abs(r)):
                ext.append((xs[i], e))
        # keep an alternating subsequence with the largest magnitudes
        alt = []
        for t, e in ext:
            if alt and (e > 0) == (alt[-1][1] > 0):
                if abs(e) > abs(alt[-1][1]):
                    alt[-1] = (t, e)
            else:
                alt.append((t, e))
        while len(alt) > nfree + 1:
            # drop the smaller end, or the smallest adjacent pair
            if abs(alt[0][1]) < abs(alt[-1][1]):
                alt.pop(0)
            else:
                alt.pop()
        if len(alt) < nfree + 1:
            break  # error already at rounding level
        newpts = [t for t, _ in alt]
        emax = max(abs(e) for _, e in alt)
        emin = min(abs(e) for _, e in alt)
        pts = newpts
        if emax - emin < emax * mp.mpf("1e-6"):
            break
    es = [abs(_err(g, w, co, t)) for t in xs]
    return co, max(es)


def round_to(x, kind):
    if kind == "f32":
        return mp.mpf(struct.unpack("f", struct.pack("f", float(x)))[0])
    if kind == "f64":
        return mp.mpf(float(x))
    return x


def fit(g, w, a, b, nco, fixed=None, rnd=None, **kw):
    """Fit; with rnd, round coefficients from the highest degree down, refitting the rest."""
    fixed = dict(fixed or {})
    co, e = remez(g, w, a, b, nco, fixed, **kw)
    if not rnd:
        return co, e
    for k in range(nco - 1, -1, -1):
        if k in fixed:
            continue
        fixed[k] = round_to(co[k], rnd)
        if len(fixed) < nco:
            co, e = remez(g, w, a, b, nco, fixed, **kw)
    co = [fixed[k] for k in range(nco)]
    xs = [mp.mpf(a) + (mp.mpf(b) - mp.mpf(a)) * i / 4000 for i in range(4001)]
    e = max(abs(_err(g, w, co, t)) for t in xs)
    return co, e


def hexf(x, kind):
    if kind == "f32":
        v = struct.unpack("f", struct.pack("f", float(x)))[0]
        return "%sf  /* %s */" % (float(v).hex(), repr(v))
    v = float(x)
    return "%s  /* %s */" % (v.hex(), repr(v))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--g", required=True, help="target g(t), mpmath expression in t")
    ap.add_argument("--w", default="1", help="weight w(t)")
    ap.add_argument("--a", required=True)
    ap.add_argument("--b", required=True)
    ap.add_argument("--n", type=int, required=True, help="number of coefficients")
    ap.add_argument("--fix", default="", help="k=value,... coefficients held fixed")
    ap.add_argument("--round", default=None, choices=["f32", "f64"])
    args = ap.parse_args()
    ns = {k: getattr(mp, k) for k in dir(mp) if not k.startswith("_")}
    g = lambda t: eval(args.g, ns, {"t": t})
    w = lambda t: eval(args.w, ns, {"t": t})
    fixed = {}
    for item in filter(None, args.fix.split(",")):
        k, v = item.split("=")
        fixed[int(k)] = mp.mpf(eval(v, ns))
    co, e = fit(g, w, eval(args.a, ns), eval(args.b, ns), args.n, fixed, args.round)
    print("# max weighted error: %s  (2^%.2f)" % (mp.nstr(e, 5), float(mp.log(e, 2)) if e > 0 else -999))
    for k, c in enumerate(co):
        print("c%-2d = %s" % (k, hexf(c, args.round) if args.round else mp.nstr(c, 25)))


if __name__ == "__main__":
    sys.exit(main())
