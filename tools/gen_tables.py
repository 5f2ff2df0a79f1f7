#!/usr/bin/env python
"""tools/gen_tables.py -- generate numericals_b200/csrc/nukmath_tables.inc: the lookup tables and
polynomial coefficients of the table-driven double-float log / exp cores (pow_f64 and friends).

Nothing is copied from a library: every number below is computed here with mpmath (80 digits) and
checked (exactness of the reduced argument with rationals, table residuals, polynomial error).

log core (x = 2^k z, z in [Z0, 2 Z0), Z0 = 0.705078125, 128 sub-intervals of equal width in the
bit pattern; sub-interval 75 is [1 - 2^-9, 1 + 2^-8) so that x near 1 reduces with c = 1, log c = 0):
    log x = k ln2 + log c_i + log1p(r),  r = z / c_i - 1 = fma(z, invc_i, -1)  EXACTLY, because
    invc_i = j/128 (z < 1) or j/256 (z >= 1) has so few bits that z*invc - 1 fits in 53 bits.
    logc_i = log(1/invc_i) rounded to a multiple of 2^-43 (so k*LN2HI + logc is exact),
    logctail_i = log(1/invc_i) - logc_i rounded to double.
exp core (e^x = 2^(ki/128) e^r, ki = rint(128 x / ln2)):
    2^(j/128) = H_j (1 + tail_j), H_j the nearest double; the table stores tail_j and
    bits(H_j) - (j << 45), so adding ki << 45 installs the binary exponent floor(ki/128).
"""
import os
import struct
import sys
from fractions import Fraction

import mpmath as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import remez  # noqa: E402

mp.mp.dps = 80
N = 128
OFF = 0x3FE6900000000000


def u2d(u):
    return struct.unpack("<d", struct.pack("<Q", u))[0]


def d2u(d):
    return struct.unpack("<Q", struct.pack("<d", d))[0]


def rd(x):
    """mpf -> nearest double"""
    return float(mp.mpf(x))


def log_table():
    rows, rmin, rmax = [], 0, 0
    for i in range(N):
        a = Fraction(u2d(OFF + (i << 45)))
        b = Fraction(u2d(OFF + ((i + 1) << 45)))
        if a < 1 < b:
            D, j = 1, 1
        else:
            D = 128 if b <= 1 else 256
            j0 = int(round(D * 2 / (a + b)))
            j = min((j0 - 1, j0, j0 + 1), key=lambda jj: max(abs(a * jj / D - 1), abs(b * jj / D - 1)))
        invc = Fraction(j, D)
        lo, hi = a * invc - 1, b * invc - 1
        assert max(abs(lo), abs(hi)) < Fraction(1, 128), (i, float(lo), float(hi))
        rmin, rmax = min(rmin, lo), max(rmax, hi)
        # exactness: z = a + t*ulp; r = z*invc - 1 must be a 53-bit number for every z in the interval
        ulp = Fraction(1, 2 ** 53) if b <= 1 else Fraction(1, 2 ** 52)
        if D != 1:
            q = ulp * Fraction(1, D)           # every r is a multiple of q
            assert max(abs(lo), abs(hi)) / q < 2 ** 53, i
        logc_full = mp.log(mp.mpf(D) / j)
        logc = mp.nint(logc_full * 2 ** 43) / 2 ** 43
        tail = rd(logc_full - logc)
        assert mp.mpf(float(logc)) == logc
        rows.append((float(invc), float(logc), tail))
        assert abs(logc_full - logc - tail) < mp.mpf(2) ** -96, i
    return rows, float(rmin), float(rmax)


def exp_table(n=N, shift=45):
    """2^(j/n) = H_j (1 + tail_j); the entry holds bits(H_j) - (j << shift) so that adding ki << shift (ki = n k + j)
    puts k into the exponent field (shift = 52 - log2 n)"""
    rows = []
    for j in range(n):
        v = mp.mpf(2) ** (mp.mpf(j) / n)
        H = rd(v)
        tail = rd((v - mp.mpf(H)) / mp.mpf(H))
        rows.append((tail, (d2u(H) - (j << shift)) & 0xFFFFFFFFFFFFFFFF))
    return rows


# ---- single-float pow: 32-entry tables (logc and 2^(j/32) as doubles, 1/c as floats = 640 bytes)
OFF32 = 0x3F360000          # z in [0.7109375, 1.421875); interval 18 = [1 - 2^-7, 1 + 2^-6) has c = 1


def f2u32(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


def u2f32(u):
    return struct.unpack("<f", struct.pack("<I", u))[0]


def pow32_tables():
    invc, logc, rmin, rmax = [], [], 0, 0
    for i in range(32):
        a = Fraction(u2f32(OFF32 + (i << 18)))
        b = Fraction(u2f32(OFF32 + ((i + 1) << 18)))
        if a < 1 < b:
            c = Fraction(1)
        else:
            # 1/c rounded to 12 bits: z (24 bits) * invc is exact in double, so is z*invc - 1
            c = Fraction(int(round(2 / (a + b) * 2 ** 12)), 2 ** 12)
        assert float(c) == u2f32(f2u32(float(c)))
        lo, hi = a * c - 1, b * c - 1
        rmin, rmax = min(rmin, lo), max(rmax, hi)
        invc.append(float(c))
        logc.append(rd(-mp.log(mp.mpf(c.numerator) / c.denominator)))
    expt = []
    for j in range(32):
        H = rd(mp.mpf(2) ** (mp.mpf(j) / 32))
        expt.append((d2u(H) - (j << 47)) & 0xFFFFFFFFFFFFFFFF)
    return invc, logc, expt, float(rmin), float(rmax)


def hexd(x):
    return "0x%016xull" % d2u(float(x))


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "numericals_b200", "csrc",
                       "nukmath_tables.inc")
    if len(sys.argv) > 1:          # gen_tables.py OUT: write somewhere else (tests/test_tables_host.py)
        out = sys.argv[1]
    lt, rmin, rmax = log_table()
    et = exp_table()
    # log1p(r) = r - r^2/2 + r^3 (B0 + B1 r + ... + B5 r^5) on [rmin, rmax]; error relative to r
    g = lambda t: (mp.log1p(t) - t + t * t / 2) / t ** 3 if t != 0 else mp.mpf(1) / 3
    w = lambda t: t * t
    span = max(-rmin, rmax) * 1.0001
    B, eb = remez.fit(g, w, -span, span, 6, None, "f64", grid=600)
    # e^r = 1 + r + r^2 (C0 + C1 r + C2 r^2 + C3 r^3) on |r| <= ln2/256 (1 + 2^-10); absolute error
    ge = lambda t: (mp.exp(t) - 1 - t) / t ** 2 if t != 0 else mp.mpf(1) / 2
    we = lambda t: t * t
    se = float(mp.log(2) / 256) * 1.002
    C, ec = remez.fit(ge, we, -se, se, 4, None, "f64", grid=600)
    # sinh / cosh: a 32-entry table (512 bytes: at most 4 rows of shared memory, so a per-lane indexed 16-byte read
    # costs 4 wavefronts instead of the ~8 of the 128-entry one) and two more coefficients:
    # e^r = 1 + r + r^2 (D0 + ... + D5 r^5) on |r| <= ln2/64 (1 + 2^-10); absolute error
    et32 = exp_table(32, 47)
    s32e = float(mp.log(2) / 64) * 1.002
    Dc, ed = remez.fit(ge, we, -s32e, s32e, 6, None, "f64", grid=600)
    l2n32_hi = mp.nint(mp.log(2) / 32 * 2 ** 42) / 2 ** 42   # ln2/32 to 37 bits: ki*hi exact for |ki| < 2^16
    l2n32_lo = rd(mp.log(2) / 32 - l2n32_hi)
    ln2 = mp.log(2)
    ln2hi = mp.nint(ln2 * 2 ** 42) / 2 ** 42                 # 42 bits: k*LN2HI exact for |k| < 2^11
    ln2lo = rd(ln2 - ln2hi)
    l2n_hi = mp.nint(ln2 / N * 2 ** 42) / 2 ** 42            # ln2/128 to 35 bits: ki*hi exact for |ki| < 2^18
    l2n_lo = rd(ln2 / N - l2n_hi)
    with open(out, "w") as f:
        f.write("// nukmath_tables.inc -- GENERATED by tools/gen_tables.py; do not edit.\n")
        f.write("// log1p polynomial: r in [%.6g, %.6g], max error relative to r 2^%.2f\n" %
                (rmin, rmax, float(mp.log(eb, 2))))
        f.write("// exp polynomial: |r| <= %.6g, max absolute error 2^%.2f\n" % (se, float(mp.log(ec, 2))))
        f.write("#define NM_TL_OFF 0x%016xull\n" % OFF)
        f.write("#define NM_TL_LN2HI %s\n" % float(ln2hi).hex())
        f.write("#define NM_TL_LN2LO %s\n" % float(ln2lo).hex())
        f.write("#define NM_TE_INVLN2N %s\n" % rd(N / ln2).hex())
        f.write("#define NM_TE_LN2HIN %s\n" % float(l2n_hi).hex())
        f.write("#define NM_TE_LN2LON %s\n" % float(l2n_lo).hex())
        f.write("// 32-entry exp (sinh / cosh): |r| <= %.6g, max absolute error 2^%.2f\n" % (s32e, float(mp.log(ed, 2))))
        f.write("#define NM_TE32_INVLN2N %s\n" % rd(32 / ln2).hex())
        f.write("#define NM_TE32_LN2HIN %s\n" % float(l2n32_hi).hex())
        f.write("#define NM_TE32_LN2LON %s\n" % float(l2n32_lo).hex())
        f.write("#define NM_TAB_POLY32 { %s }\n" % ", ".join(float(c).hex() for c in Dc))
        f.write("// kTabPoly: -2*B0..-2*B5 (log1p; the kernel multiplies by -r^2/2), C0..C3 (exp)\n")
        f.write("#define NM_TAB_POLY { %s }\n" % ", ".join(float(c).hex() for c in [-2 * b for b in B] + list(C)))
        f.write("// image layout (uint64 words): [0,256) log {invc, logc} pairs; [256,384) logctail;\n")
        f.write("// [384,640) exp {tail, bits(H) - (j << 45)} pairs; [640,704) 32-entry exp {tail, bits(H) - (j << 47)} pairs\n")
        f.write("#define NM_TAB_WORDS 704\n#define NM_TAB_IMAGE { \\\n")
        words = []
        for invc, logc, _ in lt:
            words += [hexd(invc), hexd(logc)]
        words += [hexd(t) for _, _, t in lt]
        for tail, sb in et:
            words += [hexd(tail), "0x%016xull" % sb]
        for tail, sb in et32:
            words += [hexd(tail), "0x%016xull" % sb]
        assert len(words) == 704
        for k in range(0, 704, 4):
            f.write("  " + ", ".join(words[k:k + 4]) + (", \\\n" if k + 4 < 704 else " }\n"))
    # ---- single-float pow
    invc, logc, expt, r0, r1 = pow32_tables()
    g32 = lambda t: (mp.log1p(t) - t) / t ** 2 if t != 0 else -mp.mpf(1) / 2
    w32 = lambda t: t                                        # error relative to r
    L32, el = remez.fit(g32, w32, r0 * 1.0001, r1 * 1.0001, 5, None, "f64", grid=600)
    s32 = float(mp.log(2) / 64) * 1.002
    E32, ee = remez.fit(ge, we, -s32, s32, 3, None, "f64", grid=600)
    with open(out, "a") as f:
        f.write("// ---- single-float pow: log1p(r) - r = r^2 L(r), r in [%.6g, %.6g], error relative to r 2^%.2f;\n" %
                (r0, r1, float(mp.log(el, 2))))
        f.write("// e^r - 1 - r = r^2 E(r), |r| <= %.6g, absolute error 2^%.2f\n" % (s32, float(mp.log(ee, 2))))
        f.write("#define NM_TF_OFF 0x%08xu\n" % OFF32)
        f.write("// kTabPF: L0..L4, E0..E2, ln2, 32/ln2, -ln2/32, 1.5*2^52, ln2 * 2^-20\n")
        consts = [float(c) for c in list(L32) + list(E32)] + [rd(mp.log(2)), rd(32 / mp.log(2)), rd(-mp.log(2) / 32), 1.5 * 2.0 ** 52, rd(mp.log(2)) * 2.0 ** -20]
        f.write("#define NM_TABF_POLY { %s }\n" % ", ".join(c.hex() for c in consts))
        f.write("// image (uint64 words): [0,64) {1/c, log c} pairs of doubles (one 16-byte read per element, no widening of a\n")
        f.write("// single-float 1/c); [64,96) bits(2^(j/32)) - (j << 47)\n")
        words = []
        for k in range(32):
            words += [hexd(invc[k]), hexd(logc[k])]
        words += ["0x%016xull" % v for v in expt]
        f.write("#define NM_TABF_WORDS 96\n#define NM_TABF_IMAGE { \\\n")
        for k in range(0, 96, 4):
            f.write("  " + ", ".join(words[k:k + 4]) + (", \\\n" if k + 4 < 96 else " }\n"))
    print("pow32: r in [%g, %g], log poly 2^%.2f, exp poly 2^%.2f" % (r0, r1, float(mp.log(el, 2)), float(mp.log(ee, 2))))
    print("wrote", os.path.normpath(out))
    print("log: r in [%g, %g], poly err 2^%.2f; exp: poly err 2^%.2f" %
          (rmin, rmax, float(mp.log(eb, 2)), float(mp.log(ec, 2))))


if __name__ == "__main__":
    main()
