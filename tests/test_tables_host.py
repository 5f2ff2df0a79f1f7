"""CPU: the lookup tables of the table-driven kernels (numericals_b200/csrc/nukmath_tables.inc).

 (1) the committed file is what tools/gen_tables.py generates (nothing hand-edited, generator and header in sync);
 (2) the properties the kernels rely on hold for the numbers in the file itself: r = z * invc - 1 fits in 53 bits
     over every sub-interval (so the device's single fma is exact), log c on the 2^-43 grid, |log c - (logc + tail)|
     tiny, 2^(j/128) = H_j (1 + tail_j), and the 32-entry single-float-pow tables likewise."""
import os
import re
import struct
import subprocess
import sys
from fractions import Fraction

import mpmath as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "numericals_b200", "csrc", "nukmath_tables.inc")
mp.mp.dps = 60


def u2d(u):
    return struct.unpack("<d", struct.pack("<Q", u))[0]


def words(name):
    txt = open(INC).read()
    m = re.search(r"#define %s \{(.*?)\}" % name, txt, re.S)
    return [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", m.group(1))]


def test_generated_file_is_current(tmp_path):
    out = str(tmp_path / "tables.inc")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py"), out], stdout=subprocess.DEVNULL)
    assert open(out).read() == open(INC).read(), "numericals_b200/csrc/nukmath_tables.inc is not what tools/gen_tables.py writes"


def test_double_tables_have_the_properties_the_kernels_use():
    w = words("NM_TAB_IMAGE")
    assert len(w) == 704
    off = int(re.search(r"#define NM_TL_OFF 0x([0-9a-f]+)ull", open(INC).read()).group(1), 16)
    for i in range(128):
        invc, logc, tail = Fraction(u2d(w[2 * i])), u2d(w[2 * i + 1]), u2d(w[256 + i])
        a, b = Fraction(u2d(off + (i << 45))), Fraction(u2d(off + ((i + 1) << 45)))
        ulp = Fraction(1, 2 ** 53) if b <= 1 else Fraction(1, 2 ** 52)
        q = ulp / invc.denominator                       # every z*invc - 1 of the interval is a multiple of q
        rmax = max(abs(a * invc - 1), abs(b * invc - 1))
        assert rmax < Fraction(1, 128) and rmax / q < 2 ** 53, i
        assert (Fraction(logc) * 2 ** 43).denominator == 1, i
        exact = mp.log(mp.mpf(invc.denominator) / invc.numerator)
        assert abs(exact - logc - tail) < mp.mpf(2) ** -96, i
        if a < 1 < b:
            assert invc == 1 and logc == 0.0 and tail == 0.0
    for j in range(128):
        tail, sb = u2d(w[384 + 2 * j]), w[384 + 2 * j + 1]
        H = u2d((sb + (j << 45)) & 0xFFFFFFFFFFFFFFFF)
        v = mp.mpf(2) ** (mp.mpf(j) / 128)
        assert 1.0 <= H < 2.0 and abs(v - mp.mpf(H)) <= mp.mpf(2) ** -53, j
        assert abs((v - mp.mpf(H)) / mp.mpf(H) - tail) < mp.mpf(2) ** -105, j


def test_exp32_table():
    """the 32-entry table of sinh / cosh: 2^(j/32) = H (1 + tail) to 2^-105"""
    w = words("NM_TAB_IMAGE")
    for j in range(32):
        tail, sb = u2d(w[640 + 2 * j]), w[640 + 2 * j + 1]
        H = u2d((sb + (j << 47)) & 0xFFFFFFFFFFFFFFFF)
        v = mp.mpf(2) ** (mp.mpf(j) / 32)
        assert 1.0 <= H < 2.0 and abs(v - mp.mpf(H)) <= mp.mpf(2) ** -53, j
        assert abs((v - mp.mpf(H)) / mp.mpf(H) - tail) < mp.mpf(2) ** -105, j


def test_single_float_pow_tables():
    w = words("NM_TABF_IMAGE")
    assert len(w) == 96
    off = int(re.search(r"#define NM_TF_OFF 0x([0-9a-f]+)u", open(INC).read()).group(1), 16)
    for i in range(32):
        c = Fraction(u2d(w[2 * i]))
        assert c.denominator <= 2 ** 12, i                # 24-bit z times a 12-bit 1/c: exact in double
        a = Fraction(struct.unpack("<f", struct.pack("<I", off + (i << 18)))[0])
        b = Fraction(struct.unpack("<f", struct.pack("<I", off + ((i + 1) << 18)))[0])
        assert max(abs(a * c - 1), abs(b * c - 1)) <= Fraction(1, 64), i
        assert abs(-mp.log(mp.mpf(c.numerator) / c.denominator) - u2d(w[2 * i + 1])) <= abs(mp.mpf(u2d(w[2 * i + 1]))) * mp.mpf(2) ** -53 + mp.mpf(10) ** -300
        H = u2d((w[64 + i] + (i << 47)) & 0xFFFFFFFFFFFFFFFF)
        assert abs(mp.mpf(2) ** (mp.mpf(i) / 32) - mp.mpf(H)) <= mp.mpf(2) ** -53
