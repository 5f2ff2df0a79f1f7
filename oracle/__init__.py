"""oracle/ -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline /
--impl reference).  The product package numericals_b200 never imports this.

Loader for the CPU restatement of the BMAS C contract (oracle/bmas_oracle.c).  Parity status:
PINNED at the level the reference itself pins this path -- its known-answer tests
(sum.lisp:236-283, maximum.lisp:237-290, mean.lisp:242-289, variance.lisp:144-193,
n-arg-fn-tests.lisp) are ported verbatim to tests/golden/reference_kats.json and checked against
this oracle in tests/test_oracle_kats.py.  The reference cannot be executed here (Common Lisp, no
implementation in the image; BMAS un-vendored), so there is no oracle/_ref.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")


def build(force=False):
    """Compile bmas_oracle.c (gcc) into oracle/_build/liboracle_{avx2,avx512}.so."""
    avx2 = os.path.join(BUILD_DIR, "liboracle_avx2.so")
    if force or not os.path.exists(avx2) or os.path.getmtime(avx2) < os.path.getmtime(
            os.path.join(_HERE, "bmas_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-j2"])
    return avx2


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line.split()
                    return all(x in flags for x in ("avx512f", "avx512dq", "avx512bw", "avx512vl"))
    except OSError:
        pass
    return False


_handles = {}


def load(widest=False):
    """The parity oracle (fixed AVX2+FMA SLEEF entry points).  widest=True returns the AVX-512
    build when the host supports it: ONLY for CPU-baseline timing, never for parity."""
    key = "avx512" if (widest and _cpu_has_avx512()) else "avx2"
    if key not in _handles:
        build()
        path = os.path.join(BUILD_DIR, "liboracle_%s.so" % key)
        if not os.path.exists(path):
            key, path = "avx2", os.path.join(BUILD_DIR, "liboracle_avx2.so")
        h = C.CDLL(path, mode=C.RTLD_LOCAL)
        h.ORACLE_dsum_exact_d.restype = C.c_double
        h.ORACLE_ssum_exact_d.restype = C.c_double
        h._oracle_kind = key
        _handles[key] = h
    return _handles[key]
