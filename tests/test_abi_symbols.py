"""CPU checks of the drop-in boundary: libnuk.so loads without a GPU, exports EVERY symbol that
include/nuk.h + include/nuk_bmas.h declare, and the oracle exports the same BMAS symbol set."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = subprocess.check_output(["gcc", "-E", "-P", os.path.join(ROOT, "include", "nuk.h")]).decode()
    names = set(re.findall(r"\b(BMAS_\w+|nuk_\w+)\s*\(", src))
    names.discard("nuk_stream_t")
    return sorted(names)


def test_header_declares_the_reference_table():
    names = declared_symbols()
    bmas = [n for n in names if n.startswith("BMAS_")]
    # 24 arith + 20 max/min + 60 compare + 14 abs/round + 2 sqrt + 24 copy/cast + 28 transcendental
    # + 4 pow/atan2 + 26 reductions + 20 arg-reductions + 6 dot + 40 bitwise + sra
    assert len(bmas) == 269, len(bmas)
    for must in ("BMAS_sadd", "BMAS_dadd", "BMAS_i8add", "BMAS_u8mul", "BMAS_slt", "BMAS_u64gt", "BMAS_ssin",
                 "BMAS_dpow", "BMAS_cast_ds", "BMAS_cast_u16d", "BMAS_ssum", "BMAS_u8hmax", "BMAS_i64sra",
                 "BMAS_dhimax", "BMAS_i16dot"):
        assert must in bmas


def test_libnuk_exports_every_declared_symbol():
    from numericals_b200 import _lib
    lib = _lib.lib()          # must load on a machine without a GPU
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.nuk_version()


def test_oracle_exports_the_same_bmas_symbols(oracle):
    missing = [n for n in declared_symbols() if n.startswith("BMAS_") and not hasattr(oracle, n)]
    assert not missing, missing


def test_no_device_means_loud_failure_not_fallback():
    """Without a CUDA device every compute entry point must fail (no CPU path)."""
    from numericals_b200 import _lib
    lib = _lib.lib()
    if lib.nuk_device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    x = np.ones(8, dtype=np.float32)
    o = np.zeros(8, dtype=np.float32)
    lib.nuk_clear_error()
    _lib.bmas("sadd")(C.c_long(8), C.c_void_p(x.ctypes.data), C.c_long(1), C.c_void_p(x.ctypes.data), C.c_long(1),
                      C.c_void_p(o.ctypes.data), C.c_long(1))
    assert lib.nuk_last_status() != 0
    assert not o.any(), "output was written without a device: a CPU path exists"
    import numericals_b200 as nu
    with pytest.raises(nu.NukError):
        nu.zeros(4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "numericals_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_lisp_stub_binds_exported_symbols_and_defines_what_it_exports():
    """lisp/nuk-cffi.lisp cannot be executed here (no Lisp in the image); at least every C name it binds
    must be exported by libnuk.so and every symbol its package exports must have a definition."""
    from numericals_b200 import _lib
    lib = _lib.lib()
    text = open(os.path.join(ROOT, "lisp", "nuk-cffi.lisp")).read()
    cnames = re.findall(r'\(cffi:defcfun \("(\w+)"', text)
    assert len(cnames) >= 40
    missing = [n for n in cnames if not hasattr(lib, n)]
    assert not missing, missing
    export = re.search(r"\(:export(.*?)\)\)\s*\(in-package", text, re.S).group(1)
    exported = set(re.findall(r"#:([^\s()]+)", export))
    defined = set(re.findall(r"\((?:defun|defmacro|defconstant|defvar|define-condition)\s+([^\s()]+)", text))
    defined |= set(re.findall(r'\(cffi:defcfun \("\w+"\s+([^\s()]+)\)', text))
    for name in re.findall(r"\(defstruct \(([^\s()]+)", text):
        defined |= {name, "make-" + name, name + "-pointer"}
    undefined = sorted(exported - defined)
    assert not undefined, undefined
    # the enum values the stub hard-codes are the header's
    hdr = subprocess.check_output(["gcc", "-E", "-P", os.path.join(ROOT, "include", "nuk.h")]).decode()
    for lisp, c in (("+logb+ 13", "NUK_LOGB"), ("+not+ 21", "NUK_NOT"), ("+sumsq+ 3", "NUK_SUMSQ"), ("+u8+ 10", "NUK_U8")):
        assert ("(defconstant " + lisp + ")") in text, lisp
    assert "NUK_LOGB" in hdr and "NUK_PTR_DEVICE = 1" in hdr
