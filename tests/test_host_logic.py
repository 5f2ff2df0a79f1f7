"""Host-side logic that needs no GPU: translation table, max-type, broadcast resolver and
descriptor compaction (evaluated by libnuk's own resolver code), staging arithmetic."""
import ctypes as C

import numpy as np
import pytest

import numericals_b200 as nu
from numericals_b200 import _lib as L
from numericals_b200 import dtypes as T


def test_translation_table_matches_reference_rows():
    # src/basic-math/package.lisp:131-140: unsigned add/sub use the signed symbols, mul does not
    assert nu.c_name("(unsigned-byte 32)", "add") == "i32add"
    assert nu.c_name("(unsigned-byte 32)", "subtract") == "i32sub"
    assert nu.c_name("(unsigned-byte 32)", "multiply") == "u32mul"
    assert nu.c_name("(unsigned-byte 8)", "sum") == "i8sum"
    assert nu.c_name("(unsigned-byte 8)", "maximum") == "u8hmax"
    assert nu.c_name("single-float", "abs") == "sfabs" and nu.c_name("(signed-byte 16)", "abs") == "i16abs"
    assert nu.c_name("double-float", "expt") == "dpow" and nu.c_name("single-float", "two-arg-<") == "slt"
    with pytest.raises(nu.NoCFunction):
        nu.c_name("(signed-byte 32)", "divide")      # package.lisp:140: floats only
    with pytest.raises(nu.NoCFunction):
        nu.c_name("(signed-byte 32)", "sin")
    with pytest.raises(nu.NoCFunction):
        nu.c_name("single-float", "lognot")


def test_every_table_entry_is_exported(libnuk):
    for table in (T.TWO_ARG, T.COMPARE, T.ONE_ARG, T.REDUCE):
        for name in table:
            for et in T.ELEMENT_TYPES:
                try:
                    sym = nu.c_name(et, name)
                except nu.NoCFunction:
                    continue
                assert hasattr(libnuk, "BMAS_" + sym), (name, et, sym)


@pytest.mark.parametrize("a,b,want", [
    ("single-float", "double-float", "double-float"),
    ("(signed-byte 32)", "single-float", "single-float"),
    ("(unsigned-byte 8)", "(signed-byte 8)", "(signed-byte 16)"),
    ("(unsigned-byte 16)", "(signed-byte 8)", "(signed-byte 32)"),
    ("(unsigned-byte 32)", "(signed-byte 32)", "(signed-byte 64)"),
    ("(unsigned-byte 64)", "(signed-byte 64)", "single-float"),
    ("(unsigned-byte 8)", "(unsigned-byte 32)", "(unsigned-byte 32)"),
    ("(unsigned-byte 8)", "(signed-byte 64)", "(signed-byte 64)"),
    ("(signed-byte 8)", "(signed-byte 16)", "(signed-byte 16)"),
])
def test_max_type(a, b, want, refnu):
    assert nu.max_type(a, b) == want and nu.max_type(b, a) == want       # common.lisp:254-280
    assert refnu.max_type(a, b) == want


def test_broadcast_rule_against_numpy_and_oracle(refnu):
    cases = [((8192, 1), (1, 8192)), ((4096, 4096), (4096,)), ((10, 1, 10), (10, 45, 10), (2, 10, 45, 10)),
             ((45, 45), (10, 45, 45)), ((), (3,)), ((3,), (4,)), ((2, 3), (3, 2)), ((1,), (1,)), ((5, 1, 7), (6, 1))]
    for shapes in cases:
        ok, dims = nu.broadcast_compatible_p(*shapes)
        rok, rdims = refnu.broadcast_compatible_p(*shapes)
        try:
            want = np.broadcast_shapes(*shapes)
            assert ok and rok and dims == tuple(want) == rdims, shapes
        except ValueError:
            assert not ok and not rok, shapes


def test_strides_for_broadcast(refnu):
    assert nu.strides_for_broadcast((8192, 1), (8192, 8192)) == (1, 0)
    assert nu.strides_for_broadcast((1, 8192), (8192, 8192)) == (0, 1)
    assert nu.strides_for_broadcast((4096,), (4096, 4096)) == (0, 1)
    assert nu.strides_for_broadcast((10, 45, 10), (2, 10, 45, 10)) == (0, 450, 10, 1)
    assert refnu.strides_for_broadcast((10, 45, 10), (2, 10, 45, 10)) == (0, 450, 10, 1)
    with pytest.raises(nu.IncompatibleBroadcastDimensions):
        nu.strides_for_broadcast((3,), (4,))


def _coalesce(dims, *strides):
    lib = L.lib()
    d = L.i64_array(list(dims) + [0] * (8 - len(dims)))
    st = L.i64_array(sum(([*s] + [0] * (8 - len(s)) for s in strides), []))
    r = lib.nuk_coalesce(len(strides), len(dims), d, st)
    return [d[i] for i in range(r)], [[st[o * 8 + i] for i in range(r)] for o in range(len(strides))]


def test_descriptor_compaction():
    # contiguous 3-d collapses to one axis
    assert _coalesce((4, 5, 6), (30, 6, 1), (30, 6, 1)) == ([120], [[1], [1]])
    # (4096 x 4096) * (4096): the row vector's (0, 1) strides block the merge
    d, s = _coalesce((4096, 4096), (4096, 1), (4096, 1), (0, 1))
    assert d == [4096, 4096] and s == [[4096, 1], [4096, 1], [0, 1]]
    # size-1 axes vanish; a scalar operand (all strides 0) never blocks
    d, s = _coalesce((1, 8, 1, 16), (0, 16, 0, 1), (0, 0, 0, 0))
    assert d == [128] and s == [[1], [0]]
    # slice with a step stays separate
    d, s = _coalesce((10, 5), (20, 2), (5, 1))
    assert d == [10, 5]


def test_view_descriptors_are_pure_metadata():
    """aref* / transpose / reshape / broadcast are descriptor operations (no device needed)."""
    from numericals_b200.array import DenseArray

    class FakeStorage:
        ptr, nbytes = 4096, 4 * 100 * 60
    a = DenseArray(FakeStorage(), "single-float", (100, 60))
    v = a[::-2, 10:50:3]
    assert v.dimensions == (50, 14) and v.strides == (-120, 3) and v.offset == 99 * 60 + 10
    t = a.transpose()
    assert t.dimensions == (60, 100) and t.strides == (1, 60)
    b = DenseArray(FakeStorage(), "single-float", (60,)).broadcast_to((100, 60))
    assert b.strides == (0, 1)
    assert a.reshape(60, 100).strides == (100, 1)
    assert v.ptr == 4096 + 4 * v.offset
