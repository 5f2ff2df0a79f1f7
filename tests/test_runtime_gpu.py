"""Runtime properties of libnuk the reference's call pattern depends on:

* re-entrancy: the reference calls the C layer from concurrent lparallel workers on disjoint slices
  (src/transcendental/lparallel.lisp:70-96, dense-numericals-src/utils/lparallel.lisp:54-76; its test ladder
  forces that path, src/basic-math/test-definition-macros.lisp:30,37) -- eight host threads, device and host
  pointers, results bit-equal to the serial run;
* host operands as a Lisp vector really is: PAGEABLE memory, any stride, several chunks -- and page-locked
  memory (nuk_host_alloc) on the direct-DMA path;
* CUDA-graph replay of a captured call sequence; NUK_PTR_DEVICE pointer mode;
* hmax / hmin NaN placement against the oracle's x[0]-seeded fold (oracle/bmas_oracle.c DEF_HMINMAX).
"""
import ctypes as C
import threading

import numpy as np
import pytest

from _helpers import NP_OF, call_bmas, rnd

pytestmark = pytest.mark.gpu


def _same(a, b):
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        return np.array_equal(na, nb) and np.where(na, 0, a).tobytes() == np.where(nb, 0, b).tobytes()
    return a.tobytes() == b.tobytes()


def test_eight_host_threads_disjoint_slices_bit_equal_to_serial(nu, libnuk):
    from numericals_b200 import _lib
    n, T = 1 << 22, 8
    per = n // T
    rng = np.random.default_rng(3)
    xs, ys = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    xd = rng.random(n)
    dx, dy, dxd = nu.asarray(xs), nu.asarray(ys), nu.asarray(xd)
    sadd, ssin, dsum = _lib.bmas("sadd"), _lib.bmas("ssin"), _lib.bmas("dsum")

    def run(threaded, host):
        o_add, o_sin = np.zeros(n, np.float32), np.zeros(n, np.float32)
        d_add, d_sin = nu.zeros(n, type="single-float"), nu.zeros(n, type="single-float")
        sums = [None] * T
        errs = []

        def work(t):
            try:
                lo = t * per
                if host:
                    px, py = xs.ctypes.data + 4 * lo, ys.ctypes.data + 4 * lo
                    pa, ps, pd = o_add.ctypes.data + 4 * lo, o_sin.ctypes.data + 4 * lo, xd.ctypes.data + 8 * lo
                else:
                    px, py = dx.ptr + 4 * lo, dy.ptr + 4 * lo
                    pa, ps, pd = d_add.ptr + 4 * lo, d_sin.ptr + 4 * lo, dxd.ptr + 8 * lo
                for _ in range(3):
                    call_bmas(sadd, per, px, 1, py, 1, pa, 1)
                    call_bmas(ssin, per, px, 1, ps, 1)
                    sums[t] = dsum(C.c_long(per), C.c_void_p(pd), C.c_long(1))
                assert libnuk.nuk_last_status() == 0, libnuk.nuk_last_error()
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        if threaded:
            ts = [threading.Thread(target=work, args=(t,)) for t in range(T)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        else:
            for t in range(T):
                work(t)
        assert not errs, errs
        libnuk.nuk_device_sync()
        if host:
            return o_add, o_sin, sums
        return d_add.to_numpy(), d_sin.to_numpy(), sums

    for host in (False, True):
        a0, s0, r0 = run(False, host)
        a1, s1, r1 = run(True, host)
        assert a0.tobytes() == a1.tobytes() and s0.tobytes() == s1.tobytes() and r0 == r1, host
        assert np.array_equal(a0, xs + ys)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_operands_many_chunks_every_stride(pinned, nu, libnuk, oracle):
    """1 MiB chunks so that a 3-slot ring wraps several times; pageable (numpy) and page-locked buffers."""
    from numericals_b200 import _lib
    old = libnuk.nuk_get_option(b"stage_mb")
    libnuk.nuk_set_option(b"stage_mb", 1)
    held = []

    def buf(a):
        if not pinned:
            return a
        p = C.c_void_p()
        _lib.check(libnuk.nuk_host_alloc(C.byref(p), a.nbytes))
        held.append(p)
        v = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(a.nbytes,)).view(a.dtype)
        v[:] = a
        return v

    try:
        rng = np.random.default_rng(9)
        for p, n, (ix, iy, io) in (("s", 3_000_001, (1, 1, 1)), ("d", 1_000_003, (1, -1, 1)), ("s", 700_001, (3, -2, 2)),
                                   ("i16", 2_000_001, (-1, 1, -1)), ("u8", 5_000_001, (2, 1, 3)), ("d", 400_001, (1, 0, -3))):
            dt = NP_OF[p]
            ln = n * max(abs(ix), abs(iy), abs(io), 1)
            x, y = buf(rnd(rng, dt, ln)), buf(rnd(rng, dt, ln))
            init = rnd(rng, dt, ln)
            want, got = init.copy(), buf(init.copy())

            def first(inc):
                return 0 if inc >= 0 else (n - 1) * -inc

            for f, o in ((getattr(oracle, "BMAS_" + p + "mul"), want), (_lib.bmas(p + "mul"), got)):
                f.restype = None
                call_bmas(f, n, x.ctypes.data + first(ix) * x.itemsize, ix, y.ctypes.data + first(iy) * x.itemsize, iy,
                          o.ctypes.data + first(io) * x.itemsize, io)
            assert libnuk.nuk_last_status() == 0, libnuk.nuk_last_error()
            assert _same(np.asarray(got), want), (p, n, ix, iy, io, pinned)     # incl. the untouched gaps of a strided OUT
            # reductions / dot / arg-reduction over the same host operands
            red = "sum" if p in ("s", "d", "i16") else "hmax"
            fo, fd = getattr(oracle, "BMAS_%s%s" % (p, red)), _lib.bmas(p + red)
            fo.restype = fd.restype
            a = fo(C.c_long(n), C.c_void_p(x.ctypes.data + first(ix) * x.itemsize), C.c_long(ix))
            b = fd(C.c_long(n), C.c_void_p(x.ctypes.data + first(ix) * x.itemsize), C.c_long(ix))
            if p in ("s", "d") and red == "sum":
                assert abs(a - b) <= 1e-5 * abs(a) + 1e-3, (p, a, b)
            else:
                assert a == b, (p, red, a, b)
            fo, fd = getattr(oracle, "BMAS_%shimax" % p), _lib.bmas(p + "himax")
            fo.restype = C.c_long
            assert fo(C.c_long(n), C.c_void_p(x.ctypes.data + first(ix) * x.itemsize), C.c_long(ix)) == \
                fd(C.c_long(n), C.c_void_p(x.ctypes.data + first(ix) * x.itemsize), C.c_long(ix)), p
            if p in ("i16",):
                fo, fd = getattr(oracle, "BMAS_%sdot" % p), _lib.bmas(p + "dot")
                fo.restype = fd.restype
                args = (C.c_long(n), C.c_void_p(x.ctypes.data + first(ix) * x.itemsize), C.c_long(ix),
                        C.c_void_p(y.ctypes.data + first(iy) * x.itemsize), C.c_long(iy))
                assert fo(*args) == fd(*args)
    finally:
        libnuk.nuk_set_option(b"stage_mb", old)
        for p in held:
            libnuk.nuk_host_free(p)
    assert libnuk.nuk_host_mover_threads() >= 1


def test_graph_replay_and_pointer_mode(nu, libnuk):
    from numericals_b200 import _lib
    n = 10 ** 6
    rng = np.random.default_rng(4)
    a, b = rng.random(n), rng.random(n)
    da, db, do = nu.asarray(a), nu.asarray(b), nu.zeros(n, type="double-float")
    dadd, dmul = _lib.bmas("dadd"), _lib.bmas("dmul")
    s = C.c_void_p()
    _lib.check(libnuk.nuk_stream_create(C.byref(s)))
    prev = libnuk.nuk_get_stream()
    try:
        libnuk.nuk_set_stream(s)
        _lib.check(libnuk.nuk_set_pointer_mode(1))          # NUK_PTR_DEVICE: no driver query per operand
        assert libnuk.nuk_get_pointer_mode() == 1
        call_bmas(dadd, n, da.ptr, 1, db.ptr, 1, do.ptr, 1)                        # warm-up outside capture
        _lib.check(libnuk.nuk_graph_begin(s))
        call_bmas(dadd, n, da.ptr, 1, db.ptr, 1, do.ptr, 1)                        # out = a + b
        call_bmas(dmul, n, do.ptr, 1, db.ptr, 1, do.ptr, 1)                        # out = out * b  (in place)
        g = C.c_void_p()
        _lib.check(libnuk.nuk_graph_end(s, C.byref(g)))
        for _ in range(5):
            _lib.check(libnuk.nuk_graph_launch(g, s))
        _lib.check(libnuk.nuk_stream_sync(s))
        assert np.array_equal(do.to_numpy(), (a + b) * b)
        _lib.check(libnuk.nuk_graph_destroy(g))
        # a stride-0 operand may still be the reference's one-element HOST scalar in device mode
        two = np.array([2.0])
        call_bmas(dmul, n, da.ptr, 1, two.ctypes.data, 0, do.ptr, 1)
        _lib.check(libnuk.nuk_stream_sync(s))
        assert np.array_equal(do.to_numpy(), a * 2.0)
    finally:
        libnuk.nuk_set_pointer_mode(0)
        libnuk.nuk_set_stream(prev)
        libnuk.nuk_stream_destroy(s)


@pytest.mark.parametrize("p", ["s", "d"])
def test_hmax_hmin_nan_placement_matches_the_oracle(p, nu, libnuk, oracle):
    """BMAS hmax/hmin fold from x[0]: a NaN in the first position is the result, a NaN anywhere else is
    ignored (oracle/bmas_oracle.c DEF_HMINMAX).  Device pointers, host pointers, strided, multi-chunk."""
    from numericals_b200 import _lib
    dt = NP_OF[p]
    rng = np.random.default_rng(8)
    old = libnuk.nuk_get_option(b"stage_mb")
    libnuk.nuk_set_option(b"stage_mb", 1)
    try:
        for n in (1, 5, 4097, 700_001):
            for where in ("none", "first", "middle", "last", "first+last", "all"):
                h = rnd(rng, dt, n, -5, 5)
                idx = {"none": [], "first": [0], "middle": [n // 2], "last": [n - 1], "first+last": [0, n - 1],
                       "all": range(n)}[where]
                for i in idx:
                    h[i] = np.nan
                d = nu.asarray(h)
                for red in ("hmax", "hmin"):
                    fo, fd = getattr(oracle, "BMAS_" + p + red), _lib.bmas(p + red)
                    fo.restype = fd.restype
                    want = fo(C.c_long(n), C.c_void_p(h.ctypes.data), C.c_long(1))
                    for ptr in (d.ptr, h.ctypes.data):
                        got = fd(C.c_long(n), C.c_void_p(ptr), C.c_long(1))
                        assert (np.isnan(want) and np.isnan(got)) or want == got, (p, red, n, where, want, got)
                # the nu: API full reduction is the same hmax over the flattened array (maximum.lisp:65)
                fmax = getattr(oracle, "BMAS_" + p + "hmax")
                got, want = nu.maximum(d), fmax(C.c_long(n), C.c_void_p(h.ctypes.data), C.c_long(1))
                assert (np.isnan(want) and np.isnan(got)) or want == got, (p, n, where)
            # reversed view: logical element 0 is the LAST address
            h = rnd(rng, dt, n, -5, 5)
            h[-1] = np.nan
            fo, fd = getattr(oracle, "BMAS_" + p + "hmax"), _lib.bmas(p + "hmax")
            last = h.ctypes.data + (n - 1) * h.itemsize
            want, got = fo(C.c_long(n), C.c_void_p(last), C.c_long(-1)), fd(C.c_long(n), C.c_void_p(last), C.c_long(-1))
            assert np.isnan(want) and np.isnan(got), (p, n)
    finally:
        libnuk.nuk_set_option(b"stage_mb", old)


def test_thread_moving_between_devices_keeps_its_scratch_apart(nu, libnuk):
    """A thread that switches device (nuk_init) must get that device's scratch / result slots: the
    single-process multi-GPU mode does this once per call."""
    if libnuk.nuk_device_count() < 2:
        pytest.skip("needs two devices")
    from numericals_b200 import _lib
    h = np.random.default_rng(1).random((300, 16384))
    try:
        for dev in (0, 1, 0, 1):
            _lib.check(libnuk.nuk_init(dev))
            x = nu.asarray(h)
            assert abs(float(nu.sum(x)) - h.sum()) < 1e-6 * h.size
            assert np.allclose(nu.sum(x, axes=0).to_numpy(), h.sum(axis=0), rtol=1e-12)   # TMA kernel: per-device opt-in
            assert int(nu.arg_maximum(x)) == int(h.argmax())
            assert float(nu.vdot(x, x)) == pytest.approx(float((h * h).sum()), rel=1e-12)
    finally:
        libnuk.nuk_init(0)
