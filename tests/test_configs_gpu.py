"""Parity at the FULL sizes of BASELINE.json's configs (the other GPU tests use shapes the oracle
finishes in milliseconds):

  C2  nu:exp nu:sin nu:log nu:tanh on 2^28 single-floats: <= 1 ULP from the oracle on EVERY element
  C3  (8192 x 1) + (1 x 8192) and (4096 x 4096) * (4096), OUT preallocated: bit-exact
      (benchmarks/two-arg-fn.lisp:50-81 shapes; src/basic-math/two-arg-fn-non-comparison.lisp:151-186)
  C4  16384 x 16384 double-float sum / maximum / mean: axis 0 bit-exact in the reference's row-accumulate
      order (src/basic-math/sum.lisp:52-71), axis 1 and full against the exactly rounded sum
  C5  nu:multiply + nu:sum over 2^32 + 12345 single-floats of RANDOM data: every 32 Mi-element window that
      matters for 64-bit indexing (first, the ones straddling 2^31 and 2^32, the ragged tail, and a spread of
      others) bit-exact against oracle BMAS_smul, the whole array bit-exact against an independent IEEE
      multiply, the sum inside the pairwise bound, and an int32 sum over the same length bit-exact (a
      skipped or re-read range cannot hide in a wrap-around integer sum) -- src/basic-math/sum.lisp:117-127.
"""
import ctypes as C
import math
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAX_ULP = 1  # north-star tolerance for the transcendentals, in units in the last place of the oracle's result


def _ulp32(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ka = np.where(ia < 0, -(1 << 31) - ia, ia)
    kb = np.where(ib < 0, -(1 << 31) - ib, ib)
    d = np.abs(ka - kb)
    return np.where(np.isnan(a) & np.isnan(b), 0, d)


def _oracle_unary_threaded(oracle, name, x, threads=16):
    """the oracle over equal contiguous slices on host threads (ctypes releases the GIL)"""
    out = np.empty_like(x)
    f = getattr(oracle, "BMAS_" + name)
    f.restype = None
    n = x.size
    per = -(-n // threads)

    def work(t):
        lo, hi = t * per, min(n, (t + 1) * per)
        if lo < hi:
            f(C.c_long(hi - lo), C.c_void_p(x.ctypes.data + lo * x.itemsize), C.c_long(1),
              C.c_void_p(out.ctypes.data + lo * x.itemsize), C.c_long(1))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return out


# ------------------------------------------------------------------------------------------------ C2
def test_c2_four_ufuncs_2p28_every_element_within_1ulp(nu, oracle):
    n = 1 << 28
    x = np.random.default_rng(0).random(n, dtype=np.float32)      # U[0,1) seed 0: the reference's test domain
    dx = nu.asarray(x)
    dout = nu.empty((n,), "single-float")
    for fn in ("exp", "sin", "log", "tanh"):
        getattr(nu, fn)(dx, out=dout, broadcast=False)
        got = dout.to_numpy()
        want = _oracle_unary_threaded(oracle, "s" + fn, x)
        worst = 0
        for lo in range(0, n, 1 << 24):
            worst = max(worst, int(_ulp32(got[lo:lo + (1 << 24)], want[lo:lo + (1 << 24)]).max()))
        assert worst <= MAX_ULP, (fn, worst)


# ------------------------------------------------------------------------------------------------ C3
def test_c3_broadcast_shapes_bit_exact(nu, refnu):
    rng = np.random.default_rng(0)
    col = rng.random((8192, 1), dtype=np.float32)
    row = rng.random((1, 8192), dtype=np.float32)
    out = nu.empty((8192, 8192), "single-float")
    r = nu.add(nu.asarray(col), nu.asarray(row), out=out)
    assert r is out
    want = refnu.two_arg("add", col, row)                         # 8192 per-row oracle calls, the reference's loop
    assert out.to_numpy().tobytes() == want.tobytes()
    m = rng.random((4096, 4096), dtype=np.float32)
    v = rng.random(4096, dtype=np.float32)
    out = nu.empty((4096, 4096), "single-float")
    nu.multiply(nu.asarray(m), nu.asarray(v), out=out)
    assert out.to_numpy().tobytes() == refnu.two_arg("mul", m, v).tobytes()
    # in place on the matrix operand (nu:multiply! a v): OUT aliases X at full size
    dm = nu.asarray(m)
    nu.multiply(dm, nu.asarray(v), out=dm)
    assert dm.to_numpy().tobytes() == refnu.two_arg("mul", m, v).tobytes()


# ------------------------------------------------------------------------------------------------ C4
def test_c4_reductions_16384sq_f64(nu, refnu, oracle):
    n = 16384
    h = np.random.default_rng(0).random((n, n))                   # U[0,1) seed 0, 2 GiB
    x = nu.asarray(h)
    # axis 0: the reference's out <- out + row_i order, bit for bit (TMA column kernel is the default here)
    assert nu.sum(x, axes=0).to_numpy().tobytes() == refnu.reduce("sum", h, 0).tobytes()
    assert nu.maximum(x, axes=0).to_numpy().tobytes() == refnu.reduce("maximum", h, 0).tobytes()
    assert nu.mean(x, axes=0).to_numpy().tobytes() == refnu.mean(h, 0).tobytes()
    # axis 1 and full: deterministic tree, checked against the exactly rounded sum with the pairwise bound
    eps = 2.0 ** -53
    exact_rows = np.array([oracle.ORACLE_dsum_exact_d(C.c_long(n), C.c_void_p(h.ctypes.data + 8 * n * i), C.c_long(1))
                           for i in range(n)])
    got = nu.sum(x, axes=1).to_numpy()
    bound = 2 * eps * math.log2(n) * np.sum(np.abs(h), axis=1)
    assert np.all(np.abs(got - exact_rows) <= bound)
    assert np.array_equal(nu.maximum(x, axes=1).to_numpy(), h.max(axis=1))
    gm = nu.mean(x, axes=1).to_numpy()
    assert np.array_equal(gm, got / np.float64(n))                # TRUE division of our own sum (mean.lisp:90-97)
    exact = oracle.ORACLE_dsum_exact_d(C.c_long(n * n), C.c_void_p(h.ctypes.data), C.c_long(1))
    tot = float(nu.sum(x))
    assert abs(tot - exact) <= 2 * eps * math.log2(n * n) * float(np.abs(h).sum()), (tot, exact)
    assert float(nu.maximum(x)) == h.max() and float(nu.minimum(x)) == h.min()
    assert float(nu.mean(x)) == tot / float(n * n)
    assert tot == float(nu.sum(x)), "full sum is not deterministic run to run"


# ------------------------------------------------------------------------------------------------ C5
W = 1 << 25   # 32 Mi-element window


def _windows(n):
    ws = [0, (1 << 31) - W // 2, (1 << 31), (1 << 32) - W // 2 - 3, (1 << 32) - 7, n - W, n - 12345 - 5]
    ws += [int(k * (n - W) / 9.0) | 1 for k in range(1, 9)]      # a spread, odd starts (misaligned windows)
    return sorted(set(max(0, min(w, n - 1)) for w in ws))


def test_c5_multiply_sum_beyond_2p32_random_data(nu, oracle):
    torch = pytest.importorskip("torch")
    from numericals_b200.array import from_pointer
    n = (1 << 32) + 12345
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 13 * n + (4 << 30):
        pytest.skip("needs %d GiB of HBM" % ((13 * n) >> 30))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    xs = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen).mul_(2).sub_(1)     # U[-1,1)
    ys = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen).add_(0.5)           # U[0.5,1.5)
    zs = torch.zeros(n, dtype=torch.float32, device="cuda")
    X, Y, Z = (from_pointer(t.data_ptr(), "single-float", (n,), owner=t) for t in (xs, ys, zs))
    torch.cuda.synchronize()
    r = nu.multiply(X, Y, out=Z, broadcast=False)
    assert r is Z
    total = nu.sum(Z)
    torch.cuda.synchronize()
    # (1) windows against the oracle's BMAS_smul, bit for bit
    smul = oracle.BMAS_smul
    smul.restype = None
    for w0 in _windows(n):
        m = min(W, n - w0)
        hx, hy = xs[w0:w0 + m].cpu().numpy(), ys[w0:w0 + m].cpu().numpy()
        want = np.empty_like(hx)
        smul(C.c_long(m), C.c_void_p(hx.ctypes.data), C.c_long(1), C.c_void_p(hy.ctypes.data), C.c_long(1),
             C.c_void_p(want.ctypes.data), C.c_long(1))
        assert zs[w0:w0 + m].cpu().numpy().tobytes() == want.tobytes(), "window at %d" % w0
    # (2) every element against an independent IEEE multiply, and the exact sum in double, chunk by chunk
    exact, sabs = 0.0, 0.0
    chunk = 1 << 28
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ref = xs[lo:hi] * ys[lo:hi]
        assert torch.equal(ref, zs[lo:hi]), "chunk at %d" % lo
        d = ref.double()
        exact += float(d.sum().item())
        sabs += float(d.abs().sum().item())
        del ref, d
    bound = 2 * 2.0 ** -24 * math.log2(n) * sabs
    assert abs(float(total) - exact) <= bound, (float(total), exact, bound)
    assert float(nu.sum(Z)) == float(total), "not deterministic"
    # slab sums add up: the full-array kernel and the per-slab kernels see the same elements
    slab = 1 << 30
    parts = [float(nu.sum(from_pointer(zs.data_ptr() + 4 * lo, "single-float", (min(slab, n - lo),), owner=zs)))
             for lo in range(0, n, slab)]
    assert abs(math.fsum(parts) - exact) <= bound
    # fused form (BMAS_sdot: multiply + sum in one pass, 8 instead of 16 B/element) over the same operands
    dot = float(nu.vdot(X, Y))
    assert abs(dot - exact) <= bound, (dot, exact)
    del zs, Z
    # (3) wrap-around int32 sum over > 2^32 elements: exact, so any index slip shows
    xi = xs.view(torch.int32)                                      # the float bit patterns as integers
    XI = from_pointer(xi.data_ptr(), "(signed-byte 32)", (n,), owner=xi)
    want = 0
    for lo in range(0, n, chunk):
        want += int(xi[lo:min(n, lo + chunk)].sum(dtype=torch.int64).item())
    want = (want + (1 << 31)) % (1 << 32) - (1 << 31)
    assert int(nu.sum(XI)) == want
    assert int(nu.maximum(XI)) == int(xi.max().item()) and int(nu.minimum(XI)) == int(xi.min().item())
    # arg-maximum beyond 2^32: plant the extremum past the 32-bit boundary
    ys[(1 << 32) + 4321] = 7.0
    torch.cuda.synchronize()
    assert int(nu.arg_maximum(Y)) == (1 << 32) + 4321
