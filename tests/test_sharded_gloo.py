"""N > 1 path on CPU: world_size-2 gloo processes exercise the slab partition (flat and leading-axis),
operand slicing / replication for broadcast maps, and the local-reduce + rank-ordered-combine protocol of
numericals_b200.sharded with an oracle-backed local backend (test infrastructure: the product backend is
libnuk on the rank's GPU with the exchange inside libnuk's kernels, covered by tests/test_multigpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleBackend:
    """local compute through the CPU oracle, exchange through torch.distributed (tests only)"""

    _OPS = {"add": "add", "subtract": "sub", "multiply": "mul", "divide": "div"}
    _FOLD = {"sum": "add", "maximum": "max", "minimum": "min"}

    def __init__(self, dist, world):
        from oracle import reference_nu
        self.ref, self.dist, self.world = reference_nu, dist, world

    def dims(self, x):
        return np.shape(x)

    def view(self, x, axis, lo, hi):
        key = [slice(None)] * np.ndim(x)
        key[axis] = slice(lo, hi)
        return x[tuple(key)]

    def map1(self, name, x, out=None):
        return self.ref.one_arg(name, np.ascontiguousarray(x), out=out)

    def map2(self, name, x, y, out=None):
        return self.ref.two_arg(self._OPS[name], np.ascontiguousarray(x), np.ascontiguousarray(np.asarray(y, dtype=x.dtype)), out=out)

    def reduce_axis(self, name, x, axis, keep_dims):
        return self.ref.reduce(name, np.ascontiguousarray(x), axis, keep_dims)

    def moments_axis(self, x, axis, keep_dims):
        x = np.ascontiguousarray(x)
        return self.ref.reduce("sum", x, axis, keep_dims), self.ref.reduce("sum", self.ref.two_arg("mul", x, x), axis, keep_dims)

    def _gather(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a).reshape(-1).copy())
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return [o.numpy().reshape(np.shape(a)) for o in outs]

    def allreduce(self, name, a):
        parts = self._gather(a)
        acc = parts[0].copy()
        for p in parts[1:]:                                   # rank order, one rounded step per rank
            acc = self.ref.two_arg(self._FOLD[name], acc, p)
        return acc

    def reduce_full(self, name, x):
        part = np.asarray(self.ref.reduce(name, np.ascontiguousarray(x)), dtype=x.dtype).reshape(1)
        return self.allreduce(name, part)[0]

    def moments_full(self, x):
        x = np.ascontiguousarray(x)
        s = np.asarray(self.ref.reduce("sum", x), dtype=x.dtype).reshape(1)
        q = np.asarray(self.ref.reduce("sum", self.ref.two_arg("mul", x, x)), dtype=x.dtype).reshape(1)
        return self.allreduce("sum", s)[0], self.allreduce("sum", q)[0]

    def divide_scalar(self, a, n):
        a[...] = a / a.dtype.type(n)
        return a

    def variance_finish(self, s, q, count, ddof):
        t = s.dtype.type
        ss = (s * s) / t(count)
        return (q - ss) / t(count - ddof)

    def scalar_type(self, x):
        return x.dtype.type


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from numericals_b200.sharded import ShardedArray, slab_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = OracleBackend(dist, world)
        full_x = np.random.default_rng(5).random(n).astype(np.float32)
        full_y = np.random.default_rng(6).random(n).astype(np.float32)
        lo, hi = slab_bounds(n, world, rank)
        x = ShardedArray.scatter(full_x, rank, world, be)
        y = ShardedArray.scatter(full_y, rank, world, be)
        assert (x.lo, x.hi, x.axis) == (lo, hi, 0)
        z = x.map2("multiply", y)                       # local, no communication
        s = z.sum()                                     # local reduce + rank-ordered combine
        flat = (rank, lo, hi, z.local.tobytes(), float(s), float(z.maximum()), float(x.mean()),
                float(x.variance()), float(x.map1("sin").sum()))
        # leading-axis sharding of a matrix (dense-numericals-src/utils/lparallel.lisp:41-49)
        rows, cols = 37, 50
        m = np.random.default_rng(7).random((rows, cols))
        rv = np.random.default_rng(8).random(cols)           # row vector: stride 0 along the cut axis -> replicated
        cv = np.random.default_rng(9).random((rows, 1))      # column vector: extends along the cut axis -> sliced
        M = ShardedArray.scatter(m, rank, world, be)
        P = M.map2("multiply", rv).map2("add", cv)
        mat = (P.local.tobytes(), P.lo, P.hi,
               P.sum(axes=0).tobytes(), P.sum(axes=1).local.tobytes(), P.sum(axes=1).axis,
               P.maximum(axes=0).tobytes(), P.mean(axes=0).tobytes(), P.mean(axes=1, keep_dims=True).local.tobytes(),
               P.variance(axes=0, ddof=1).tobytes(), P.variance(axes=1).local.tobytes(), float(P.sum()), float(P.variance()))
        # a (3, rows, 4) array: the first axis (3 >= 2 workers) is cut, not the longest one
        t3 = np.random.default_rng(10).random((3, rows, 4))
        T3 = ShardedArray.scatter(t3, rank, world, be)
        t3r = (T3.axis, T3.sum(axes=1).local.tobytes(), T3.sum(axes=0).tobytes())
        q.put((flat, mat, t3r))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [100003, 7])
def test_two_rank_protocol(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0][0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [g[0] for g in got]
    x = np.random.default_rng(5).random(n).astype(np.float32)
    y = np.random.default_rng(6).random(n).astype(np.float32)
    z = x * y
    # slabs tile [0, n) exactly, in order, ceiling(n/2) first (lparallel.lisp:66-81)
    assert res[0][1] == 0 and res[0][2] == res[1][1] == -(-n // 2) and res[1][2] == n
    assert np.frombuffer(res[0][3] + res[1][3], dtype=np.float32).tobytes() == z.tobytes()
    # every rank holds the SAME combined results (rank-ordered combine, no rank-dependent rounding)
    assert res[0][4:] == res[1][4:]
    s, mx, mean, var, ssin = res[0][4:]
    assert abs(s - float(np.sum(z.astype(np.float64)))) < 1e-5 * n
    assert mx == float(z.max())
    assert abs(mean - float(x.astype(np.float64).mean())) < 1e-6
    assert abs(var - float(x.astype(np.float64).var())) < 1e-4
    assert abs(ssin - float(np.sin(x.astype(np.float64)).sum())) < 1e-5 * n

    # ---- matrix, rows cut in two
    rows, cols = 37, 50
    m = np.random.default_rng(7).random((rows, cols))
    p = m * np.random.default_rng(8).random(cols) + np.random.default_rng(9).random((rows, 1))
    m0, m1 = got[0][1], got[1][1]
    assert (m0[1], m0[2], m1[1], m1[2]) == (0, 19, 19, 37)
    assert np.frombuffer(m0[0] + m1[0]).reshape(rows, cols).tobytes() == p.tobytes()      # maps are bit-exact
    assert m0[3:4] == m1[3:4] and m0[6:8] == m1[6:8] and m0[9] == m1[9] and m0[11:] == m1[11:]   # replicated results agree
    assert np.allclose(np.frombuffer(m0[3]), p.sum(axis=0), rtol=1e-14)
    assert np.allclose(np.frombuffer(m0[4] + m1[4]), p.sum(axis=1), rtol=1e-14) and m0[5] == 0
    assert np.array_equal(np.frombuffer(m0[6]), p.max(axis=0))
    assert np.allclose(np.frombuffer(m0[7]), p.mean(axis=0), rtol=1e-14)
    assert np.allclose(np.frombuffer(m0[8] + m1[8]), p.mean(axis=1), rtol=1e-14)
    assert np.allclose(np.frombuffer(m0[9]), p.var(axis=0, ddof=1), rtol=1e-10)
    assert np.allclose(np.frombuffer(m0[10] + m1[10]), p.var(axis=1), rtol=1e-10)
    assert abs(m0[11] - p.sum()) < 1e-10 * p.size and abs(m0[12] - p.var()) < 1e-10
    # ---- rank 3: first axis long enough is axis 0 (3 >= 2)
    t3 = np.random.default_rng(10).random((3, rows, 4))
    a0, a1 = got[0][2], got[1][2]
    assert a0[0] == a1[0] == 0
    assert np.allclose(np.frombuffer(a0[1] + a1[1]).reshape(3, 4), t3.sum(axis=1), rtol=1e-14)
    assert a0[2] == a1[2] and np.allclose(np.frombuffer(a0[2]).reshape(rows, 4), t3.sum(axis=0), rtol=1e-14)


def test_slab_bounds_cover_and_match_reference_rule():
    from numericals_b200.sharded import shard_axis, slab_bounds
    for n in (0, 1, 7, 8, 100, 2 ** 32):
        for w in (1, 2, 3, 4, 8):
            b = [slab_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            per = -(-n // w)
            assert all(hi - lo <= per for lo, hi in b)
    assert shard_axis((2, 5, 9), 4) == 1 and shard_axis((8, 5), 8) == 0 and shard_axis((3, 3), 4) is None


def test_c_abi_sharding_rule_matches_python():
    """nuk_shard_bounds / nuk_shard_axis (what a Lisp host calls) == the Python restatement"""
    import ctypes as C
    from numericals_b200 import _lib
    from numericals_b200.sharded import shard_axis, slab_bounds
    lib = _lib.lib()
    lo, hi = C.c_int64(), C.c_int64()
    for n in (0, 1, 7, 8, 100, 2 ** 32, 2 ** 32 + 12345):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                assert lib.nuk_shard_bounds(n, w, r, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == slab_bounds(n, w, r)
    for dims in ((2, 5, 9), (8, 5), (3, 3), (16384, 16384), ()):
        for w in (1, 2, 4, 8):
            want = shard_axis(dims, w)
            assert lib.nuk_shard_axis(len(dims), _lib.i64_array(dims), w) == (-1 if want is None else want)
    assert lib.nuk_shard_bounds(5, 2, 2, C.byref(lo), C.byref(hi)) != 0
