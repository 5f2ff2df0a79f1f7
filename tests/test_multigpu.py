"""Multi-GPU combine inside libnuk (include/nuk.h "multi-GPU"; csrc/comm.cu, comm.cuh): needs >= 2 GPUs.

* one process driving several devices (nuk_comm_init_all: what a Common Lisp image does): rank-ordered
  all-reduce, fused full reduction and moments, bit-identical on every rank and equal to the fold of the
  per-rank results in rank order;
* one process per GPU (nuk_comm_create / export / connect over cudaIpc; gloo only carries the 128-byte
  handles): numericals_b200.sharded.ShardedArray + DeviceBackend on flat and leading-axis slabs -- the
  reference's slicing rules, src/transcendental/lparallel.lisp:60-96 and
  dense-numericals-src/utils/lparallel.lisp:39-76.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    from numericals_b200 import _lib
    return _lib.lib().nuk_device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fold(parts, op):
    acc = parts[0].copy()
    for p in parts[1:]:
        acc = op(acc, p)
    return acc


def test_single_process_all_devices(nu, libnuk):
    ndev = min(_ndev(), 8)
    if ndev < 2:
        pytest.skip("needs two devices")
    from numericals_b200 import _lib
    from numericals_b200.sharded import Communicator, slab_bounds
    comms = Communicator.init_all(ndev)
    try:
        rng = np.random.default_rng(21)
        # ---- all-reduce of a vector (the combine of an axis reduction over the cut axis)
        for dt, et, red, op in ((np.float64, "double-float", 0, np.add), (np.float32, "single-float", 1, np.maximum),
                                (np.int32, "(signed-byte 32)", 0, np.add), (np.float64, "double-float", 2, np.minimum)):
            for count in (1, 16384 + 7, 300_001):          # 300001 doubles > one 1 MiB mailbox slot: two rounds
                parts = [(rng.random(count) * 100 - 50).astype(dt) for _ in range(ndev)]
                devs = []
                for i in range(ndev):
                    nu.require_device(i)
                    devs.append(nu.asarray(parts[i], type=et))
                for i in range(ndev):                        # all launches first ...
                    _lib.check(libnuk.nuk_comm_allreduce(comms[i].handle, red, devs[i].dtype_code, count,
                                                         C.c_void_p(devs[i].ptr), C.c_void_p(devs[i].ptr), None))
                want = _fold(parts, op)
                for i in range(ndev):                        # ... then the waits
                    nu.require_device(i)
                    assert devs[i].to_numpy().tobytes() == want.tobytes(), (dt, red, count, i)
                    comms[i].status()
        # ---- fused full reduction of a sharded flat array: pass 1 + (pass 2, exchange, rank-order fold)
        n = 5_000_003
        full = (rng.random(n) * 2 - 1).astype(np.float32)
        slabs, outs = [], []
        for i in range(ndev):
            nu.require_device(i)
            lo, hi = slab_bounds(n, ndev, i)
            slabs.append(nu.asarray(full[lo:hi]))
            outs.append(nu.zeros(3, type="single-float"))
        local = []
        for i in range(ndev):
            nu.require_device(i)
            local.append(np.float32(nu.sum(slabs[i])))
        for red in (0, 1, 2):
            for i in range(ndev):
                _lib.check(libnuk.nuk_comm_reduce(comms[i].handle, red, 0, 1, _lib.i64_array([slabs[i].total_size]),
                                                  _lib.i64_array([1]), C.c_void_p(slabs[i].ptr),
                                                  C.c_void_p(outs[i].ptr + 4 * red), None))
        got = []
        for i in range(ndev):
            nu.require_device(i)
            got.append(outs[i].to_numpy())
            comms[i].status()
        assert all(g.tobytes() == got[0].tobytes() for g in got), "ranks disagree"
        assert got[0][0] == _fold([np.array(v) for v in local], np.add)      # rank-ordered fold of the local sums
        assert got[0][1] == full.max() and got[0][2] == full.min()
        # ---- moments pair and the x[0]-seeded NaN rule across ranks
        full[0] = np.nan
        nu.require_device(0)
        slabs[0] = nu.asarray(full[slice(*slab_bounds(n, ndev, 0))])
        full[n - 1] = np.nan                                # a NaN that is NOT first is ignored
        nu.require_device(ndev - 1)
        slabs[ndev - 1] = nu.asarray(full[slice(*slab_bounds(n, ndev, ndev - 1))])
        for i in range(ndev):
            _lib.check(libnuk.nuk_comm_reduce(comms[i].handle, 1, 0, 1, _lib.i64_array([slabs[i].total_size]),
                                              _lib.i64_array([1]), C.c_void_p(slabs[i].ptr), C.c_void_p(outs[i].ptr), None))
        for i in range(ndev):
            nu.require_device(i)
            assert np.isnan(outs[i].to_numpy()[0]), i
    finally:
        for c in comms:
            c.close()
        nu.require_device(0)


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import numericals_b200 as nu
    from numericals_b200.sharded import Communicator, DeviceBackend, ShardedArray
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = None
    try:
        nu.require_device(rank)
        comm = Communicator.from_process_group(dist, rank, world)
        be = DeviceBackend(comm)
        n = 3_000_001
        fx = np.random.default_rng(5).random(n).astype(np.float32)
        fy = np.random.default_rng(6).random(n).astype(np.float32)
        x = ShardedArray.scatter(nu.asarray(fx), rank, world, be)
        y = ShardedArray.scatter(nu.asarray(fy), rank, world, be)
        z = x.map2("multiply", y)
        flat = (z.local.to_numpy().tobytes(), float(z.sum()), float(z.maximum()), float(x.mean()), float(x.variance()),
                float(x.map1("sin").sum()))
        rows, cols = 3001, 4100
        m = np.random.default_rng(7).random((rows, cols))
        rv = np.random.default_rng(8).random(cols)
        cv = np.random.default_rng(9).random((rows, 1))
        M = ShardedArray.scatter(nu.asarray(m), rank, world, be)
        P = M.map2("multiply", nu.asarray(rv)).map2("add", nu.asarray(cv))
        mat = (P.local.to_numpy().tobytes(), P.sum(axes=0).to_numpy().tobytes(), P.sum(axes=1).local.to_numpy().tobytes(),
               P.maximum(axes=0).to_numpy().tobytes(), P.mean(axes=0).to_numpy().tobytes(),
               P.mean(axes=1, keep_dims=True).local.to_numpy().tobytes(), P.variance(axes=0, ddof=1).to_numpy().tobytes(),
               P.variance(axes=1).local.to_numpy().tobytes(), float(P.sum()), float(P.variance()))
        comm.status()
        q.put((rank, flat, mat))
    finally:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


def test_process_per_gpu_sharded_arrays():
    world = min(_ndev(), 8)
    if world < 2:
        pytest.skip("needs two devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n = 3_000_001
    fx = np.random.default_rng(5).random(n).astype(np.float32)
    fy = np.random.default_rng(6).random(n).astype(np.float32)
    z = fx * fy
    assert b"".join(g[1][0] for g in got) == z.tobytes()                    # maps: bit-exact, slabs tile the array
    assert all(g[1][1:] == got[0][1][1:] for g in got), "ranks disagree on the combined scalars"
    s, mx, mean, var, ssin = got[0][1][1:]
    assert abs(s - float(z.astype(np.float64).sum())) < 2 * 2.0 ** -24 * 22 * float(np.abs(z).sum())
    assert mx == float(z.max())
    assert abs(mean - float(fx.astype(np.float64).mean())) < 1e-6 and abs(var - float(fx.astype(np.float64).var())) < 1e-5
    assert abs(ssin - float(np.sin(fx.astype(np.float64)).sum())) < 1e-6 * n
    rows, cols = 3001, 4100
    p = np.random.default_rng(7).random((rows, cols)) * np.random.default_rng(8).random(cols) + \
        np.random.default_rng(9).random((rows, 1))
    mats = [g[2] for g in got]
    assert np.frombuffer(b"".join(m[0] for m in mats)).reshape(rows, cols).tobytes() == p.tobytes()
    for k in (1, 3, 4, 6, 8, 9):
        assert all(m[k] == mats[0][k] for m in mats), "replicated result %d differs between ranks" % k
    assert np.allclose(np.frombuffer(mats[0][1]), p.sum(axis=0), rtol=1e-13)
    assert np.allclose(np.frombuffer(b"".join(m[2] for m in mats)), p.sum(axis=1), rtol=1e-13)
    assert np.array_equal(np.frombuffer(mats[0][3]), p.max(axis=0))
    assert np.allclose(np.frombuffer(mats[0][4]), p.mean(axis=0), rtol=1e-13)
    assert np.allclose(np.frombuffer(b"".join(m[5] for m in mats)), p.mean(axis=1), rtol=1e-13)
    assert np.allclose(np.frombuffer(mats[0][6]), p.var(axis=0, ddof=1), rtol=1e-9)
    assert np.allclose(np.frombuffer(b"".join(m[7] for m in mats)), p.var(axis=1), rtol=1e-9)
    assert abs(mats[0][8] - p.sum()) < 1e-10 * p.size and abs(mats[0][9] - p.var()) < 1e-9


def _lonely_rank(q):
    """child process: rank 0 of a 2-rank communicator issues a collective its peer never joins"""
    sys.path.insert(0, ROOT)
    os.environ["NUK_COMM_TIMEOUT_MS"] = "500"
    import numericals_b200 as nu
    from numericals_b200 import _lib
    from numericals_b200.sharded import Communicator
    import time
    lib = _lib.lib()
    comms = Communicator.init_all(2)
    nu.require_device(0)
    a = nu.asarray(np.arange(1024, dtype=np.float64))
    t0 = time.time()
    _lib.check(lib.nuk_comm_allreduce(comms[0].handle, 0, a.dtype_code, 1024, C.c_void_p(a.ptr), C.c_void_p(a.ptr), None))
    lib.nuk_device_sync()                       # returns: the kernel gives up after the timeout, it does not hang
    waited = time.time() - t0
    st = lib.nuk_comm_status(comms[0].handle)
    st2 = lib.nuk_comm_allreduce(comms[0].handle, 0, a.dtype_code, 1024, C.c_void_p(a.ptr), C.c_void_p(a.ptr), None)
    q.put((st, st2, waited, lib.nuk_last_error().decode()))


def test_absent_peer_times_out_instead_of_hanging():
    """A peer that never arrives must not leave a spinning kernel behind: after NUK_COMM_TIMEOUT_MS the kernels
    give up, nuk_comm_status and every later collective return NUK_ERR_COMM (5)."""
    if _ndev() < 2:
        pytest.skip("needs two devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_lonely_rank, args=(q,))
    p.start()
    st, st2, waited, msg = q.get(timeout=120)
    p.join(timeout=60)
    assert st == 5 and st2 == 5, (st, st2, msg)
    assert 0.4 <= waited < 30, waited
    assert "did not arrive" in msg
