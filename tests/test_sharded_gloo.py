"""N > 1 path on CPU: world_size-2 gloo processes exercise the slab partition and the
gather-partials + fixed-order-combine protocol of numericals_b200.sharded with an oracle-backed
local backend (test infrastructure: the product backend is libnuk on the rank's GPU)."""
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
    """local compute through the CPU oracle (tests only)"""

    def __init__(self):
        from oracle import reference_nu
        self.ref = reference_nu
        self._ops = {"add": "add", "subtract": "sub", "multiply": "mul", "divide": "div"}

    def map1(self, name, x, out=None):
        return self.ref.one_arg(name, x, out=out)

    def map2(self, name, x, y, out=None):
        return self.ref.two_arg(self._ops[name], x, np.asarray(y, dtype=x.dtype), out=out)

    def reduce_local(self, name, x):
        return self.ref.reduce(name, x)

    def combine(self, name, partials, dtype):
        return self.ref.reduce(name, np.asarray(partials, dtype=dtype))

    def gather_tensor(self, value, dtype, dist, world_size):
        import torch
        t = torch.tensor([value], dtype=getattr(torch, np.dtype(dtype).name))
        outs = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(outs, t)
        return [o.item() for o in outs]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from numericals_b200.sharded import ShardedArray, slab_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_x = np.random.default_rng(5).random(n).astype(np.float32)
        full_y = np.random.default_rng(6).random(n).astype(np.float32)
        lo, hi = slab_bounds(n, world, rank)
        be = OracleBackend()
        x = ShardedArray(full_x[lo:hi].copy(), n, rank, world, be, dist)
        y = ShardedArray(full_y[lo:hi].copy(), n, rank, world, be, dist)
        z = x.map2("multiply", y)                       # local, no communication
        s = z.sum()                                     # local reduce + all-gather + fixed-order combine
        q.put((rank, lo, hi, z.local.tobytes(), float(s), float(z.maximum()), float(x.mean()),
               float(x.variance()), float(x.map1("sin").sum())))
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
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.default_rng(5).random(n).astype(np.float32)
    y = np.random.default_rng(6).random(n).astype(np.float32)
    z = x * y
    # slabs tile [0, n) exactly, in order, ceiling(n/2) first (lparallel.lisp:66-81)
    assert res[0][1] == 0 and res[0][2] == res[1][1] == -(-n // 2) and res[1][2] == n
    assert np.frombuffer(res[0][3] + res[1][3], dtype=np.float32).tobytes() == z.tobytes()
    # every rank holds the SAME combined results (fixed-order combine, no rank-dependent rounding)
    assert res[0][4:] == res[1][4:]
    s, mx, mean, var, ssin = res[0][4:]
    assert abs(s - float(np.sum(z.astype(np.float64)))) < 1e-5 * n
    assert mx == float(z.max())
    assert abs(mean - float(x.astype(np.float64).mean())) < 1e-6
    assert abs(var - float(x.astype(np.float64).var())) < 1e-4
    assert abs(ssin - float(np.sin(x.astype(np.float64)).sum())) < 1e-5 * n


def test_slab_bounds_cover_and_match_reference_rule():
    from numericals_b200.sharded import slab_bounds
    for n in (0, 1, 7, 8, 100, 2 ** 32):
        for w in (1, 2, 3, 4, 8):
            b = [slab_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            per = -(-n // w)
            assert all(hi - lo <= per for lo, hi in b)
