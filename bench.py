#!/usr/bin/env python
"""bench.py -- the hot path of numericals on B200, measured.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): the single-float one-arg transcendentals nu:exp nu:sin
nu:log nu:tanh on a contiguous 2^28-element array, device-resident, :out preallocated.  One STEP =
one pass of each of the four ufuncs over the array (4 launches of k_map_flat).  With N GPUs every
rank owns one 2^28-element slab of an (N * 2^28)-element array (weak scaling, no data-path
collective: elementwise maps shard by contiguous slabs).

  value      Gelem/s over the whole job, inputs resident in HBM (arrays are 1 GiB: larger than L2)
  e2e        same metric through the reference-facing BMAS_s<fn> symbols with HOST (page-locked)
             buffers: chunked H2D -> kernel -> D2H inside the timed region; e2e.pageable: the same
             calls on plain (pageable) numpy buffers, which is what a Lisp heap vector is
  roofline   dominant kernel (the slowest of the four maps): algorithmic 8 B/element x 2^28 / its
             CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (BMAS contract restated over SLEEF u10) timed on this box's host cores
  extra      C1 / C3 / C4 / C5 of BASELINE.json configs, each with its own algorithmic GB/s; C5 and the sharded
             C4 go through libnuk's own communicator (include/nuk.h "multi-GPU"): torch.distributed only
             carries the 128-byte handles, barriers and the max-over-ranks of the timings

Nothing here reads /root/reference.  The oracle is only executed by cpu_baseline and --impl reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ELEMS = 1 << 28
UFUNCS = ("exp", "sin", "log", "tanh")
METRIC = "f32 one-arg transcendental ufunc throughput (nu:exp nu:sin nu:log nu:tanh, 2^28 contiguous, device-resident)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json: torch copy_ 1 Gi bf16)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_pass(handle, fn_names, x, out, threads):
    """one pass of each ufunc over x with `threads` equal contiguous slices (the reference's
    lparallel rule, src/transcendental/lparallel.lisp:60-96); returns seconds"""
    n = x.size
    per = -(-n // threads)
    fns = []
    for name in fn_names:
        f = getattr(handle, "BMAS_s" + name)
        f.restype = None
        fns.append(f)

    def work(t):
        lo, hi = t * per, min(n, (t + 1) * per)
        if lo >= hi:
            return
        for f in fns:
            f(C.c_long(hi - lo), C.c_void_p(x.ctypes.data + 4 * lo), C.c_long(1),
              C.c_void_p(out.ctypes.data + 4 * lo), C.c_long(1))

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    return time.perf_counter() - t0


def cpu_baseline(sample_elems, reps, threads):
    """The oracle (kind 'port': BMAS-equivalent loops over the same SLEEF 3.8 u10 the reference's
    BMAS wraps) on the host cores, on a bounded sample of the C2 workload."""
    import oracle as O
    h = O.load(widest=True)
    rng = np.random.default_rng(0)
    x = rng.random(sample_elems).astype(np.float32)
    out = np.empty_like(x)
    cpu_pass(h, UFUNCS, x, out, threads)                       # warm-up, first touch
    best_all = min(cpu_pass(h, UFUNCS, x, out, threads) for _ in range(reps))
    best_one = min(cpu_pass(h, UFUNCS, x[: sample_elems // 8], out[: sample_elems // 8], 1) for _ in range(max(1, reps // 2)))
    return {
        "value": len(UFUNCS) * sample_elems / best_all / 1e9, "unit": "Gelem/s", "cores": threads, "kind": "port",
        "simd": h._oracle_kind,
        "value_1thread": len(UFUNCS) * (sample_elems // 8) / best_one / 1e9,
        "sample": "%d f32 elements x 4 ufuncs (exp sin log tanh), best of %d passes, %d threads over equal "
                  "contiguous slices; 1-thread figure on %d elements (the reference's contiguous one-arg "
                  "polymorph is a single BMAS call, src/basic-math/one-arg-fn-float.lisp:56-75)"
                  % (sample_elems, reps, threads, sample_elems // 8),
    }


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = cpu_threads()
    sample = N_ELEMS                      # the whole C2 array: same config as the GPU arm
    import oracle as O
    h = O.load(widest=True)
    rng = np.random.default_rng(0)
    x = rng.random(sample).astype(np.float32)
    out = np.empty_like(x)
    for _ in range(max(1, args.warmup)):
        cpu_pass(h, UFUNCS, x, out, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_pass(h, UFUNCS, x, out, threads)
    ms = 1e3 * t / args.steps
    val = len(UFUNCS) * sample / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gelem/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic U[0,1)",
        "config": {"workload": "C2: nu:exp nu:sin nu:log nu:tanh, single-float, contiguous, :out preallocated; "
                               "each step = the four maps over the full 2^28-element array on the host cores",
                   "elements_per_array": sample, "ufuncs": list(UFUNCS)},
        "cpu_baseline": {"value": val, "unit": "Gelem/s", "cores": threads, "kind": "port", "simd": h._oracle_kind,
                         "sample": "2^28 f32 elements x 4 ufuncs per step (the full array), %d threads over equal slices "
                                   "(reference rule: src/transcendental/lparallel.lisp:60-96); the reference itself "
                                   "is Common Lisp over the un-vendored BMAS library and cannot run here, so this is "
                                   "the oracle port over the same SLEEF 3.8 u10" % threads},
        "e2e": {"value": val, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def bind_near_gpu(torch, local_rank, world):
    """Host threads of this rank (and libnuk's mover threads, which inherit the mask) onto the cores that are
    local to the rank's GPU: the e2e legs stream 8 GiB per step through host memory, and on a two-socket box a
    page-locked buffer that lands on the other socket halves the link.  No-op when the platform exposes a
    single NUMA node (numa_node = -1) or when there is only one rank."""
    info = {"numa_node": None, "cpus": None}
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(path + "/numa_node").read().strip())
        cpulist = open(path + "/local_cpulist").read().strip()
        info["numa_node"] = node
        if world > 1 and node >= 0 and cpulist:
            cpus = set()
            for part in cpulist.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= set(os.sched_getaffinity(0))
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"] = len(cpus)
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)[:80]
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C1/C3/C4/C5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=2, help="0 skips the end-to-end legs")
    ap.add_argument("--timed-only", action="store_true",
                    help="only the timed region (no e2e / extra / cpu legs): the command profiled for the ncu launch list")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner)
    # are sent to stderr for the duration of the run; emit() writes to the saved descriptor.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(saved_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)
    args.warmup = max(args.warmup, 3)
    if args.timed_only:
        args.no_extra, args.no_cpu, args.e2e_steps = True, True, 0

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)

    import numericals_b200 as nu
    from numericals_b200 import _lib
    from numericals_b200.array import from_pointer

    affinity = bind_near_gpu(torch, local_rank, world)
    nu.require_device(local_rank)
    lib = _lib.lib()
    stream = torch.cuda.current_stream()
    lib.nuk_set_stream(C.c_void_p(stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- data: this rank's slab, synthetic U[0,1) (the reference's test domain), resident in HBM
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    xt = torch.rand(N_ELEMS, dtype=torch.float32, device="cuda", generator=gen).clamp_(min=1e-6)
    ot = torch.empty_like(xt)
    x = from_pointer(xt.data_ptr(), "single-float", (N_ELEMS,), owner=xt)
    out = from_pointer(ot.data_ptr(), "single-float", (N_ELEMS,), owner=ot)
    fns = [getattr(nu, f) for f in UFUNCS]

    def step():
        for f in fns:
            f(x, out=out, broadcast=False)

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps; per-launch CUDA events on the launching stream
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in fns]
          for _ in range(args.steps)]
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = nu.launch_count()
    t_start.record(stream)
    for k in range(args.steps):
        for j, f in enumerate(fns):
            ev[k][j][0].record(stream)
            f(x, out=out, broadcast=False)
            ev[k][j][1].record(stream)
    t_stop.record(stream)
    barrier()
    launches = nu.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(t_start.elapsed_time(t_stop))
    ms_step = ms_total / args.steps
    per_fn_ms = [float(np.mean([ev[k][j][0].elapsed_time(ev[k][j][1]) for k in range(args.steps)])) for j in range(len(fns))]
    value = world * len(UFUNCS) * N_ELEMS / (ms_step / 1e3) / 1e9

    peak, peak_src = peaks()
    per_ufunc = {}
    for name, ms in zip(UFUNCS, per_fn_ms):
        gbs = 8.0 * N_ELEMS / (ms / 1e3) / 1e9
        per_ufunc[name] = {"ms": ms, "gelem_s": N_ELEMS / (ms / 1e3) / 1e9, "hbm_gbs": gbs, "frac_of_measured_peak": gbs / peak,
                           "frac_of_8tbs": gbs / 8000.0}
    dom = max(range(len(UFUNCS)), key=lambda j: per_fn_ms[j])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_map_flat<%s>" % UFUNCS[dom])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_map_flat<%sF> (nu:%s, f32, 2^28)" % (UFUNCS[dom].capitalize(), UFUNCS[dom]),
                "achieved": per_ufunc[UFUNCS[dom]]["hbm_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_ufunc[UFUNCS[dom]]["hbm_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 8 * N_ELEMS,
                "all_four_maps_frac": float(np.mean([v["frac_of_measured_peak"] for v in per_ufunc.values()]))}

    # ---- parity spot check inside the bench (cheap): device result == host compile is covered by tests;
    #      here just make sure the kernels did the work (finite, in range)
    chk = float(ot[:: 1 << 16].abs().max().item())
    assert 0.0 <= chk <= 1.0, "tanh output out of range: kernels did not run?"

    # ---- e2e: the reference-facing C ABI with HOST buffers, copies inside the timed region.
    #      Two legs: page-locked buffers (nuk_host_alloc: direct DMA) and PAGEABLE numpy buffers (what a
    #      Lisp heap vector is: staged through libnuk's page-locked ring by its host mover threads).
    e2e = None
    if args.e2e_steps > 0:
        bm = [_lib.bmas("s" + f) for f in UFUNCS]
        src = np.random.default_rng(99 + rank).random(N_ELEMS, dtype=np.float32) + np.float32(1e-6)

        def e2e_leg(px, po, steps):
            def one():
                for f in bm:
                    f(C.c_long(N_ELEMS), px, C.c_long(1), po, C.c_long(1))
            one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                one()                                    # synchronous: returns when OUT is on the host
            barrier()
            ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / steps
            assert _lib.last_status() == 0, lib.nuk_last_error()
            return ms

        hx, ho = C.c_void_p(), C.c_void_p()
        _lib.check(lib.nuk_host_alloc(C.byref(hx), 4 * N_ELEMS))
        _lib.check(lib.nuk_host_alloc(C.byref(ho), 4 * N_ELEMS))
        hxa = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(N_ELEMS,))
        hxa[:] = src
        e2e_ms = e2e_leg(hx, ho, args.e2e_steps)
        hoa = np.ctypeslib.as_array(C.cast(ho, C.POINTER(C.c_float)), shape=(N_ELEMS,))
        assert np.all(np.abs(hoa[:: 1 << 14]) <= 1.0)
        pinned_tail = hoa[-4096:].copy()
        lib.nuk_host_free(hx)
        lib.nuk_host_free(ho)
        pout = np.empty(N_ELEMS, dtype=np.float32)
        pout[:] = 0                                      # touch: the pages exist before the clock starts
        pg_ms = e2e_leg(C.c_void_p(src.ctypes.data), C.c_void_p(pout.ctypes.data), args.e2e_steps)
        assert np.array_equal(pout[-4096:], pinned_tail), "pageable and page-locked legs disagree"
        e2e = {"value": world * len(UFUNCS) * N_ELEMS / (e2e_ms / 1e3) / 1e9, "unit": "Gelem/s", "ms_per_step": e2e_ms,
               "steps": args.e2e_steps, "h2d_bytes_per_step": len(UFUNCS) * 4 * N_ELEMS,
               "d2h_bytes_per_step": len(UFUNCS) * 4 * N_ELEMS,
               "pcie_gbs_each_way_per_gpu": len(UFUNCS) * 4 * N_ELEMS / (e2e_ms / 1e3) / 1e9,
               "path": "BMAS_sexp/ssin/slog/stanh(n, host x, 1, host out, 1): page-locked host buffers, %d MiB chunks, "
                       "H2D | kernel | D2H overlapped on 3 streams; wall clock (the call is synchronous)"
                       % lib.nuk_get_option(b"stage_mb"),
               "pageable": {"value": world * len(UFUNCS) * N_ELEMS / (pg_ms / 1e3) / 1e9, "unit": "Gelem/s",
                            "ms_per_step": pg_ms, "frac_of_page_locked": e2e_ms / pg_ms,
                            "path": "same calls on plain numpy buffers (pageable, like a Lisp heap vector): %d host "
                                    "mover threads copy each chunk through a page-locked ring"
                                    % lib.nuk_host_mover_threads()}}
        del src, pout

    extra = {}
    if not args.no_extra:
        extra = side_measurements(nu, lib, torch, dist, stream, rank, world, barrier, max_over_ranks, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(1 << 27, 5, cpu_threads())

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic U[1e-6,1) per rank (torch.rand on device)",
            "config": {"workload": "C2 (BASELINE.json configs[1]): nu:exp nu:sin nu:log nu:tanh, single-float, "
                                   "2^28-element contiguous array per GPU, device-resident, :out preallocated, "
                                   ":broadcast nil; one step = the four maps",
                       "elements_per_array_per_gpu": N_ELEMS, "ufuncs": list(UFUNCS),
                       "sharding": "contiguous slab per rank, no data-path collective",
                       "l2": "inputs (1 GiB) and outputs (1 GiB) are each ~8x the 126 MB L2; no flush needed"},
            "hbm_gbs": value * 8.0, "per_ufunc": per_ufunc, "roofline": roofline, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": cpu, "host_affinity": affinity,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def side_measurements(nu, lib, torch, dist, stream, rank, world, barrier, max_over_ranks, peak):
    """BASELINE.json configs C1, C3, C4 (rank-local, reported for rank 0's GPU) and C5 (sharded)."""
    from numericals_b200.array import from_pointer
    res = {}

    def timeit(fn, reps, flush=False):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not flush:
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        tot = 0.0
        for _ in range(reps):
            lib.nuk_flush_l2(C.c_void_p(stream.cuda_stream))
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    def wrap(t, et):
        return from_pointer(t.data_ptr(), et, tuple(t.shape), owner=t)

    if rank == 0:
        # C1: nu:add, double-float, 10^6 contiguous (benchmarks/two-arg-fn.lisp:50-81). L2-resident, launch-bound.
        a = torch.ones(1000, 1000, dtype=torch.float64, device="cuda")
        b = torch.ones_like(a)
        o = torch.empty_like(a)
        A, B, O_ = wrap(a, "double-float"), wrap(b, "double-float"), wrap(o, "double-float")
        from numericals_b200 import _lib as _L
        dadd = _L.bmas("dadd")
        pa, pb, po = C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(o.data_ptr())
        n1, one = C.c_long(10 ** 6), C.c_long(1)
        ms_abi = timeit(lambda: dadd(n1, pa, one, pb, one, po, one), 500)       # the C-ABI call the Lisp side makes
        ms_api = timeit(lambda: nu.add(A, B, out=O_, broadcast=False), 200)    # through the Python mirror
        ms_cold = timeit(lambda: dadd(n1, pa, one, pb, one, po, one), 20, flush=True)
        assert bool((o == 2).all())
        # the same call with the device-storage contract (no driver pointer query per operand) and as a
        # CUDA-graph replay of 20 captured calls: the launch-latency floor of a 10^6-element map
        lib.nuk_set_pointer_mode(1)
        ms_dev = timeit(lambda: dadd(n1, pa, one, pb, one, po, one), 500)
        gs = C.c_void_p()
        _L.check(lib.nuk_stream_create(C.byref(gs)))
        lib.nuk_set_stream(gs)
        dadd(n1, pa, one, pb, one, po, one)
        _L.check(lib.nuk_graph_begin(gs))
        for _ in range(20):
            dadd(n1, pa, one, pb, one, po, one)
        graph = C.c_void_p()
        _L.check(lib.nuk_graph_end(gs, C.byref(graph)))
        ev0, ev1 = C.c_void_p(), C.c_void_p()
        lib.nuk_event_create(C.byref(ev0)); lib.nuk_event_create(C.byref(ev1))
        lib.nuk_graph_launch(graph, gs)
        lib.nuk_stream_sync(gs)
        lib.nuk_event_record(ev0, gs)
        for _ in range(50):
            lib.nuk_graph_launch(graph, gs)
        lib.nuk_event_record(ev1, gs)
        lib.nuk_stream_sync(gs)
        gms = C.c_float()
        lib.nuk_event_elapsed_ms(ev0, ev1, C.byref(gms))
        us_graph = gms.value * 1e3 / (50 * 20)
        lib.nuk_graph_destroy(graph)
        lib.nuk_event_destroy(ev0); lib.nuk_event_destroy(ev1)
        lib.nuk_set_stream(C.c_void_p(stream.cuda_stream))
        lib.nuk_stream_destroy(gs)
        lib.nuk_set_pointer_mode(0)
        assert bool((o == 2).all())
        res["C1_add_f64_1e6"] = {"us_per_call_c_abi": ms_abi * 1e3, "us_per_call_python_api": ms_api * 1e3,
                                 "us_per_call_c_abi_device_pointer_mode": ms_dev * 1e3, "us_per_call_graph": us_graph,
                                 "graph_algorithmic_gbs": 24e6 / (us_graph / 1e6) / 1e9,
                                 "us_per_call_cold_l2": ms_cold * 1e3, "gelem_s": 1e6 / (ms_abi / 1e3) / 1e9,
                                 "algorithmic_gbs": 24e6 / (ms_abi / 1e3) / 1e9,
                                 "cold_l2_algorithmic_gbs": 24e6 / (ms_cold / 1e3) / 1e9,
                                 "note": "24 MB working set is L2-resident and back-to-back calls are launch-"
                                         "latency bound: GB/s above the HBM peak is L2 bandwidth, not DRAM"}
        del a, b, o
        # C3a: (8192 x 1) + (1 x 8192) -> (8192 x 8192), f32, OUT preallocated: ~4 B per output element
        col = torch.rand(8192, 1, device="cuda")
        row = torch.rand(1, 8192, device="cuda")
        o = torch.empty(8192, 8192, device="cuda")
        Cc, R, O_ = wrap(col, "single-float"), wrap(row, "single-float"), wrap(o, "single-float")
        ms = timeit(lambda: nu.add(Cc, R, out=O_), 20)
        by = 4 * 8192 * 8192 + 2 * 4 * 8192
        assert torch.equal(o, col + row)
        res["C3a_col_plus_row_8192"] = {"ms": ms, "gelem_s": 8192 * 8192 / (ms / 1e3) / 1e9,
                                        "algorithmic_gbs": by / (ms / 1e3) / 1e9, "frac_of_measured_peak": by / (ms / 1e3) / 1e9 / peak}
        del o
        # C3b: (4096 x 4096) * (4096), f32: ~8 B/element (128 MiB: about the size of L2 -> flush between reps)
        m = torch.rand(4096, 4096, device="cuda")
        v = torch.rand(4096, device="cuda")
        o = torch.empty_like(m)
        M, V, O_ = wrap(m, "single-float"), wrap(v, "single-float"), wrap(o, "single-float")
        ms = timeit(lambda: nu.multiply(M, V, out=O_), 20, flush=True)
        assert torch.equal(o, m * v)
        by = 2 * 4 * 4096 * 4096 + 4 * 4096
        res["C3b_matrix_times_row_4096"] = {"ms": ms, "algorithmic_gbs": by / (ms / 1e3) / 1e9,
                                            "frac_of_measured_peak": by / (ms / 1e3) / 1e9 / peak, "l2": "flushed between reps"}
        m = torch.rand(16384, 16384, device="cuda")
        v = torch.rand(16384, device="cuda")
        o = torch.empty_like(m)
        M, V, O_ = wrap(m, "single-float"), wrap(v, "single-float"), wrap(o, "single-float")
        ms = timeit(lambda: nu.multiply(M, V, out=O_), 10)
        by = 2 * 4 * 16384 * 16384
        res["C3b_matrix_times_row_16384"] = {"ms": ms, "algorithmic_gbs": by / (ms / 1e3) / 1e9,
                                             "frac_of_measured_peak": by / (ms / 1e3) / 1e9 / peak}
        del m, o
        # C4: sum / maximum / mean over axis 0, axis 1 and everything, 16384 x 16384 double-float (2 GiB)
        d = torch.rand(16384, 16384, dtype=torch.float64, device="cuda")
        D = wrap(d, "double-float")
        o0 = nu.empty((16384,), "double-float")
        by = 8 * 16384 * 16384
        c4 = {}
        for name in ("sum", "maximum", "mean"):
            f = getattr(nu, name)
            for ax in (0, 1):
                ms = timeit(lambda: f(D, axes=ax, out=o0), 10)
                c4["%s_axis%d" % (name, ax)] = {"ms": ms, "algorithmic_gbs": by / (ms / 1e3) / 1e9,
                                                "frac_of_measured_peak": by / (ms / 1e3) / 1e9 / peak}
            ms = timeit(lambda: f(D), 10)
            c4["%s_full" % name] = {"ms": ms, "algorithmic_gbs": by / (ms / 1e3) / 1e9,
                                    "frac_of_measured_peak": by / (ms / 1e3) / 1e9 / peak,
                                    "note": "includes the 8-byte D2H of the scalar result and a stream sync"}
        nu.config.reduce_reference_order = True
        ms = timeit(lambda: nu.sum(D, axes=0, out=o0), 5)
        nu.config.reduce_reference_order = False
        c4["sum_axis0_reference_order"] = {"ms": ms, "algorithmic_gbs": by / (ms / 1e3) / 1e9,
                                           "note": "strictly sequential rows per column: bit-identical to the reference's row-accumulate order"}
        ref = d.sum(dim=0)
        got = torch.from_numpy(nu.sum(D, axes=0).to_numpy()).cuda()
        assert torch.allclose(ref, got, rtol=1e-12)
        res["C4_reductions_f64_16384sq"] = c4
        del d
    barrier()
    # ---- multi-GPU legs go through libnuk's own communicator (include/nuk.h "multi-GPU"): torch.distributed only
    #      carries the 128-byte mailbox handles; the exchange itself happens inside libnuk's reduction kernels
    from numericals_b200.sharded import Communicator, DeviceBackend, ShardedArray, slab_bounds
    comm = Communicator.from_process_group(dist, rank, world)
    sptr = C.c_void_p(stream.cuda_stream)
    # C5: nu:multiply + nu:sum on 2^32 single-float elements sharded in contiguous slabs (strong scaling):
    # 3 launches per step and rank: the map, pass 1 of the sum, and ONE kernel for pass 2 + exchange + rank-order fold
    total = 1 << 32
    lo, hi = slab_bounds(total, world, rank)
    n = hi - lo
    free_b = torch.cuda.mem_get_info()[0]
    if 12 * n + (6 << 30) < free_b:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(777 + rank)
        xs = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen).mul_(2).sub_(1)     # U[-1,1)
        ys = torch.rand(n, dtype=torch.float32, device="cuda", generator=gen).add_(0.5)           # U[0.5,1.5)
        zs = torch.empty_like(xs)
        X, Y, Z = (from_pointer(t.data_ptr(), "single-float", (n,), owner=t) for t in (xs, ys, zs))
        final = torch.zeros(2, dtype=torch.float32, device="cuda")
        dims, st = _i64(n), _i64(1)

        def c5():
            nu.multiply(X, Y, out=Z, broadcast=False)
            _check(lib.nuk_comm_reduce(comm.handle, 0, 0, 1, dims, st, C.c_void_p(Z.ptr), C.c_void_p(final.data_ptr()), sptr))

        def c5_fused():   # the same quantity without materialising z: nu:vdot's kernel shape would be 8 B/element
            _check(lib.nuk_comm_reduce(comm.handle, 0, 0, 1, dims, st, C.c_void_p(Z.ptr), C.c_void_p(final.data_ptr() + 4), sptr))

        c5()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record(stream)
        for _ in range(reps):
            c5()
        b.record(stream)
        barrier()
        ms = max_over_ranks(a.elapsed_time(b) / reps)
        a.record(stream)
        for _ in range(reps):
            c5_fused()
        b.record(stream)
        barrier()
        ms_sum = max_over_ranks(a.elapsed_time(b) / reps)
        comm.status()
        tot = float(final[0].item())
        # checks on RANDOM data: (1) z == x * y on windows at both ends of the slab and around the 2^31-element
        # mark of the slab (64-bit indexing), against torch's IEEE multiply; (2) the sum against the exact sum
        # (double accumulation of the same products, all-reduced) within the pairwise bound; (3) every rank
        # holds the same bits
        for w0 in sorted(set(max(0, min(n - (1 << 22), w)) for w in (0, (1 << 31) - (1 << 21), n - (1 << 22), n // 3 + 1))):
            assert torch.equal(zs[w0:w0 + (1 << 22)], xs[w0:w0 + (1 << 22)] * ys[w0:w0 + (1 << 22)]), "C5 window %d" % w0
        ex = torch.zeros(2, dtype=torch.float64, device="cuda")
        for c0 in range(0, n, 1 << 28):
            d = zs[c0:c0 + (1 << 28)].double()
            ex[0] += d.sum()
            ex[1] += d.abs().sum()
            del d
        if world > 1:
            dist.all_reduce(ex)
            allf = [torch.zeros_like(final) for _ in range(world)]
            dist.all_gather(allf, final)
            assert all(torch.equal(f, allf[0]) for f in allf), "ranks hold different sums"
        exact, sabs = float(ex[0].item()), float(ex[1].item())
        bound = 2 * 2.0 ** -24 * 32 * sabs
        assert abs(tot - exact) <= bound, (tot, exact, bound)
        assert float(final[1].item()) == tot
        if rank == 0:
            res["C5_multiply_sum_2p32_sharded"] = {
                "ms": ms, "n_gpus": world, "elements_total": total,
                "algorithmic_gbs_aggregate": 16.0 * total / (ms / 1e3) / 1e9,
                "per_gpu_frac_of_measured_peak": 16.0 * total / world / (ms / 1e3) / 1e9 / peak,
                "scaling": "strong", "launches_per_step_per_rank": 3,
                "sum_only_ms": ms_sum, "sum_only_gbs_aggregate": 4.0 * total / (ms_sum / 1e3) / 1e9,
                "data": "random U[-1,1) * U[0.5,1.5) per rank", "sum": tot, "sum_exact_in_double": exact,
                "abs_err": abs(tot - exact), "pairwise_bound": bound, "ranks_bit_identical": True,
                "combine": "nuk_comm_reduce: pass 2 + exchange over NVLink peer memory + rank-order fold in one kernel"}
        del xs, ys, zs
    elif rank == 0:
        res["C5_multiply_sum_2p32_sharded"] = {"skipped": "needs %d GiB per GPU" % (12 * n >> 30)}
    barrier()
    # C4 sharded: the 16384 x 16384 double-float matrix cut along its first axis (the reference's rule for
    # dense arrays, dense-numericals-src/utils/lparallel.lisp:41-49): axis 1 is local, axis 0 = local column
    # reduction + rank-ordered all-reduce of a 16384-vector, full = nuk_comm_reduce (strong scaling)
    if world > 1:
        be = DeviceBackend(comm)
        R = 16384
        rlo, rhi = slab_bounds(R, world, rank)
        gen = torch.Generator(device="cuda")
        gen.manual_seed(4242 + rank)
        dm = torch.rand(rhi - rlo, R, dtype=torch.float64, device="cuda", generator=gen)
        M = ShardedArray(from_pointer(dm.data_ptr(), "double-float", (rhi - rlo, R), owner=dm), (R, R), 0, rank, world, be)
        full_out = torch.zeros(1, dtype=torch.float64, device="cuda")
        d2, s2 = _i64(rhi - rlo, R), _i64(R, 1)

        def timed(fn, reps=10):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            barrier()
            return max_over_ranks(a.elapsed_time(b) / reps)

        by = 8.0 * R * R
        c4s = {}
        o_cols = nu.empty((R,), "double-float")              # :out preallocated, as in the single-GPU C4 leg
        o_rows = nu.empty((rhi - rlo,), "double-float")
        for name, fn in (("sum_axis0", lambda: M.sum(axes=0, out=o_cols)), ("sum_axis1", lambda: M.sum(axes=1, out=o_rows)),
                         ("maximum_axis0", lambda: M.maximum(axes=0, out=o_cols)),
                         ("mean_axis0", lambda: M.mean(axes=0, out=o_cols)),
                         ("sum_full", lambda: _check(lib.nuk_comm_reduce(comm.handle, 0, 1, 2, d2, s2, C.c_void_p(dm.data_ptr()),
                                                                         C.c_void_p(full_out.data_ptr()), sptr)))):
            ms = timed(fn)
            c4s[name] = {"ms": ms, "algorithmic_gbs_aggregate": by / (ms / 1e3) / 1e9,
                         "per_gpu_frac_of_measured_peak": by / world / (ms / 1e3) / 1e9 / peak}
        comm.status()
        # parity of the sharded results: axis 0 against the rank-ordered fold of per-rank torch sums, full sum
        # against the exact double sum; identical on every rank
        got0 = torch.from_numpy(M.sum(axes=0).to_numpy()).cuda()
        ref0 = dm.sum(dim=0)
        dist.all_reduce(ref0)
        assert torch.allclose(got0, ref0, rtol=1e-12)
        allg = [torch.zeros_like(got0) for _ in range(world)]
        dist.all_gather(allg, got0)
        assert all(torch.equal(g, allg[0]) for g in allg), "ranks hold different axis-0 sums"
        ref = dm.sum().reshape(1)
        dist.all_reduce(ref)
        assert abs(float(full_out.item()) - float(ref.item())) <= 1e-12 * float(ref.item())
        if rank == 0:
            c4s["note"] = ("rows cut in %d slabs; axis 0 = local TMA column reduction + nuk_comm_allreduce of 16384 doubles "
                           "(rank-ordered fold over peer memory), axis 1 local, full = nuk_comm_reduce" % world)
            res["C4_sharded_f64_16384sq"] = c4s
        del dm
    comm.close()
    return res


def _i64(*v):
    return (C.c_int64 * len(v))(*v)


def _check(st):
    if st != 0:
        from numericals_b200 import _lib
        raise RuntimeError(_lib.lib().nuk_last_error())


if __name__ == "__main__":
    sys.exit(main())
