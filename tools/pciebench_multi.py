#!/usr/bin/env python
"""tools/pciebench_multi.py -- the host-link ceiling of bench.py's e2e line as the number of GPUs grows.

One process drives G = 1, 2, 4, 8 GPUs at once: each GPU copies 1 GiB host->device and 1 GiB device->host
concurrently (page-locked buffers, 32 MiB chunks, one stream per direction) -- the traffic shape of the staged
BMAS_* path.  Prints GB/s each way per GPU and in aggregate, so the e2e figures of SCALE_rNN.json can be read
as a fraction of what the box's PCIe fabric / host memory delivers at that G.  Also: plain host memcpy
bandwidth by thread count (the ceiling of the pageable-operand path, which copies every byte once more)."""
import sys
import threading
import time

import numpy as np
import torch

ndev = torch.cuda.device_count()
n = 1 << 30
chunk = 32 << 20
bufs = []
for d in range(ndev):
    with torch.cuda.device(d):
        bufs.append(dict(h_in=torch.empty(n, dtype=torch.uint8).pin_memory(), h_out=torch.empty(n, dtype=torch.uint8).pin_memory(),
                         d_a=torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d),
                         d_b=torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d),
                         s1=torch.cuda.Stream(device=d), s2=torch.cuda.Stream(device=d)))


def run(G, kind):
    for d in range(G):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for off in range(0, n, chunk):
        for d in range(G):
            b = bufs[d]
            if kind in ("h2d", "both"):
                with torch.cuda.stream(b["s1"]):
                    b["d_a"][off:off + chunk].copy_(b["h_in"][off:off + chunk], non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(b["s2"]):
                    b["h_out"][off:off + chunk].copy_(b["d_b"][off:off + chunk], non_blocking=True)
    for d in range(G):
        torch.cuda.synchronize(d)
    return n / (time.perf_counter() - t0) / 1e9


print("# %d GPUs visible; 1 GiB per direction per GPU, 32 MiB chunks, page-locked host buffers" % ndev)
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    for kind in ("h2d", "d2h", "both"):
        best = max(run(G, kind) for _ in range(3))
        print("G=%d %-5s %6.1f GB/s each way per GPU  %7.1f GB/s each way aggregate" % (G, kind, best, best * G), flush=True)

src = np.ones(1 << 30, dtype=np.uint8)
dst = np.empty_like(src)
for T in (1, 2, 4, 8, 16):
    per = src.size // T

    def work(t):
        np.copyto(dst[t * per:(t + 1) * per], src[t * per:(t + 1) * per])

    best = 0.0
    for _ in range(3):
        ts = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        best = max(best, src.size / (time.perf_counter() - t0) / 1e9)
    print("host memcpy %2d threads: %6.1f GB/s copied (x2 = memory traffic)" % (T, best), flush=True)
sys.exit(0)
