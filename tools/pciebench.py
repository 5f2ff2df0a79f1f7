#!/usr/bin/env python
"""tools/pciebench.py -- what the host link of this box can do: pinned H2D alone, D2H alone and both directions
at once, by chunk size (the ceiling of bench.py's e2e line, which moves 8 bytes per f32 element)."""
import time

import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, chunk):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for off in range(0, n, chunk):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out[off:off + chunk].copy_(d_b[off:off + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0) / 1e9


for chunk in (4 << 20, 32 << 20, 256 << 20, 1 << 30):
    for kind in ("h2d", "d2h", "both"):
        best = max(run(kind, chunk) for _ in range(3))
        print("chunk %4d MiB  %-5s %6.1f GB/s per direction" % (chunk >> 20, kind, best), flush=True)
