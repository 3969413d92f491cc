"""Pinned-memory PCIe ceilings of the box: D2H alone, H2D alone, both directions at once (two streams)."""
import time

import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both", True, True)):
    t = run(a, b)
    print(f"{name}: {n / t / 1e9:.1f} GB/s per direction ({t * 1e3:.2f} ms for 512 MiB each)")
