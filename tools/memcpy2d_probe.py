"""How fast does the copy engine move the NARROW column range of AoS rows across PCIe?  cudaMemcpy2DAsync of `width` bytes out
of every `pitch`-byte row, device <-> pinned host, against the contiguous copy of the whole rows (CUDA events)."""
import ctypes
import json
import sys

import torch

rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_int, ctypes.c_void_p]
H2D, D2H = 1, 2
n = 1 << 20


def timed(fn, k=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


st = torch.cuda.current_stream().cuda_stream
for name, pitch, off, width, kind in (("link_pos D2H", 372, 144, 216, D2H), ("dof D2H", 120, 44, 64, D2H), ("src quats H2D", 336, 160, 176, H2D)):
    dev = torch.zeros(n * pitch, dtype=torch.uint8, device="cuda")
    host = torch.zeros(n * pitch, dtype=torch.uint8).pin_memory()
    if kind == D2H:
        full = lambda: host.copy_(dev, non_blocking=True)
        part = lambda: rt.cudaMemcpy2DAsync(host.data_ptr() + off, pitch, dev.data_ptr() + off, pitch, width, n, D2H, st)
    else:
        full = lambda: dev.copy_(host, non_blocking=True)
        part = lambda: rt.cudaMemcpy2DAsync(dev.data_ptr() + off, pitch, host.data_ptr() + off, pitch, width, n, H2D, st)
    tf, tp = timed(full), timed(part)
    print(json.dumps({"case": name, "rows": n, "pitch": pitch, "width": width, "contiguous_ms": round(tf, 3), "contiguous_GBps": round(n * pitch / tf / 1e6, 1),
                      "columns_ms": round(tp, 3), "columns_payload_GBps": round(n * width / tp / 1e6, 1)}), flush=True)
