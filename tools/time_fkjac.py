"""Device time per launch of the Hu FK (joint angles, limits) and the 2-link Jacobian at the configs[1] size and at larger
batches, timed as bench.py times its side measurements (event pair per launch, launches queued back to back, L2 flushed
in between).  One JSON line per case; for A/B runs of library variants:
    cp variants/<name>.so humanoid_real_time_retarget_b200/libhrt_b200.so && python tools/time_fkjac.py <name>"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "in-tree"
sizes = [int(s) for s in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["16", "18", "20"])]
peak = 6448.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
eng = hrt.default_engine(0, robot="hu")
gen = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def kernel_us(launch, reps=23, skip=3):
    evs = []
    torch.cuda.synchronize()
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs[skip:]]) * 1e3
    return float(np.median(t)), float(t.min())


for lg in sizes:
    L = 1 << lg
    ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
    out = (torch.empty(L, 33, 4, device="cuda"), torch.empty(L, 33, 3, device="cuda"))
    jout = torch.empty(L, 2, 6, 32, device="cuda")
    for name, nbytes, fn in (("fk_angles", 1080, lambda: eng.fk_angles(hrt.TREE_ROBOT, ang, clip=True, out=out)),
                             ("jacobian_K2", 1664, lambda: eng.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29], clip=True, out=jout))):
        med, mn = kernel_us(fn)
        print(json.dumps({"lib": tag, "case": name, "log2_configs": lg, "us_median": round(med, 2), "us_min": round(mn, 2),
                          "hbm_frac": round(L * nbytes / (med * 1e-6) / 1e9 / peak, 4)}), flush=True)
    del ang, out, jout
