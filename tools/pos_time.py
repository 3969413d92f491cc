"""CUDA-event time of the position-path kernel on config 3p (2^18 frames, dof only, as bench.py runs it) and on 2^20 frames.
    python tools/pos_time.py [label]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402  (input synthesis only)

eng = hrt.default_engine(0)
sk = oc.load_skeletons()
full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
res = {}
for lg in (18, 20):
    n = 1 << lg
    gg = torch.Generator().manual_seed(3)
    em = 0.4 * torch.randn(1 << 18, 59, 3, generator=gg)
    root = torch.zeros(1 << 18, 3)
    root[:, 2] = 1.0
    _, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                      torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
    rep = n >> 18
    body, lh, rh = (x.contiguous().cuda().repeat(rep, 1, 1) for x in (gt[:, full2body], gt[:, 14:34], gt[:, 39:59]))
    dof = torch.empty(n, 30, device="cuda")
    evs = []
    torch.cuda.synchronize()
    for i in range(13):                                    # launches queued back to back, one event pair each (as bench.py)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.retarget_full_body_pos(body, lh, rh, out=(None, dof, None))
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs[3:]]
    res[f"2^{lg}_ms"] = round(float(np.median(ts)), 5)
    res[f"2^{lg}_hbm_frac"] = round(n * 852 / (float(np.median(ts)) * 1e-3) / 1e9 / 6448.7, 4)
    res[f"2^{lg}_checksum"] = float(dof.double().sum().item())
print(sys.argv[1] if len(sys.argv) > 1 else "", res)
