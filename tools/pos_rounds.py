"""Fixed cost of a position-path launch: CUDA-event time against the number of whole 16-frame rounds per warp
(k x 148 CTAs x 16 warps x 16 frames), dof only.   python tools/pos_rounds.py"""
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
gg = torch.Generator().manual_seed(3)
base = 148 * 16 * 16
em = 0.4 * torch.randn(base, 59, 3, generator=gg)
root = torch.zeros(base, 3)
root[:, 2] = 1.0
_, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                  torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in (1, 2, 3, 5, 7, 10, 20):
    body, lh, rh = (x.contiguous().cuda().repeat(k, 1, 1) for x in (gt[:, full2body], gt[:, 14:34], gt[:, 39:59]))
    dof = torch.empty(k * base, 30, device="cuda")
    ts = []
    for i in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.retarget_full_body_pos(body, lh, rh, out=(None, dof, None))
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    print(f"rounds {k:2d}: {np.median(ts):7.1f} us  ({np.median(ts) / k:6.1f} us per round)", flush=True)
