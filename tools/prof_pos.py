"""Two launches of the position-path kernel (config 3p, 2^18 frames, dof only) for `ncu --set full -k regex:pos_retarget -c 2`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402  (input synthesis only)

eng = hrt.default_engine(0)
sk = oc.load_skeletons()
gg = torch.Generator().manual_seed(0)
n = 1 << 18
em = 0.4 * torch.randn(n, 59, 3, generator=gg)
root = torch.zeros(n, 3)
root[:, 2] = 1.0
_, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                  torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
body, lh, rh = gt[:, full2body].contiguous().cuda(), gt[:, 14:34].contiguous().cuda(), gt[:, 39:59].contiguous().cuda()
dof = torch.empty(n, 30, device="cuda")
for _ in range(2):
    eng.retarget_full_body_pos(body, lh, rh, out=(None, dof, None))
torch.cuda.synchronize()
print("ok")
