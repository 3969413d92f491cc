"""Two launches each of the kernels beside the headline one -- FK from joint angles (2^20 Hu configurations), FK from local
quaternions (2^20), Jacobian (2^18, K = 2) and the position path (2^18 frames, dof only) -- for ONE
`ncu --set full -k regex:"fk_limb|jacobian|pos_retarget" -c 8` capture (profiles/r01_side_kernels_ncu_summary.csv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402  (input synthesis only)

eng_hu = hrt.default_engine(0, robot="hu")
eng = hrt.default_engine(0)
gen = torch.Generator(device="cuda").manual_seed(0)
L = 1 << 20
ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
rt = torch.randn(L, 3, device="cuda", generator=gen)
rq = torch.nn.functional.normalize(torch.randn(L, 4, device="cuda", generator=gen), dim=-1)
out = (torch.empty(L, 33, 4, device="cuda"), torch.empty(L, 33, 3, device="cuda"))
for _ in range(2):
    eng_hu.fk_angles(hrt.TREE_ROBOT, ang, rt, rq, clip=True, exact=False, out=out)
lq = torch.nn.functional.normalize(torch.randn(L, 33, 4, device="cuda", generator=gen), dim=-1)
for _ in range(2):
    eng_hu.fk_local_quats(hrt.TREE_ROBOT, lq)
jout = torch.empty(1 << 18, 2, 6, 32, device="cuda")
for _ in range(2):
    eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang[:1 << 18], [20, 29], clip=True, out=jout)
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
