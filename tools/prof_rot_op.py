"""A few launches of one element-wise op on 2^24 rows, for `ncu --set full -k regex:rot_op -c 3`:  python tools/prof_rot_op.py quat_rotate"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from humanoid_real_time_retarget_b200 import rotation3d as r3d  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "quat_rotate"
eng = hrt.default_engine(0)
n = 1 << 24
gen = torch.Generator(device="cuda").manual_seed(0)
a = torch.nn.functional.normalize(torch.randn(n, 4, device="cuda", generator=gen), dim=-1)
v = torch.randn(n, 3, device="cuda", generator=gen)
o4, o3 = torch.empty(n, 4, device="cuda"), torch.empty(n, 3, device="cuda")
op, ins, outs = {"quat_rotate": (r3d.OP_QUAT_ROTATE, [a, v], [o3]), "quat_normalize": (r3d.OP_QUAT_NORMALIZE, [a], [o4]),
                 "quat_to_exp_map": (r3d.OP_QUAT_TO_EXP_MAP, [a], [o3])}[which]
for _ in range(4):
    eng.rot_op(op, n, ins, [0] * len(ins), outs)
torch.cuda.synchronize()
