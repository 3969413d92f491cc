"""Two launches of the local-quaternion FK kernel (2^20 Hu configurations) for `ncu --set full -k regex:fk_limb -c 2`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402

eng = hrt.default_engine(0, robot="hu")
gen = torch.Generator(device="cuda").manual_seed(0)
L = 1 << 20
lq = torch.nn.functional.normalize(torch.randn(L, 33, 4, device="cuda", generator=gen), dim=-1)
for _ in range(2):
    eng.fk_local_quats(hrt.TREE_ROBOT, lq)
torch.cuda.synchronize()
print("ok")
