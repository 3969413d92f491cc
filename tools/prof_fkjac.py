"""Two launches each of the FK (joint angles, 2^20 Hu configurations) and Jacobian (2^18, K = 2) kernels, for
`ncu --set full -k regex:"jacobian|fk_limb" -c 4` (a handful of launches keeps the report small)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402

eng = hrt.default_engine(0, robot="hu")
gen = torch.Generator(device="cuda").manual_seed(0)
L = 1 << 20
ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
rt = torch.randn(L, 3, device="cuda", generator=gen)
rq = torch.nn.functional.normalize(torch.randn(L, 4, device="cuda", generator=gen), dim=-1)
out = (torch.empty(L, 33, 4, device="cuda"), torch.empty(L, 33, 3, device="cuda"))
for _ in range(2):
    eng.fk_angles(hrt.TREE_ROBOT, ang, rt, rq, clip=True, exact=False, out=out)
jout = torch.empty(1 << 18, 2, 6, 32, device="cuda")
for _ in range(2):
    eng.fk_jacobian(hrt.TREE_ROBOT, ang[:1 << 18], [20, 29], clip=True, out=jout)
torch.cuda.synchronize()
print("ok")
