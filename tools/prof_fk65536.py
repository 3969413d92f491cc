"""configs[1] as the bench runs it: FK with limits and the 2-link Jacobian on 65,536 Hu configurations, L2 flushed before
every launch; for `ncu --set full -k regex:"jacobian|fk_limb" -c 6`.  Prints the CUDA-event times of the same launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402

eng = hrt.default_engine(0, robot="hu")
gen = torch.Generator(device="cuda").manual_seed(0)
L = 1 << 16
ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
rt = torch.randn(L, 3, device="cuda", generator=gen)
rq = torch.nn.functional.normalize(torch.randn(L, 4, device="cuda", generator=gen), dim=-1)
out = (torch.empty(L, 33, 4, device="cuda"), torch.empty(L, 33, 3, device="cuda"))
jout = torch.empty(L, 2, 6, 32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("fk", lambda: eng.fk_angles(hrt.TREE_ROBOT, ang, rt, rq, clip=True, exact=False, out=out)),
                 ("jac", lambda: eng.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29], clip=True, out=jout))):
    ts = []
    for i in range(3 if os.environ.get("HRT_PROF_SHORT") else 12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(round(a.elapsed_time(b) * 1e3, 2))
    print(name, "us per launch:", ts)
