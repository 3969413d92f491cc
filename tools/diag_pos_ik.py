import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import humanoid_real_time_retarget_b200 as hrt
from oracle import retarget_oracle as oc
T = torch.from_numpy
sk = oc.load_skeletons()
g = dict(np.load('/root/repo/tests/golden/full_body_pos.npz'))
eng = hrt.default_engine(0)
body, lh, rh = T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"])
_, dof0, _ = eng.retarget_full_body_pos(body, lh, rh); dof0 = dof0.cpu()
_, dof1, _ = eng.retarget_full_body_pos(body, lh, rh, flags=1); dof1 = dof1.cpu()
rob_par = sk["hu_v5_zero_pose/parents"].tolist(); rob_off = T(sk["hu_v5_zero_pose/offsets"])
pos = torch.zeros(31, 3)
for j in range(1, 31): pos[j] = rob_off[j] + pos[rob_par[j]]
def resid(dof, wo=0.2):
    tot = torch.zeros(dof.shape[0])
    for first in (12, 21):
        d0 = first - 1
        pcu, Gu = oc.arm_chain(dof0[:, d0:d0+7], pos[first], rob_off[first:first+9])
        pc, G = oc.arm_chain(dof[:, d0:d0+7], pos[first], rob_off[first:first+9])
        qe = oc.quat_normalize(oc.quat_mul(Gu, oc.quat_conjugate(G)))
        ang = 2 * torch.atan2(qe[:, :3].norm(dim=-1), qe[:, 3])
        tot += ((pcu[3]-pc[3])**2).sum(-1) + ((pcu[6]-pc[6])**2).sum(-1) + (wo*ang)**2
    return tot.sqrt()
r1 = resid(dof1)
for it in (1, 3, 10, 30, 100):
    _, d, _ = eng.retarget_full_body_pos(body, lh, rh, flags=3, ik_iters=it)
    r = resid(d.cpu())
    bite = r1 > 1e-6
    print(it, "biting frames", int(bite.sum()), "mean resid clamp %.4f refined %.4f; frac improved %.3f worse %.3f" % (r1[bite].mean(), r[bite].mean(), (r[bite] < r1[bite]-1e-7).float().mean(), (r[bite] > r1[bite]+1e-6).float().mean()))
