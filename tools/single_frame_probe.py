"""Probe: device time of ONE-frame launches (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanoid_real_time_retarget_b200 as hrt
from oracle import retarget_oracle as oc
sk = oc.load_skeletons()
eng = hrt.default_engine(0)
g = torch.Generator().manual_seed(0)
em = 0.4 * torch.randn(4, 59, 3, generator=g)
root = torch.zeros(4, 3); root[:, 2] = 1.0
_, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(), torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
body, lh, rh = gt[:1, full2body].contiguous().cuda(), gt[:1, 14:34].contiguous().cuda(), gt[:1, 39:59].contiguous().cuda()
raw = oc.synth_clip_3q(1, seed=1, sk=sk).cuda()
for _ in range(10):
    eng.retarget_full_body_pos(body, lh, rh, want_body_gq=False)
    eng.retarget_body_quat(raw, flags=0, want_local_q=False, want_link_pos=False)
    eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK, want_local_q=False)
torch.cuda.synchronize()
