"""Diagnostic: does the fused refinement decrease the IK objective frame by frame (quaternion path)?"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanoid_real_time_retarget_b200 as hrt
from oracle import retarget_oracle as oc
T = torch.from_numpy
sk = oc.load_skeletons()
eng = hrt.default_engine(0)
raw = oc.synth_clip_3q(20000, seed=5, sk=sk)
zq = oc.zero_pose_transform(raw, T(sk["t2z/vtrdyn"]))
_, d0, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP)
r0 = oc.ik_residual(d0.cpu(), zq, sk)
for it in (1, 3, 10, 30):
    _, d, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK, ik_iters=it)
    r = oc.ik_residual(d.cpu(), zq, sk)
    print(f"iters {it:3d}: mean residual {float(r0.mean()):.4f} -> {float(r.mean()):.4f}; improved {float((r < r0 - 1e-6).float().mean()):.3f} worse {float((r > r0 + 1e-6).float().mean()):.3f}; p99 worsening {float(np.quantile((r - r0).numpy(), 0.99)):.4f}")
lo, hi = torch.tensor(oc.HU_V5_DOF_LOWER), torch.tensor(oc.HU_V5_DOF_UPPER)
arm = list(range(11, 18)) + list(range(20, 27))
at = ((d.cpu()[:, arm] <= lo[arm]) | (d.cpu()[:, arm] >= hi[arm])).float().sum(-1)
print("mean # arm DOFs sitting at a limit after refinement:", float(at.mean()))
