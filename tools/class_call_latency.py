"""Per-frame latency of the reference-named classes called the way the teleop scripts call them (CPU tensors, one frame)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402  (input synthesis only)

sk = oc.load_skeletons()
g = torch.Generator().manual_seed(0)
n = 2048
em = 0.4 * torch.randn(n, 59, 3, generator=g)
root = torch.zeros(n, 3)
root[:, 2] = 1.0
_, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                  torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
body, lh, rh = gt[:, full2body].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
src = hrt.RobotZeroPose.from_asset("vtrdyn_full_zero_pose")
tgt = hrt.RobotZeroPose.from_asset("hu_v5_zero_pose")
solver = hrt.VtrdynFullBodyPosRetargeter(src, tgt, precise_gripper=True)
for i in range(200):
    solver.retarget(body[i], lh[i], rh[i], record=False)
ts = np.empty(5000)
for i in range(5000):
    k = i % n
    t0 = time.perf_counter_ns()
    rl, dof, bq = solver.retarget(body[k], lh[k], rh[k], record=False)
    ts[i] = time.perf_counter_ns() - t0
print("VtrdynFullBodyPosRetargeter.retarget per frame (CPU tensors in/out): p50 %.1f us  p99 %.1f us" %
      (np.percentile(ts, 50) / 1e3, np.percentile(ts, 99) / 1e3))

# quaternion path: Mocap2HuBodyRetargeter.retarget_from_pose on one pre-transformed frame
raw = oc.synth_clip_3q(2048, seed=3, sk=sk)
zq = hrt.vtrdyn_zero_pose_transform(raw)
mocap = hrt.RobotZeroPose.from_asset("vtrdyn_zero_pose")
bsolver = hrt.Mocap2HuBodyRetargeter(mocap, tgt)
for i in range(200):
    bsolver.retarget_from_pose(zq[i], record=False)
for i in range(5000):
    k = i % 2048
    t0 = time.perf_counter_ns()
    bsolver.retarget_from_pose(zq[k], record=False)
    ts[i] = time.perf_counter_ns() - t0
print("Mocap2HuBodyRetargeter.retarget_from_pose per frame (CPU tensors in/out): p50 %.1f us  p99 %.1f us" %
      (np.percentile(ts, 50) / 1e3, np.percentile(ts, 99) / 1e3))

us = hrt.HuUpperBodyFromMocapRetarget(mocap, tgt)
for i in range(200):
    us.retarget_from_global_translation(body[i], record=False)
for i in range(5000):
    k = i % n
    t0 = time.perf_counter_ns()
    us.retarget_from_global_translation(body[k], record=False)
    ts[i] = time.perf_counter_ns() - t0
print("HuUpperBodyFromMocapRetarget.retarget_from_global_translation per frame: p50 %.1f us  p99 %.1f us" %
      (np.percentile(ts, 50) / 1e3, np.percentile(ts, 99) / 1e3))
