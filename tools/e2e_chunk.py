"""Time the host-buffer call of the fused quaternion path (pinned in -> pinned dof + link positions) for the pipeline
chunk size in HRT_HOST_CHUNK_LOG2 (read once per process)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402

B = 1 << 20
sk = oc.load_skeletons()
raw = torch.empty((B, 21, 4)).pin_memory()
raw[:1 << 16] = oc.synth_clip_3q(1 << 16, seed=1, sk=sk)
for i in range(1, 16):
    raw[i << 16:(i + 1) << 16] = raw[:1 << 16]
dof = torch.empty((B, 30)).pin_memory()
lp = torch.empty((B, 31, 3)).pin_memory()
eng = hrt.default_engine(0)
fl = hrt.BQ_CLAMP | hrt.BQ_IK
for _ in range(3):
    eng.retarget_body_quat_host(raw, flags=fl, ik_iters=10, damping=0.1, rot_weight=0.2, out_dof=dof, out_link_pos=lp)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    eng.retarget_body_quat_host(raw, flags=fl, ik_iters=10, damping=0.1, rot_weight=0.2, out_dof=dof, out_link_pos=lp)
    ts.append(time.perf_counter() - t0)
ts.sort()
print(os.environ.get("HRT_HOST_CHUNK_LOG2", "16"), "median ms", round(ts[5] * 1e3, 3), "frames/s", round(B / ts[5] / 1e6, 2), "M")
