"""Kernel duration of one-frame launches in situ: N launches queued back to back on one stream, one sync."""
import os, sys, time
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
outp = eng._pos_outputs(1, True, True, False)
dof = torch.empty(1, 30, device="cuda"); lp = torch.empty(1, 31, 3, device="cuda")
def bench(name, fn, n=3000):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{name:40s} device {a.elapsed_time(b)/n*1e3:7.2f} us/launch   host enqueue {(t1-t0)/n*1e6:6.2f} us/launch")
bench("pos B=1", lambda: eng.retarget_full_body_pos(body, lh, rh, out=outp))
bench("body_quat closed B=1", lambda: eng.retarget_body_quat(raw, flags=0, out=(None, dof, None)))
bench("body_quat ik10 B=1", lambda: eng.retarget_body_quat(raw, flags=3, out=(None, dof, lp)))
x = torch.zeros(1, device="cuda")
bench("torch add (launch floor)", lambda: x.add_(1))
