"""the UNMODIFIED reference timed on CPU (dev container only: /root/reference does not exist on
the GPU box).  Imports the reference through tools/ref_shim.py, runs its own per-frame Python loops / batched calls on the
synthetic clips of SURVEY 8(d), and writes profiles/r01_reference_cpu_devbox.json.

    python tools/time_reference_cpu.py [--frames 2000]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ref_shim  # noqa: E402
from make_golden import FULL2BODY, clip_3p, clip_3q  # noqa: E402


def pct(ts):
    ts = np.asarray(ts) * 1e3
    return {"frames": len(ts), "frames_per_s": float(len(ts) / (ts.sum() / 1e3)), "ms_p50": float(np.percentile(ts, 50)),
            "ms_p99": float(np.percentile(ts, 99)), "ms_mean": float(ts.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2000)
    args = ap.parse_args()
    ref = ref_shim.load()
    rkm, pm = ref.rkm, ref.parse_mocap
    warnings.simplefilter("ignore")
    out = {"host": {"cpus": os.cpu_count(), "torch": torch.__version__, "note": "dev container (no GPU); the GPU box's host differs"},
           "what": "unmodified reference imported from /root/reference through tools/ref_shim.py"}
    src21 = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl"))
    src59 = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_full_zero_pose.pkl"))
    tgt = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/hu_pose/hu_v5_zero_pose.pkl"))
    L, W = args.frames, 200
    for threads in (1, os.cpu_count()):
        torch.set_num_threads(threads)
        key = f"threads_{threads}"
        res = {}
        # C1 (q): zero-pose transform batched once, then Mocap2HuBodyRetargeter.retarget_from_pose per frame
        raw = clip_3q(ref, L + W)
        t0 = time.perf_counter()
        zq = pm.vtrdyn_zero_pose_transform(raw)
        res["C1q_zero_pose_transform_batched_ms"] = (time.perf_counter() - t0) * 1e3
        solver = ref.solvers.Mocap2HuBodyRetargeter(src21, tgt)
        ts = []
        for i in range(L + W):
            t0 = time.perf_counter()
            solver.retarget_from_pose(zq[i])
            ts.append(time.perf_counter() - t0)
        res["C1q_Mocap2HuBodyRetargeter_retarget_from_pose"] = pct(ts[W:])
        # C1 (p): VtrdynFullBodyPosRetargeter.retarget (precise_gripper=True) per frame
        _, gt59 = clip_3p(ref, L + W)
        body_t, lh, rh = gt59[:, FULL2BODY], gt59[:, 14:34], gt59[:, 39:59]
        s = ref.solvers.VtrdynFullBodyPosRetargeter(src59, tgt, precise_gripper=True)
        ts = []
        for i in range(L + W):
            t0 = time.perf_counter()
            s.retarget(body_t[i], lh[i], rh[i])
            ts.append(time.perf_counter() - t0)
        res["C1p_VtrdynFullBodyPosRetargeter_retarget"] = pct(ts[W:])
        # C2: HuForwardModel.forward_kinematics, 65,536 configurations, limits on
        hu = ref_shim.load_asset(ref, "asset/zero_pose/hu_zero_pose.pkl")
        model = ref.hfm.HuForwardModel(hu.skeleton_tree, device="cpu")
        g = torch.Generator().manual_seed(0)
        B = 65536
        lo, hi = ref.hu_cfg.Hu_DOF_LOWER, ref.hu_cfg.Hu_DOF_UPPER
        ang = (lo + (hi - lo) * (torch.rand(B, 32, generator=g) * 1.2 - 0.1)).reshape(B, 32, 1)
        root_t = torch.zeros(B, 3)
        root_q = torch.zeros(B, 1, 4)
        root_q[..., 3] = 1.0
        best = 1e9
        for it in range(7):
            t0 = time.perf_counter()
            model.forward_kinematics(ang, root_t, root_q, clip_angles=True)
            dt = time.perf_counter() - t0
            if it >= 2:
                best = min(best, dt)
        res["C2_HuForwardModel_forward_kinematics_65536"] = {"ms_best_of_5": best * 1e3, "configs_per_s": B / best}
        # C3: cal_local_rotation, 65,536 frames of 21 joints
        gq = ref.r3d.quat_normalize(torch.randn(B, 21, 4, generator=g))
        best = 1e9
        for it in range(7):
            t0 = time.perf_counter()
            rkm.cal_local_rotation(gq, ref.vtrdyn_cfg.vtrdyn_parent_indices if hasattr(ref.vtrdyn_cfg, "vtrdyn_parent_indices")
                                   else src21.parent_indices.tolist())
            dt = time.perf_counter() - t0
            if it >= 2:
                best = min(best, dt)
        res["C3_cal_local_rotation_65536x21"] = {"ms_best_of_5": best * 1e3, "frames_per_s": B / best}
        out[key] = res
        print(key, json.dumps(res)[:600], flush=True)
    q = out["threads_1"]["C1q_Mocap2HuBodyRetargeter_retarget_from_pose"]["frames_per_s"]
    p = out["threads_1"]["C1p_VtrdynFullBodyPosRetargeter_retarget"]["frames_per_s"]
    out["extrapolation_1M_frames_hours"] = {"quaternion_path": (1 << 20) / q / 3600, "position_path": (1 << 20) / p / 3600,
                                            "note": "stated, not run (BASELINE.md C1)"}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_reference_cpu_devbox.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
