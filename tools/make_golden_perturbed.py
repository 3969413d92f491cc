"""Goldens for solver instances built from NON-bundled zero poses (dev container only).

    python tools/make_golden_perturbed.py

The reference's solver classes read every offset from the RobotZeroPose objects handed to their
constructors (retarget/retarget_solver/full_body_pos_retargeter.py:69-107,139,162,184;
retarget_solver.py:49-86; full_body_retargeter.py:59-99,152; body_retargeter.py:35,38).  This script
perturbs the bundled source and target zero poses (limb lengths scaled 0.8-1.25x per joint, bone directions
turned by up to ~0.05 rad), runs the UNMODIFIED reference solvers built from them, and stores inputs,
perturbed tables and outputs in tests/golden/perturbed_zero_pose.npz.  The CUDA classes must reproduce
these from the same constructor arguments (tests/test_gpu_boundary.py).
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402
from make_golden import FULL2BODY, save  # noqa: E402


def perturb_zero_pose(ref, state, gen, scale_lo=0.8, scale_hi=1.25, turn=0.05):
    """A RobotZeroPose of the reference whose offsets are the bundled ones scaled and slightly turned."""
    rkm, r3d = ref.rkm, ref.r3d
    tree = state.skeleton_tree
    J = tree.num_joints
    off = tree.local_translation.clone()
    s = scale_lo + (scale_hi - scale_lo) * torch.rand(J, 1, generator=gen)
    rot = r3d.exp_map_to_quat(turn * torch.randn(J, 3, generator=gen))
    off = r3d.quat_rotate(rot, off) * s
    off[0] = tree.local_translation[0]
    parents = tree.parent_indices
    glob = off.clone()
    for j in range(1, J):
        glob[j] = off[j] + glob[int(parents[j])]
    new_tree = ref.sk3d.SkeletonTree(tree.node_names, parents.clone(), off.clone())
    return rkm.RobotZeroPose(local_translation=off, global_translation=glob, parent_indices=parents.clone(),
                             num_joints=J, node_names=tree.node_names, skeleton_tree=new_tree)


def main():
    ref = ref_shim.load()
    rkm, r3d = ref.rkm, ref.r3d
    g = torch.Generator().manual_seed(4321)
    src59 = perturb_zero_pose(ref, ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_full_zero_pose.pkl"), g)
    src21 = perturb_zero_pose(ref, ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl"), g)
    tgt = perturb_zero_pose(ref, ref_shim.load_asset(ref, "asset/hu_pose/hu_v5_zero_pose.pkl"), g)
    L = 96

    # position inputs: poses of the PERTURBED source skeletons (SURVEY 8(d) config 3p recipe)
    def clip(zp, scale, seed):
        gg = torch.Generator().manual_seed(seed)
        lq = r3d.exp_map_to_quat(scale * torch.randn(L, zp.num_joints, 3, generator=gg))
        root = torch.zeros(L, 3)
        root[:, 2] = 1.0
        return rkm.cal_forward_kinematics(lq, root, zp.parent_indices.tolist(), zp.local_translation)

    gq59, gt59 = clip(src59, 0.4, 21)
    body_t, lh, rh = gt59[:, FULL2BODY], gt59[:, 14:34], gt59[:, 39:59]
    body_q = gq59[:, FULL2BODY]
    out = dict(src59_offsets=src59.local_translation, src59_global_t=src59.global_translation,
               src59_parents=src59.parent_indices, src21_offsets=src21.local_translation,
               src21_global_t=src21.global_translation, src21_parents=src21.parent_indices,
               tgt_offsets=tgt.local_translation, tgt_global_t=tgt.global_translation, tgt_parents=tgt.parent_indices,
               body_t=body_t, lhand_t=lh, rhand_t=rh, body_q=body_q)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for precise in (True, False):
            s = ref.solvers.VtrdynFullBodyPosRetargeter(src59, tgt, precise_gripper=precise)
            a, b, c = [], [], []
            for i in range(L):
                x, y, z = s.retarget(body_t[i], lh[i], rh[i])
                a.append(x); b.append(y.clone()); c.append(z)
            tag = "precise" if precise else "binary"
            out[f"pos_{tag}_local_q"] = torch.stack(a)
            out[f"pos_{tag}_dof"] = torch.stack(b)
            out[f"pos_{tag}_body_gq"] = torch.stack(c)
            if precise:
                # BaseHumanoidRetargeter.motion_global_translation: FK on the TARGET zero pose's offsets
                out["pos_motion_global_t"] = s.motion_global_translation
                out["pos_motion_global_q"] = s.motion_global_rotation

        s = ref.solvers.VtrdynFullBodyRetargeter(src59, tgt)
        a, b = [], []
        for i in range(L):
            x, y = s.retarget(body_q[i], body_t[i], None, lh[i], None, rh[i])
            a.append(x); b.append(y.clone())
        out["full_local_q"], out["full_dof"] = torch.stack(a), torch.stack(b)

        _, gt21 = clip(src21, 0.4, 22)
        gt21 = gt21 * torch.tensor([-1., -1., 1.])
        s = ref.solvers.HuUpperBodyFromMocapRetarget(src21, tgt)
        a, b = [], []
        for i in range(L):
            x, y = s.retarget_from_global_translation(gt21[i])
            a.append(x); b.append(y)
        out["upper_global_t"], out["upper_local_q"], out["upper_dof"] = gt21, torch.stack(a), torch.stack(b)

        # quaternion path: only the source PARENTS and the target joint count are read (body_retargeter.py:35,38)
        gq21, _ = clip(src21, 0.5, 23)
        zq = ref.parse_mocap.vtrdyn_zero_pose_transform(gq21)
        s = ref.solvers.Mocap2HuBodyRetargeter(src21, tgt)
        a, b = [], []
        for i in range(L):
            x, y = s.retarget_from_pose(zq[i])
            a.append(x); b.append(y)
        out["bq_zero_pose_q"], out["bq_local_q"], out["bq_dof"] = zq, torch.stack(a), torch.stack(b)
        out["bq_motion_global_t"] = s.motion_global_translation
    save("perturbed_zero_pose", **out)


if __name__ == "__main__":
    main()
