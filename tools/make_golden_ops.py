"""Generate tests/golden/{rotation_ops2,skeleton_state,main_path,forward_vector}.npz by running the UNMODIFIED reference
(dev container only; same shim as tools/make_golden.py).

    python tools/make_golden_ops.py

rotation_ops2 : every remaining free function of rotation3d.py / transform3d.py on seeded inputs
skeleton_state: SkeletonState / SkeletonMotion of poselib (FK with non-identity tree.quat, global->local,
                velocities, retarget_to) on seeded inputs
main_path     : retarget/main.py RetargetHuV5fromMocap on a synthetic clip (rescale, rebuild, arm solves)
"""
import os
import pickle
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402
from make_golden import rand_quat, save  # noqa: E402


def ops(ref):
    r3d, t3d = ref.r3d, ref.t3d
    g = torch.Generator().manual_seed(4321)
    n = 384
    qa, qb = rand_quat(n, g), rand_quat(n, g)
    qa[::5] *= -1.0
    raw = torch.randn(n, 4, generator=g) * 2
    v = torch.randn(n, 3, generator=g)
    w = torch.randn(n, 3, generator=g)
    nn = torch.randn(n, 3, generator=g)
    ang = (torch.rand(n, generator=g) - 0.5) * 720.0
    ang_rad = (torch.rand(n, generator=g) - 0.5) * 20.0
    tr_a = torch.cat([qa, torch.randn(n, 3, generator=g)], dim=-1)
    tr_b = torch.cat([qb, torch.randn(n, 3, generator=g)], dim=-1)
    R = r3d.rot_matrix_from_quaternion(qa)
    tt = torch.rand(n, 1, generator=g)
    out = dict(qa=qa, qb=qb, raw=raw, v=v, w=w, nn=nn, ang=ang, ang_rad=ang_rad, tr_a=tr_a, tr_b=tr_b, R=R, tt=tt)
    out["quat_pos"] = r3d.quat_pos(raw)
    out["quat_abs"] = r3d.quat_abs(raw)
    out["quat_unit"] = r3d.quat_unit(raw)
    out["quat_conjugate"] = r3d.quat_conjugate(raw)
    out["quat_from_angle_axis_deg"] = r3d.quat_from_angle_axis(ang, v, degree=True)
    a2, ax2 = r3d.quat_angle_axis(qa.clone())
    out["quat_angle_axis_angle"], out["quat_angle_axis_axis"] = a2, ax2
    out["quat_yaw_z"] = r3d.quat_yaw_rotation(qa, z_up=True)
    out["quat_yaw_y"] = r3d.quat_yaw_rotation(qa, z_up=False)
    out["transform_inverse"] = r3d.transform_inverse(tr_a)
    out["transform_mul"] = r3d.transform_mul(tr_a, tr_b)
    out["transform_apply"] = r3d.transform_apply(tr_a, v)
    out["rot_matrix_det"] = r3d.rot_matrix_det(R)
    out["rot_matrix_from_quaternion"] = R
    out["rot_matrix_from_quaternion_raw"] = r3d.rot_matrix_from_quaternion(raw)
    for k, name in enumerate(["x", "y", "z", "xy", "xz"]):
        out[f"project_{name}"] = getattr(r3d, f"project_quat_to_axis_{name}")(qa)
    for k in range(3):
        out[f"extract_{k}"] = r3d.extract_rotation_along_axis(qa, k)
    out["normalize_angle"] = r3d.normalize_angle(ang_rad)
    qpos = r3d.quat_normalize(qa)
    a3, ax3 = r3d.quat_to_angle_axis(qpos)
    out["quat_to_angle_axis_angle"], out["quat_to_angle_axis_axis"] = a3, ax3
    a4, ax4 = r3d.exp_map_to_angle_axis(v)
    out["exp_map_to_angle_axis_angle"], out["exp_map_to_angle_axis_axis"] = a4, ax4
    out["angle_axis_to_exp_map"] = r3d.angle_axis_to_exp_map(ang_rad, v)
    out["quat_between_two_vecs"] = t3d.quat_between_two_vecs(v, w)
    out["quat_between_two_vecs_early"] = t3d.quat_between_two_vecs(v * 1e-8, w)
    out["proj_in_plane"] = torch.stack([t3d.proj_in_plane(v[i], nn[i]) for i in range(n)])
    out["radians_between_vecs"] = torch.stack([t3d.radians_between_vecs(v[i], w[i], nn[i]) for i in range(n)])
    qbn = r3d.quat_normalize(qb)
    out["quat_slerp"] = t3d.quat_slerp(qpos, qbn, tt)
    out["quat_slerp_near"] = t3d.quat_slerp(qpos, r3d.quat_normalize(qpos + 1e-5 * raw), tt)
    axes30 = list(ref.hu_v5_cfg.Hu_DOF_AXIS)
    q30 = r3d.quat_normalize(rand_quat(30, g))
    out["q30"], out["dof_axis30"] = q30, np.asarray(axes30, dtype=np.int64)
    out["quat_to_dof_pos"] = t3d.quat_to_dof_pos(q30, axes30)
    out["coord_transform"] = t3d.coord_transform(v, order=[2, 0, 1], dir=torch.Tensor([-1, 1, -1]))
    for seq in ["xyz", "zyx", "XYZ", "YXZ", "ZYX", "ZXZ", "yzy", "XZY"]:
        rows = [t3d.quat_in_xyz_axis(qa[i], seq) for i in range(n)]       # SciPy 1.18 takes one rotation per call here
        for m in range(3):
            out[f"euler_{seq}_{m + 1}"] = torch.stack([r[m] for r in rows])
    out["quat_to_eular"] = r3d.quat_to_eular(qa.numpy())
    out["cal_joint_quat_n4"] = t3d.cal_joint_quat(torch.randn(n, 4, 3, generator=torch.Generator().manual_seed(5)),
                                                  torch.randn(n, 4, 3, generator=torch.Generator().manual_seed(6)))
    save("rotation_ops2", **out)


def skeleton(ref):
    sk3d, r3d = ref.sk3d, ref.r3d
    g = torch.Generator().manual_seed(77)
    zp = ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl")
    base = zp.skeleton_tree
    J = base.num_joints
    tq = r3d.quat_normalize(rand_quat(J, g) * torch.tensor([0.3, 0.3, 0.3, 1.0]))
    tree = sk3d.SkeletonTree(base.node_names, base.parent_indices, base.local_translation.clone(), tq)
    T = 40
    em = 0.3 * torch.randn(T, J, 3, generator=g)
    walk = torch.cumsum(0.05 * torch.randn(T, J, 3, generator=g), dim=0)
    lq = r3d.exp_map_to_quat(em + walk)
    rt = torch.cumsum(0.02 * torch.randn(T, 3, generator=g), dim=0)
    st = sk3d.SkeletonState.from_rotation_and_root_translation(tree, lq, rt, is_local=True)
    out = dict(tree_quat=tq, parents=base.parent_indices, offsets=base.local_translation, local_q=lq, root_t=rt,
               state_tensor=st.tensor, global_rotation=st.global_rotation, global_translation=st.global_translation)
    st_g = sk3d.SkeletonState.from_rotation_and_root_translation(tree, st.global_rotation.clone(), rt, is_local=False)
    out["local_back"] = st_g.local_rotation
    out["global_repr_tensor"] = st.global_repr().tensor
    mot = sk3d.SkeletonMotion.from_skeleton_state(st, fps=30)
    out["global_velocity"], out["global_angular_velocity"] = mot.global_velocity, mot.global_angular_velocity
    out["velocity_nofilter"] = sk3d.SkeletonMotion._compute_velocity(st.global_translation, 1 / 30, guassian_filter=False)
    out["angular_velocity_nofilter"] = sk3d.SkeletonMotion._compute_angular_velocity(st.global_rotation, 1 / 30, guassian_filter=False)
    out["zero_pose_tensor"] = sk3d.SkeletonState.zero_pose(tree).tensor
    out["zero_pose_global_translation"] = sk3d.SkeletonState.zero_pose(tree).global_translation
    # naive T-pose-relative retarget (skeleton3d.py:742-889): vtrdyn clip -> Hu tree with the mapping dict of Hu.py
    hu = ref_shim.load_asset(ref, "asset/zero_pose/hu_zero_pose.pkl")      # the 33-joint Hu has the mapping's 'neck_link'
    tp = ref_shim.load_asset(ref, "asset/t_pose/vtrdyn_t_pose.pkl")
    mapping = ref.hu_cfg.VTRDYN2HU_JOINT_MAPPING
    src = sk3d.SkeletonState.from_rotation_and_root_translation(tp.skeleton_tree, r3d.exp_map_to_quat(em[:12]), rt[:12], is_local=True)
    rot = r3d.quat_from_angle_axis(torch.tensor(0.3), torch.tensor([0., 0., 1.]))
    res = src.retarget_to(mapping, tp.local_rotation, tp.root_translation, hu.skeleton_tree, hu.local_rotation, hu.root_translation,
                          rotation_to_target_skeleton=rot, scale_to_target_skeleton=0.9)
    out["rt_src_local_q"], out["rt_src_root_t"], out["rt_rot"] = src.local_rotation, src.root_translation, rot
    out["rt_out_tensor"], out["rt_out_is_local"] = res.tensor, np.asarray(res.is_local)
    out["rt_mapping_keys"] = np.asarray(list(mapping.keys()))
    out["rt_mapping_vals"] = np.asarray(list(mapping.values()))
    save("skeleton_state", **out)


def main_path(ref):
    """retarget/main.py through the shim: the file imports a few stale names."""
    import importlib
    import types
    from unittest import mock
    sys.modules["retarget.robot_kinematics_model"] = ref.rkm
    import retarget.utils as ru
    ru.get_mocap_translation = mock.MagicMock()
    cwd = os.getcwd()
    os.chdir(ref.root)
    try:
        main = importlib.import_module("retarget.main")
    finally:
        os.chdir(cwd)
    zp = ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl")
    hu = ref_shim.load_asset(ref, "asset/hu_pose/hu_v5_zero_pose.pkl")
    src = ref.rkm.RobotZeroPose.from_skeleton_state(zp)
    tgt = ref.rkm.RobotZeroPose.from_skeleton_state(hu)
    # synthetic clip: FK of random local rotations on the zero-pose tree, limbs stretched by a per-bone factor
    g = torch.Generator().manual_seed(99)
    L, J = 48, 21
    em = 0.35 * torch.randn(L, J, 3, generator=g)
    lq = ref.r3d.exp_map_to_quat(em)
    tree = zp.skeleton_tree
    stretch = 1.0 + 0.2 * torch.rand(J, 1, generator=g)
    root = torch.zeros(L, 3)
    root[:, 2] = 0.9
    _, gt = ref.rkm.cal_forward_kinematics(lq, root, tree.parent_indices.tolist(), tree.local_translation * stretch)
    gt = gt * torch.tensor([-1.0, -1.0, 1.0])          # the solver flips x,y back (main.py:170)
    r = main.RetargetHuV5fromMocap(src, tgt)
    flipped = ref.t3d.coord_transform(gt, dir=torch.Tensor([-1, -1, 1]))
    rescaled = main.Retarget.rescale_motion_to_standard_size(flipped, src)
    rebuilt = r._rebuild_with_vtrdyn_zero_pose(rescaled)
    main.plot_skeleton_H = mock.MagicMock()
    r.retarget_from_global_translation(gt)
    mocap_motion, retargeted = main.plot_skeleton_H.call_args[0][0]
    save("main_path", global_t=gt, rescaled=rescaled, rebuilt_global_rotation=rebuilt.global_rotation,
         rebuilt_global_translation=rebuilt.global_translation, rebuilt_local_rotation=rebuilt.local_rotation,
         rebuilt_velocity=rebuilt.global_velocity, rebuilt_angular_velocity=rebuilt.global_angular_velocity,
         robot_local_rotation=retargeted.local_rotation, robot_global_translation=retargeted.global_translation,
         robot_tensor=retargeted.tensor)


def forward_vector(ref):
    """SkeletonState.compute_forward_vector (skeleton3d.py:542-566) on a turning, walking vtrdyn clip; a clip shorter
    than the filter radius (80 frames at the default width 20) and a narrower filter as edge cases."""
    sk3d, r3d = ref.sk3d, ref.r3d
    g = torch.Generator().manual_seed(78)
    tree = ref_shim.load_asset(ref, "asset/t_pose/vtrdyn_t_pose.pkl").skeleton_tree
    J = tree.num_joints
    T = 300
    lq = r3d.exp_map_to_quat(0.2 * torch.randn(T, J, 3, generator=g) + torch.cumsum(0.03 * torch.randn(T, J, 3, generator=g), dim=0))
    rt = torch.cumsum(0.02 * torch.randn(T, 3, generator=g), dim=0)
    st = sk3d.SkeletonState.from_rotation_and_root_translation(tree, lq, rt, is_local=True)
    idx = (17, 13, 4, 1)                       # vtrdyn: left / right shoulder, left / right hip (VTRDYN.py:2-26)
    out = dict(local_q=lq, root_t=rt, idx=np.asarray(idx), global_translation=st.global_translation,
               fwd_default=st.compute_forward_vector(*idx), fwd_width5=st.compute_forward_vector(*idx, gaussian_filter_width=5))
    short = sk3d.SkeletonState.from_rotation_and_root_translation(tree, lq[:30], rt[:30], is_local=True)
    out["fwd_short"] = short.compute_forward_vector(*idx)
    save("forward_vector", **out)


if __name__ == "__main__":
    ref = ref_shim.load()
    which = sys.argv[1:] or ["ops", "skeleton", "main", "forward"]
    if "forward" in which:
        forward_vector(ref)
    if "ops" in which:
        ops(ref)
    if "skeleton" in which:
        skeleton(ref)
    if "main" in which:
        main_path(ref)
