"""Small ragged-size run of every kernel, meant to be run under `compute-sanitizer --tool memcheck`."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanoid_real_time_retarget_b200 as hrt
from oracle import retarget_oracle as oc
T = torch.from_numpy
sk = oc.load_skeletons()
eng, eng_hu = hrt.default_engine(0), hrt.default_engine(0, robot="hu")
g = torch.Generator().manual_seed(0)
for B in (1, 17, 33, 100):
    raw = oc.synth_clip_3q(B, seed=B, sk=sk)
    for flags in (0, hrt.BQ_CLAMP, hrt.BQ_CLAMP | hrt.BQ_IK, hrt.BQ_CLAMP | hrt.BQ_IK | hrt.BQ_ACTIVE_SET):
        eng.retarget_body_quat(raw, flags=flags, ik_iters=2)
        eng.retarget_body_quat(raw, flags=flags, ik_iters=2, want_local_q=False)
    eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK | hrt.BQ_PACKED_IK, ik_iters=2, want_local_q=False)
    em = 0.4 * torch.randn(B, 59, 3, generator=g)
    root = torch.zeros(B, 3); root[:, 2] = 1.0
    gq, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(), T(sk["vtrdyn_full_zero_pose/offsets"]))
    f2b = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, f2b].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
    eng.retarget_full_body_pos(body, lh, rh)
    eng.retarget_full_body_pos(body, lh, rh, flags=3, ik_iters=2, want_body_gq=False)
    eng.retarget_upper_body(body)
    eng.retarget_full_body(gq[:, f2b].contiguous(), body, lh, rh)
    eng.retarget_main_arms(gq[:, f2b].contiguous(), body)
    b23 = torch.zeros(B, 23, 3); b23[:, hrt.robot_config.BODY_23_TO_21] = body
    eng.retarget_full_body_pos_wire(b23, lh, rh)
    eng.rescale_motion(hrt.TREE_SOURCE, body, dir=[-1, -1, 1]); eng.rebuild_global_rotation(hrt.TREE_SOURCE, body)
    if B > 1:
        eng.motion_velocity(body, 1 / 30); eng.motion_angular_velocity(gq[:, f2b].contiguous(), 1 / 30)
    ang = torch.rand(B, 32, generator=g) - 0.5
    eng_hu.fk_angles(hrt.TREE_ROBOT, ang); eng_hu.fk_angles(hrt.TREE_ROBOT, ang, exact=True)
    eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29]); eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29, 5]); eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, [32])
    q = oc.quat_normalize(torch.randn(B, 21, 4, generator=g))
    eng.zero_pose_transform(hrt.TREE_SOURCE, q); eng.local_from_global(hrt.TREE_SOURCE, q); eng.fk_local_quats(hrt.TREE_SOURCE, q)
    r = hrt.rotation3d
    a, b, v = q.reshape(-1, 4), oc.quat_normalize(torch.randn(B * 21, 4, generator=g)), torch.randn(B * 21, 3, generator=g)
    r.quat_mul(a, b); r.quat_rotate(a, v); r.quat_to_exp_map(a); r.exp_map_to_quat(v); r.rot_matrix_from_quaternion(a)
    r.quat_from_rotation_matrix(r.rot_matrix_from_quaternion(a)); r.transform_mul(torch.cat([a, v], -1), torch.cat([b, v], -1))
    r.quat_mul(a[1:], b[1:]); r.quat_rotate(a[1:], v[1:])               # misaligned bases take the scalar path
    hrt.transform3d.quat_in_xyz_axis(a, "XYZ"); hrt.transform3d.cal_joint_quat(torch.randn(B, 5, 3, generator=g), torch.randn(B, 5, 3, generator=g))
    hrt.transform3d.quat_between_two_vecs(v, v.flip(0)); hrt.transform3d.quat_slerp(a, b, torch.rand(B * 21, 1, generator=g))
o = np.empty(30, np.float32)
for persistent in (False, True):
    eng.stream_pos_open(persistent=persistent)
    for i in range(5):
        eng.stream_pos_frame(body[i].numpy(), lh[i].numpy(), rh[i].numpy(), None, o)
    eng.stream_pos_close()
eng.stream_open(flags=3)
eng.stream_frame(raw[0].numpy(), None, o, None)
eng.stream_close()
torch.cuda.synchronize()
print("sanitize probe done")
