"""CPU oracle for the mocap -> humanoid retarget hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (humanoid_real_time_retarget_b200/) never does: it fails loudly
when the CUDA library is missing.

What this is: a *batched* torch-CPU fp32 restatement (leading frame axis B) of the reference's
per-frame Python, written from the reference's behaviour, each function citing the reference
file:line it follows (paths relative to the reference root).  The reference solvers cannot batch
(torch.dot / 1-D only), so a restatement is needed to check millions of frames in seconds.

Parity pinning (see tests/test_oracle_vs_golden.py, tools/make_golden.py):
  * every function below that has a reference counterpart is checked against outputs of the
    UNMODIFIED reference, imported in the dev container through tools/ref_shim.py, stored as
    fixtures under tests/golden/*.npz together with the generating script;
  * the reference's own known answers are pinned too: poselib test_rotation.py:27-37 and
    retarget/rotation_test.py:96-152;
  * the Euler split restates SciPy's algorithm (scipy.spatial.transform, un-vendored, unpinned by
    the reference; goldens generated with scipy 1.18.1): Bernardes & Viollet 2022 as implemented
    in scipy/spatial/transform/_rotation_xp.py:365-401,1052-1118.
  * Jacobian and IK refinement have NO reference implementation: **parity
    unpinned** for those two; the spec is ours and the oracle is the spec.
"""
from __future__ import annotations

import contextlib
import math
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                     "humanoid_real_time_retarget_b200", "data", "skeletons.npz")


def load_skeletons():
    """The extracted asset tables (tools/extract_assets.py)."""
    z = np.load(_DATA, allow_pickle=False)
    return {k: z[k] for k in z.files}


# ----------------------------------------------------------------------------------------------
# Static tables (retarget/robot_config/Hu_v5.py:12-18, Hu.py:4-25) -- must be bit-exact
# ----------------------------------------------------------------------------------------------
HU_V5_DOF_AXIS = [2, 0, 1, 1, 1, 2, 0, 1, 1, 1, 2,
                  1, 0, 2, 1, 0, 1, 2, 1, 1,
                  1, 0, 2, 1, 0, 1, 2, 1, 1, 2]
HU_DOF_AXIS = [2, 0, 1, 1, 1, 0, 2, 0, 1, 1, 1, 0, 2,
               1, 0, 2, 1, 0, 1, 2, 1, 1,
               1, 0, 2, 1, 0, 1, 2, 1, 1, 2]
HU_DOF_LOWER = [-0.1745, -0.3491, -1.5708, 0.0997, -0.6981, -0.3665,
                -0.1745, -0.3491, -1.5708, 0.0997, -0.6981, -0.3665,
                -1.0472,
                -3.1416, 0., -1.5708, 0., -1.5708, -0.785, -0.7854, 0., -0.044,
                -3.1416, -1.5708, -1.5708, 0., -1.5708, -0.785, -0.7854, 0., -0.044,
                -1.]
HU_DOF_UPPER = [0.1745, 0.3491, 0.8727, 2.618, 0.6981, 0.3665,
                0.1745, 0.3491, 0.8727, 2.618, 0.6981, 0.3665,
                1.0472,
                1.0472, 1.5708, 1.5708, 1.5708, 1.5708, 0.785, 0.7854, 0.044, 0.,
                1.0472, 0., 1.5708, 1.5708, 1.5708, 0.785, 0.7854, 0.044, 0.,
                1.]
# Hu v5 limits: the reference ships no consistent 30-entry table; ours = the
# 32-entry Hu table with the two toe DOFs (entries 5 and 11) removed.
HU_V5_DOF_LOWER = [v for i, v in enumerate(HU_DOF_LOWER) if i not in (5, 11)]
HU_V5_DOF_UPPER = [v for i, v in enumerate(HU_DOF_UPPER) if i not in (5, 11)]


# ----------------------------------------------------------------------------------------------
# poselib/poselib/core/rotation3d.py restated (xyzw quaternions, fp32, one torch op per
# reference op so the rounding sequence is the same)
# ----------------------------------------------------------------------------------------------
def quat_mul(a, b):
    """rotation3d.py:15-27."""
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2
    z = w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2
    return torch.stack([x, y, z, w], dim=-1)


def quat_pos(q):
    """rotation3d.py:31-38: flip sign where w < 0."""
    z = (q[..., 3:] < 0).float()
    return (1 - 2 * z) * q


def quat_unit(q):
    """rotation3d.py:51-56."""
    norm = q.norm(p=2, dim=-1).unsqueeze(-1)
    return q / norm.clamp(min=1e-9)


def quat_normalize(q):
    """rotation3d.py:93-98: sign-flip THEN normalise."""
    return quat_unit(quat_pos(q))


def quat_conjugate(q):
    """rotation3d.py:60-64 (quat_inverse :215-219 is the same thing)."""
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


quat_inverse = quat_conjugate


def quat_mul_norm(a, b):
    """rotation3d.py:197-202."""
    return quat_normalize(quat_mul(a, b))


def quat_rotate(rot, vec):
    """rotation3d.py:206-211: imag(q * (v,0) * conj(q)) through two full products."""
    other = torch.cat([vec, torch.zeros_like(vec[..., :1])], dim=-1)
    return quat_mul(quat_mul(rot, other), quat_conjugate(rot))[..., :3]


def quat_identity(shape: Sequence[int]):
    """rotation3d.py:112-119."""
    q = torch.zeros(tuple(shape) + (4,), dtype=torch.float32)
    q[..., 3] = 1
    return q


def quat_from_angle_axis(angle, axis):
    """rotation3d.py:123-143."""
    theta = (angle / 2).unsqueeze(-1)
    axis = axis / axis.norm(p=2, dim=-1, keepdim=True).clamp(min=1e-9)
    xyz = axis * theta.sin()
    w = theta.cos()
    return quat_normalize(torch.cat([xyz, w], dim=-1))


def quat_from_rotation_matrix(m):
    """rotation3d.py:147-193, including the four SEQUENTIAL masked sign fix-ups whose
    comparisons see already-mutated values and sign(0) == 0."""
    m = m.unsqueeze(0)
    d0, d1, d2 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    w = (((d0 + d1 + d2 + 1.0) / 4.0).clamp(0.0, None)) ** 0.5
    x = (((d0 - d1 - d2 + 1.0) / 4.0).clamp(0.0, None)) ** 0.5
    y = (((-d0 + d1 - d2 + 1.0) / 4.0).clamp(0.0, None)) ** 0.5
    z = (((-d0 - d1 + d2 + 1.0) / 4.0).clamp(0.0, None)) ** 0.5
    c0 = (w >= x) & (w >= y) & (w >= z)
    x = torch.where(c0, x * (m[..., 2, 1] - m[..., 1, 2]).sign(), x)
    y = torch.where(c0, y * (m[..., 0, 2] - m[..., 2, 0]).sign(), y)
    z = torch.where(c0, z * (m[..., 1, 0] - m[..., 0, 1]).sign(), z)
    c1 = (x >= w) & (x >= y) & (x >= z)
    w = torch.where(c1, w * (m[..., 2, 1] - m[..., 1, 2]).sign(), w)
    y = torch.where(c1, y * (m[..., 1, 0] + m[..., 0, 1]).sign(), y)
    z = torch.where(c1, z * (m[..., 0, 2] + m[..., 2, 0]).sign(), z)
    c2 = (y >= w) & (y >= x) & (y >= z)
    w = torch.where(c2, w * (m[..., 0, 2] - m[..., 2, 0]).sign(), w)
    x = torch.where(c2, x * (m[..., 1, 0] + m[..., 0, 1]).sign(), x)
    z = torch.where(c2, z * (m[..., 2, 1] + m[..., 1, 2]).sign(), z)
    c3 = (z >= w) & (z >= x) & (z >= y)
    w = torch.where(c3, w * (m[..., 1, 0] - m[..., 0, 1]).sign(), w)
    x = torch.where(c3, x * (m[..., 2, 0] + m[..., 0, 2]).sign(), x)
    y = torch.where(c3, y * (m[..., 2, 1] + m[..., 1, 2]).sign(), y)
    return quat_normalize(torch.stack([x, y, z, w], dim=-1)).squeeze(0)


def normalize_angle(x):
    """rotation3d.py:583-584."""
    return torch.atan2(torch.sin(x), torch.cos(x))


def quat_to_angle_axis(q):
    """rotation3d.py:588-608."""
    min_theta = 1e-5
    sin_theta = torch.sqrt(1 - q[..., 3] * q[..., 3])
    angle = 2 * torch.acos(q[..., 3])
    angle = normalize_angle(angle)
    axis = q[..., 0:3] / sin_theta.unsqueeze(-1)
    mask = torch.abs(sin_theta) > min_theta
    default_axis = torch.zeros_like(axis)
    default_axis[..., -1] = 1
    angle = torch.where(mask, angle, torch.zeros_like(angle))
    axis = torch.where(mask.unsqueeze(-1), axis, default_axis)
    return angle, axis


def quat_to_exp_map(q):
    """rotation3d.py:621-627."""
    angle, axis = quat_to_angle_axis(q)
    return angle.unsqueeze(-1) * axis


def exp_map_to_angle_axis(exp_map):
    """rotation3d.py:630-646."""
    min_theta = 1e-5
    angle = torch.norm(exp_map, dim=-1)
    axis = exp_map / angle.unsqueeze(-1)
    angle = normalize_angle(angle)
    default_axis = torch.zeros_like(exp_map)
    default_axis[..., -1] = 1
    mask = torch.abs(angle) > min_theta
    angle = torch.where(mask, angle, torch.zeros_like(angle))
    axis = torch.where(mask.unsqueeze(-1), axis, default_axis)
    return angle, axis


def exp_map_to_quat(exp_map):
    """rotation3d.py:649-652 / transform3d.py:147-150."""
    angle, axis = exp_map_to_angle_axis(exp_map)
    return quat_from_angle_axis(angle, axis)


def quat_to_dof_pos(quat, dof_axis: List[int]):
    """transform3d.py:177-183: exp_map(q)[..., i, dof_axis[i]].  quat: (..., D, 4)."""
    eye = torch.eye(3)[dof_axis]                       # (D,3)
    exp_map = quat_to_exp_map(quat) * eye              # reference multiplies before gathering
    idx = torch.tensor(dof_axis, dtype=torch.long)
    return exp_map[..., torch.arange(len(dof_axis)), idx]


# ----------------------------------------------------------------------------------------------
# robot_kinematics_model/kinematics.py
# ----------------------------------------------------------------------------------------------
def cal_forward_kinematics(local_q, root_t, parents: Sequence[int], offsets):
    """kinematics.py:13-39.  local_q (B,J,4), root_t (B,3), offsets (J,3).
    Root: G_r[0] = l[0] (NOT normalised), G_t[0] = root_t (offset[0] ignored)."""
    gr, gt = [], []
    for j, p in enumerate(parents):
        if p == -1:
            gr.append(local_q[..., j, :])
            gt.append(root_t)
        else:
            gr.append(quat_mul_norm(gr[p], local_q[..., j, :]))
            gt.append(quat_rotate(gr[p], offsets[j, :]) + gt[p])
    return torch.stack(gr, dim=-2), torch.stack(gt, dim=-2)


def cal_local_rotation(global_q, parents: Sequence[int]):
    """kinematics.py:41-63."""
    out = quat_identity(global_q.shape[:-1])
    for j, p in enumerate(parents):
        if p == -1:
            out[..., j, :] = global_q[..., j, :]
        else:
            out[..., j, :] = quat_mul_norm(quat_inverse(global_q[..., p, :]), global_q[..., j, :])
    return out


def hu_forward_kinematics(angles, root_t, root_q, parents, offsets, dof_axis, lower, upper,
                          clip_angles: bool):
    """hu_forward_model.py:17-33.  angles (L,D,1); root_q (L,1,4).  The forward VALUE of the
    straight-through clamp is (clamp(x) - x) + x, which is not always bit-equal to clamp(x)."""
    L = angles.shape[0]
    D = len(dof_axis)
    if clip_angles:
        lo = torch.tensor(lower, dtype=torch.float32).reshape(1, -1, 1)
        hi = torch.tensor(upper, dtype=torch.float32).reshape(1, -1, 1)
        clamped = torch.clamp(angles.clone(), min=lo, max=hi)
        # (clamp(x) - x).detach() + x  (hu_forward_model.py:27-33): the forward value is the clamped angle, the gradient
        # passes straight through
        angles = (clamped - angles).detach() + angles
    axis = torch.eye(3)[dof_axis].repeat(L, 1, 1)
    lq = quat_from_angle_axis(angles.reshape(-1), axis.reshape(-1, 3)).reshape(L, D, 4)
    lq = torch.cat([root_q, lq], dim=1)
    return cal_forward_kinematics(lq, root_t, parents, offsets)


def fk_vjp_analytic(angles, root_t, root_q, g_gq, g_gt, parents, offsets, dof_axis, lower, upper, clip_angles: bool):
    """Vector-Jacobian product of hu_forward_kinematics written out by hand (the spec of the CUDA `fk_vjp_kernel`; checked
    against torch.autograd of hu_forward_kinematics in tests/test_oracle_ik.py).  Upstream gradients g_gq (L,J,4), g_gt (L,J,3)
    -> (g_angles (L,D), g_root_t (L,3), g_root_q (L,4)).  float64 throughout.

    With a_i = R_parent(i) e_i the world axis of hinge i, p_i its pivot, f_k = g_gt[k] and
    tau_k = 1/2 (w g_u + u x g_u - g_w u) for q_k = (u, w), g_gq[k] = (g_u, g_w)   (d q_k = 1/2 [dphi, 0] q_k):
        d L / d theta_i = a_i . sum_{k in subtree(i)} [ tau_k + (p_k - p_i) x f_k ]
    (the straight-through clamp contributes no factor; the normalisation in quat_mul_norm and the sign flip of quat_pos
    map tangent vectors to themselves).  The root quaternion enters un-normalised (kinematics.py:27-35): link 0's rotation
    IS root_q (direct term), every other link turns with root_q's direction (tangential term) and the children of the
    root sit at root_t + |root_q|^2 R offset (radial term)."""
    dd = torch.float64
    L, D = angles.shape[0], len(dof_axis)
    J = D + 1
    th = angles.reshape(L, D).to(dd)
    if clip_angles:
        th = torch.minimum(torch.maximum(th, torch.tensor(lower, dtype=dd)), torch.tensor(upper, dtype=dd))
    q0 = root_q.reshape(L, 4).to(dd)
    t0 = root_t.reshape(L, 3).to(dd)
    off = torch.as_tensor(offsets, dtype=dd)
    g_gq, g_gt = g_gq.to(dd), g_gt.to(dd)
    s2 = (q0 * q0).sum(-1, keepdim=True)
    n0 = q0 / s2.sqrt()
    eye = torch.eye(3, dtype=dd)
    q, p, a = [n0], [t0], [None]
    for j in range(1, J):
        pj = parents[j]
        e = eye[dof_axis[j - 1]].expand(L, 3)
        a.append(quat_rotate(q[pj], e))
        step = quat_rotate(q[pj], off[j].expand(L, 3))
        p.append(p[pj] + (step * s2 if pj == 0 else step))
        half = 0.5 * th[:, j - 1]
        lq = torch.cat([e * half.sin().unsqueeze(-1), half.cos().unsqueeze(-1)], dim=-1)
        qq = quat_mul(q[pj], lq)
        qq = qq / qq.norm(dim=-1, keepdim=True)
        # tau is ODD in q_k: it must be taken at the PUBLISHED representative (quat_pos: w >= 0), the one g_gq refers to
        q.append(torch.where(qq[:, 3:] < 0, -qq, qq))

    def tau(k, qk):
        u, w = qk[:, :3], qk[:, 3:]
        gu, gw = g_gq[:, k, :3], g_gq[:, k, 3:]
        return 0.5 * (w * gu + torch.cross(u, gu, dim=-1) - gw * u)

    S = [torch.zeros(L, 3, dtype=dd)] + [tau(k, q[k]) + torch.cross(p[k], g_gt[:, k], dim=-1) for k in range(1, J)]
    F = [torch.zeros(L, 3, dtype=dd)] + [g_gt[:, k].clone() for k in range(1, J)]
    g_ang = torch.zeros(L, D, dtype=dd)
    g_s = torch.zeros(L, dtype=dd)
    for j in range(J - 1, 0, -1):
        g_ang[:, j - 1] = (a[j] * (S[j] - torch.cross(p[j], F[j], dim=-1))).sum(-1)
        pj = parents[j]
        if pj == 0:
            g_s = g_s + 2.0 * ((p[j] - t0) * F[j]).sum(-1)
        S[pj] = S[pj] + S[j]
        F[pj] = F[pj] + F[j]
    g_phi = S[0] - torch.cross(t0, F[0], dim=-1)
    g_root_t = F[0] + g_gt[:, 0]
    u, w = q0[:, :3], q0[:, 3:]
    tang = torch.cat([w * g_phi + torch.cross(g_phi, u, dim=-1), -(g_phi * u).sum(-1, keepdim=True)], dim=-1)
    g_root_q = g_gq[:, 0] + (2.0 / s2) * tang + (g_s.unsqueeze(-1) / s2) * q0
    return g_ang, g_root_t, g_root_q


# ----------------------------------------------------------------------------------------------
# retarget/utils/parse_mocap.py:106-114 (and :81-89, :126-134; zero_pose_transform.py:33-41)
# ----------------------------------------------------------------------------------------------
def zero_pose_transform(global_q, t2z, axis=(0., 0., 1.)):
    """q' = norm(norm(q * R(pi/2, axis)) * inv(T2Z)).  axis=(1,0,0) is the 'broadcast' variant."""
    rot = quat_from_angle_axis(torch.tensor(torch.pi / 2), torch.tensor(list(axis)))
    out = quat_mul_norm(global_q.clone(), rot)
    return quat_mul_norm(out, quat_inverse(t2z))


# ----------------------------------------------------------------------------------------------
# SciPy Euler split (transform3d.py:52-59 -> scipy Rotation), fp64 numpy
# ----------------------------------------------------------------------------------------------
_AX = {"x": 0, "y": 1, "z": 2}


def euler_from_quat_f64(q: np.ndarray, seq: str) -> np.ndarray:
    """scipy/_rotation_xp.py:365-401 + 1052-1118 restated.  q (...,4) xyzw any float dtype;
    returns (...,3) float64 angles in [-pi, pi], gimbal eps 1e-7 -> third angle 0."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))      # from_quat normalises
    intrinsic = seq.isupper()
    extrinsic = not intrinsic
    axes = [_AX[c] for c in seq.lower()]
    if intrinsic:
        axes = axes[::-1]
    i, j, k = axes
    symmetric = i == k
    if symmetric:
        k = 3 - i - j
    sign = float((i - j) * (j - k) * (k - i) // 2)
    if symmetric:
        a, b, c, d = q[..., 3], q[..., i], q[..., j], q[..., k] * sign
    else:
        a = q[..., 3] - q[..., j]
        b = q[..., i] + q[..., k] * sign
        c = q[..., j] + q[..., 3]
        d = q[..., k] * sign - q[..., i]
    eps = 1e-7
    half_sum = np.arctan2(b, a)
    half_diff = np.arctan2(d, c)
    ang = np.zeros(q.shape[:-1] + (3,), dtype=np.float64)
    ang[..., 1] = 2 * np.arctan2(np.hypot(c, d), np.hypot(a, b))
    first, third = (0, 2) if extrinsic else (2, 0)
    case1 = np.abs(ang[..., 1]) <= eps
    case2 = np.abs(ang[..., 1] - np.pi) <= eps
    case0 = ~(case1 | case2)
    ang[..., 0] = np.where(case1, 2 * half_sum, 2 * half_diff * (-1 if extrinsic else 1))
    ang[..., first] = np.where(case0, half_sum - half_diff, ang[..., first])
    a3 = np.where(case0, half_sum + half_diff, ang[..., third])
    if not symmetric:
        a3 = a3 * sign
        ang[..., 1] = ang[..., 1] - np.pi / 2
    ang[..., third] = a3
    # Wrap as the compiled backend that serves numpy inputs does (scipy _rotation_cy.pyx; the
    # goldens pin it): one conditional +-2*pi, so exactly +-pi is left alone.  (_rotation_xp.py
    # wraps with a modulo instead; the two differ only at exactly +pi.)
    ang = np.where(ang < -np.pi, ang + 2 * np.pi, np.where(ang > np.pi, ang - 2 * np.pi, ang))
    return ang


def quat_in_xyz_axis(q: torch.Tensor, seq: str):
    """transform3d.py:52-59: three single-axis quats (fp64 sin/cos of half angle -> fp32)."""
    ang = euler_from_quat_f64(q.detach().cpu().numpy(), seq)
    outs = []
    for n in range(3):
        e = np.zeros(ang.shape[:-1] + (4,), dtype=np.float64)
        e[..., 3] = np.cos(ang[..., n] / 2.0)
        e[..., _AX[seq[n].lower()]] = np.sin(ang[..., n] / 2.0)
        outs.append(torch.from_numpy(e.astype(np.float32)))
    return tuple(outs)


# ----------------------------------------------------------------------------------------------
# retarget/retarget_solver/body_retargeter.py:34-81  (the quaternion-input path, config 3q)
# ----------------------------------------------------------------------------------------------
def retarget_body_quat(source_global_q, src_parents, num_robot_joints=31, dof_axis=None):
    """Mocap2HuBodyRetargeter.retarget_from_pose, batched.  source_global_q (B,21,4) already
    zero-pose re-referenced.  Returns robot_local_rotation (B,31,4), dof_pos (B,30)."""
    dof_axis = HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B = source_global_q.shape[0]
    sl = cal_local_rotation(source_global_q, src_parents)
    rl = quat_identity((B, num_robot_joints))
    l_sh_pitch, l_sh_roll, l_sh_yaw = quat_in_xyz_axis(sl[:, 18], "YXZ")
    r_sh_pitch, r_sh_roll, r_sh_yaw = quat_in_xyz_axis(sl[:, 14], "YXZ")
    l_el_yaw, l_el_pitch, l_el_roll = quat_in_xyz_axis(sl[:, 19], "ZYX")
    r_el_yaw, r_el_pitch, r_el_roll = quat_in_xyz_axis(sl[:, 15], "ZYX")
    rl[:, 12] = l_sh_pitch
    rl[:, 13] = l_sh_roll
    rl[:, 14] = quat_mul_norm(l_el_yaw, l_sh_yaw)
    rl[:, 21] = r_sh_pitch
    rl[:, 22] = r_sh_roll
    rl[:, 23] = quat_mul_norm(r_el_yaw, r_sh_yaw)
    rl[:, 15] = l_el_pitch
    rl[:, 16] = l_el_roll
    rl[:, 24] = r_el_pitch
    rl[:, 25] = r_el_roll
    dof = quat_to_dof_pos(rl[:, 1:], dof_axis)
    return rl, dof


def retarget_body_quat_test(source_global_q, src_parents, num_robot_joints=31, dof_axis=None):
    """body_retargeter.py:83-99 (retarget_test): direct copy of 4 local quats."""
    dof_axis = HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B = source_global_q.shape[0]
    sl = cal_local_rotation(source_global_q, src_parents)
    rl = quat_identity((B, num_robot_joints))
    rl[:, 13] = sl[:, 18]
    rl[:, 15] = sl[:, 19]
    rl[:, 22] = sl[:, 14]
    rl[:, 24] = sl[:, 15]
    return rl, quat_to_dof_pos(rl[:, 1:], dof_axis)


# ----------------------------------------------------------------------------------------------
# retarget/spatial_transform/transform3d.py position-path ops, batched over a leading B axis
# ----------------------------------------------------------------------------------------------
def coord_transform(p, order=None, dir=None):
    """transform3d.py:24-29."""
    if order is not None:
        p = p[..., order]
    if dir is not None:
        p = p * dir
    return p


def quat_between_two_vecs(v1, v2):
    """transform3d.py:9-21 incl. the whole-batch early-out."""
    if torch.norm(v1, dim=-1).max() <= 1e-6 or torch.norm(v2, dim=-1).max() <= 1e-6:
        return torch.tensor([[0, 0, 0, 1]] * v1.shape[0], dtype=torch.float32)
    v1 = v1 / torch.linalg.norm(v1, dim=-1, keepdim=True)
    v2 = v2 / torch.linalg.norm(v2, dim=-1, keepdim=True)
    cross = torch.cross(v1, v2, dim=-1)
    dots = torch.sum(v1 * v2, dim=-1, keepdim=True)
    return quat_normalize(torch.cat([cross, 1 + dots], dim=-1))


def cal_joint_quat(zero_local_t, motion_local_t):
    """transform3d.py:32-50: Kabsch.  A = M^T Z; U,S,Vt = svd(A); R = U Vt; where det(R) < 0 flip
    the LAST ROW of Vt; quat_from_rotation_matrix(R).  Both args (b,n,3)."""
    A = torch.einsum("bij,bjk->bik", motion_local_t.permute(0, 2, 1), zero_local_t)
    U, _, Vt = torch.linalg.svd(A)
    R = torch.einsum("bij,bjk->bik", U, Vt)
    det = torch.linalg.det(R)
    Vt = Vt.clone()
    Vt[det < 0, -1, :] *= -1
    R = torch.einsum("bij,bjk->bik", U, Vt)
    return quat_from_rotation_matrix(R)


def cal_joint_quat_exact_svd(zero_local_t, motion_local_t):
    """The same Kabsch step with the 3x3 SVD taken in float64 and R rounded once to fp32: what an exact SVD returns.
    NOT the reference's arithmetic (that is MKL's fp32 sgesdd behind torch.linalg.svd, closed source); it is the
    restatement the CUDA kernels' fp64 solve is compared with, and the yardstick for how much of the distance to the
    reference is LAPACK rounding that no other implementation can reproduce (tools/parity_study.py)."""
    A = torch.einsum("bij,bjk->bik", motion_local_t.double().permute(0, 2, 1), zero_local_t.double())
    U, _, Vt = torch.linalg.svd(A)
    R = torch.einsum("bij,bjk->bik", U, Vt)
    Vt = Vt.clone()
    Vt[torch.linalg.det(R) < 0, -1, :] *= -1
    return quat_from_rotation_matrix(torch.einsum("bij,bjk->bik", U, Vt).float())


@contextlib.contextmanager
def exact_kabsch():
    """Within this context every solver below takes its Kabsch rotations from cal_joint_quat_exact_svd."""
    global cal_joint_quat
    saved = cal_joint_quat
    cal_joint_quat = cal_joint_quat_exact_svd
    try:
        yield
    finally:
        cal_joint_quat = saved


def _dot(a, b):
    """torch.dot of two 3-vectors == left-to-right fp32 sum of rounded products (probed)."""
    p = a * b
    return (p[..., 0] + p[..., 1]) + p[..., 2]


def proj_in_plane(v, n):
    """transform3d.py:62-75, batched (v (B,3), n (3,) or (B,3))."""
    n_norm = torch.linalg.norm(n, dim=-1)
    v_proj_n = (_dot(v, n) / n_norm ** 2).unsqueeze(-1) * n
    return v - v_proj_n


def radians_between_vecs(v1, v2, n):
    """transform3d.py:78-100, batched: acos(clamp(v1.v2)) * sign(n.(v1 x v2)) after normalising."""
    v1 = v1 / torch.linalg.norm(v1, dim=-1, keepdim=True)
    v2 = v2 / torch.linalg.norm(v2, dim=-1, keepdim=True)
    normal = n / torch.linalg.norm(n, dim=-1, keepdim=True)
    v1, v2, normal = torch.broadcast_tensors(v1, v2, normal)
    cos_theta = _dot(v1, v2).clamp(-1.0, 1.0)
    angle = torch.acos(cos_theta)
    cross = torch.cross(v1, v2, dim=-1)
    direction = _dot(normal, cross)
    return angle * torch.sign(direction)


_EYE = torch.eye(3, dtype=torch.float32)


def cal_shoulderPR(v1, v0, parent_q):
    """full_body_pos_retargeter.py:247-278 (= retarget_solver.py:127-158 =
    full_body_retargeter.py:210-241).  v1 (B,3) measured bone, v0 (3,) zero-pose bone,
    parent_q (B,4).  Returns pitch quat about y, roll quat about x."""
    v1 = quat_rotate(quat_inverse(parent_q), v1)
    v0 = v0.unsqueeze(0)
    v1p = proj_in_plane(v1, _EYE[1])
    v0p = proj_in_plane(v0, _EYE[1])
    th1 = radians_between_vecs(_EYE[0], v1p, _EYE[1])
    th0 = radians_between_vecs(_EYE[0], v0p, _EYE[1])
    pitch = quat_from_angle_axis(th1 - th0, _EYE[1].expand(v1.shape[0], 3))
    ph1 = radians_between_vecs(v1p, v1, torch.cross(v1p, _EYE[1].expand_as(v1p), dim=-1))
    ph0 = radians_between_vecs(v0p, v0, torch.cross(v0p, _EYE[1].expand_as(v0p), dim=-1))
    roll = quat_from_angle_axis(ph1 - ph0, _EYE[0].expand(v1.shape[0], 3))
    return pitch, roll


def cal_elbowP_and_shoulderY(v1, v0, parent_q):
    """full_body_pos_retargeter.py:221-243 (triplicated likewise): yaw about z, elbow pitch about y."""
    v1 = quat_rotate(quat_inverse(parent_q), v1)
    v0 = v0.unsqueeze(0)
    v1p = proj_in_plane(v1, _EYE[2])
    v0p = proj_in_plane(v0, _EYE[2])
    th1 = radians_between_vecs(_EYE[0], v1p, _EYE[2])
    th0 = radians_between_vecs(_EYE[0], v0p, _EYE[2])
    yaw = quat_from_angle_axis(th1 - th0, _EYE[2].expand(v1.shape[0], 3))
    ph1 = radians_between_vecs(v1p, v1, torch.cross(_EYE[2].expand_as(v1p), v1p, dim=-1))
    ph0 = radians_between_vecs(v0p, v0, torch.cross(_EYE[2].expand_as(v0p), v0p, dim=-1))
    pitch = quat_from_angle_axis(ph1 - ph0, _EYE[1].expand(v1.shape[0], 3))
    return yaw, pitch


def _arm_from_positions(rl, body_t, parent_q_left, parent_q_right, z_lu, z_ll, z_ru, z_rl):
    """Shared arm block of a29/a31/a32 (e.g. full_body_pos_retargeter.py:73-115)."""
    l_pitch, l_roll = cal_shoulderPR(body_t[:, 19] - body_t[:, 18], z_lu, parent_q_left)
    l_elbow_parent = quat_mul(quat_mul(parent_q_left, l_pitch), l_roll)
    l_yaw, l_elbow = cal_elbowP_and_shoulderY(body_t[:, 20] - body_t[:, 19], z_ll, l_elbow_parent)
    rl[:, 12], rl[:, 13], rl[:, 14], rl[:, 15] = l_pitch, l_roll, l_yaw, l_elbow
    r_pitch, r_roll = cal_shoulderPR(body_t[:, 15] - body_t[:, 14], z_ru, parent_q_right)
    r_elbow_parent = quat_mul(quat_mul(parent_q_right, r_pitch), r_roll)
    r_yaw, r_elbow = cal_elbowP_and_shoulderY(body_t[:, 16] - body_t[:, 15], z_rl, r_elbow_parent)
    rl[:, 21], rl[:, 22], rl[:, 23], rl[:, 24] = r_pitch, r_roll, r_yaw, r_elbow
    return rl


def retarget_upper_body(source_global_t, src_offsets, num_robot_joints=31, dof_axis=None):
    """HuUpperBodyFromMocapRetarget.retarget_from_global_translation, retarget_solver.py:40-99.
    source_global_t (B,21,3) in vtrdyn order; src_offsets (21,3) = vtrdyn zero-pose offsets."""
    dof_axis = HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B = source_global_t.shape[0]
    t = coord_transform(source_global_t, dir=torch.tensor([-1., -1., 1.]))
    rl = quat_identity((B, num_robot_joints))
    torso = cal_joint_quat(src_offsets[[17, 13, 11]].unsqueeze(0).expand(B, 3, 3),
                           t[:, [17, 13, 11]] - t[:, [10]])
    rl = _arm_from_positions(rl, t, torso, torso, src_offsets[19], src_offsets[20],
                             src_offsets[15], src_offsets[16])
    return rl, quat_to_dof_pos(rl[:, 1:], dof_axis)


def _wrist_split(rl, base_q, wrist_global_q, arm0):
    """full_body_pos_retargeter.py:128-147: wrist local quat -> intrinsic XYZ -> 3 robot joints."""
    chain = quat_mul(quat_mul(quat_mul(rl[:, arm0], rl[:, arm0 + 1]), rl[:, arm0 + 2]), rl[:, arm0 + 3])
    parent = quat_mul_norm(base_q, chain)
    local = quat_mul_norm(quat_inverse(parent), wrist_global_q)
    a, b, c = quat_in_xyz_axis(local, "XYZ")
    rl[:, arm0 + 4], rl[:, arm0 + 5], rl[:, arm0 + 6] = a, b, c
    return rl


def retarget_full_body_pos(body_t, lhand_t, rhand_t, src_offsets, src_global_t,
                           precise_gripper=True, num_robot_joints=31, dof_axis=None):
    """VtrdynFullBodyPosRetargeter.retarget, full_body_pos_retargeter.py:25-217, batched.
    body_t (B,21,3), l/rhand_t (B,20,3); src_* = vtrdyn_full zero pose (59 joints).
    Returns robot_local_rotation (B,31,4), dof_pos (B,30), body_global_rotation (B,59,4)."""
    dof_axis = HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B = body_t.shape[0]
    rl = quat_identity((B, num_robot_joints))
    bq = quat_identity((B, src_offsets.shape[0]))
    torso = cal_joint_quat(src_offsets[[11, 36, 34]].unsqueeze(0).expand(B, 3, 3),
                           body_t[:, [17, 13, 11]] - body_t[:, [10]])
    rl = _arm_from_positions(rl, body_t, torso, torso, src_offsets[13], src_offsets[14],
                             src_offsets[38], src_offsets[39])
    bq[:, 10] = torso
    lw = cal_joint_quat(src_offsets[[16, 20, 24, 28, 32]].unsqueeze(0).expand(B, 5, 3),
                        lhand_t[:, [2, 6, 10, 14, 17]] - lhand_t[:, [0]])
    rl = _wrist_split(rl, bq[:, 10], lw, 12)
    rw = cal_joint_quat(src_offsets[[41, 45, 49, 53, 56]].unsqueeze(0).expand(B, 5, 3),
                        rhand_t[:, [2, 6, 10, 14, 17]] - rhand_t[:, [0]])
    rl = _wrist_split(rl, bq[:, 10], rw, 21)
    bq[:, 14] = lw
    bq[:, 39] = rw
    dof = quat_to_dof_pos(rl[:, 1:], dof_axis)
    # gripper, :177-217
    orig = (src_global_t[[18, 22, 26, 30, 33], 0] - src_global_t[14, 0]).mean()
    lh = quat_rotate(quat_inverse(lw).unsqueeze(1), lhand_t)
    l_avg = (lh - lh[:, [0]])[:, [4, 8, 12, 16, 19], 0].mean(dim=-1)
    rh = quat_rotate(quat_inverse(rw).unsqueeze(1), rhand_t)
    r_avg = (rh - rh[:, [0]])[:, [4, 8, 12, 16, 19], 0].mean(dim=-1)
    if precise_gripper:
        ls = torch.clip(l_avg / orig - 0.5, 0, 0.5) / 0.5
        rs = torch.clip(r_avg / orig - 0.5, 0, 0.5) / 0.5
        dof[:, 18], dof[:, 19] = ls * 0.044, ls * -0.044
        dof[:, 27], dof[:, 28] = rs * 0.044, rs * -0.044
    else:
        lc = l_avg / orig < 0.7
        rc = r_avg / orig < 0.7
        zero = torch.zeros(B)
        dof[:, 18] = torch.where(lc, zero, zero + 0.044)
        dof[:, 19] = torch.where(lc, zero, zero - 0.044)
        dof[:, 27] = torch.where(rc, zero, zero + 0.044)
        dof[:, 28] = torch.where(rc, zero, zero - 0.044)
    return rl, dof, bq


def retarget_full_body(body_q, body_t, lhand_t, rhand_t, src_offsets, num_robot_joints=31,
                       dof_axis=None):
    """VtrdynFullBodyRetargeter.retarget, full_body_retargeter.py:19-177, batched.  Arms from
    positions with MEASURED parent quats body_q[17]/[13]; wrists from measured body_q[20]/[16];
    binary gripper with finger tips rotated BY the wrist quat (not its inverse)."""
    dof_axis = HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B = body_t.shape[0]
    rl = quat_identity((B, num_robot_joints))
    rl = _arm_from_positions(rl, body_t, body_q[:, 17], body_q[:, 13], src_offsets[13],
                             src_offsets[14], src_offsets[38], src_offsets[39])
    rl = _wrist_split(rl, body_q[:, 17], body_q[:, 20], 12)
    rl = _wrist_split(rl, body_q[:, 13], body_q[:, 16], 21)
    dof = quat_to_dof_pos(rl[:, 1:], dof_axis)
    orig = (src_offsets[[18, 22, 26, 30, 33], 0] - src_offsets[24, 0]).mean()
    lh = quat_rotate(body_q[:, 20].unsqueeze(1), lhand_t)
    l_avg = (lh - lh[:, [0]])[:, [3, 7, 11, 15, 19], 0].mean(dim=-1)
    rh = quat_rotate(body_q[:, 16].unsqueeze(1), rhand_t)
    r_avg = (rh - rh[:, [0]])[:, [3, 7, 11, 15, 19], 0].mean(dim=-1)
    lc = l_avg / orig < 0.7
    rc = r_avg / orig < 0.7
    zero = torch.zeros(B)
    dof[:, 18] = torch.where(lc, zero, zero + 0.044)
    dof[:, 19] = torch.where(lc, zero, zero - 0.044)
    dof[:, 27] = torch.where(rc, zero, zero + 0.044)
    dof[:, 28] = torch.where(rc, zero, zero - 0.044)
    return rl, dof


# ----------------------------------------------------------------------------------------------
# Builder-specified parts: geometric Jacobian and damped-least-squares IK refinement.
# NO reference implementation exists -> PARITY UNPINNED; this code is the spec
# (DESIGN.md section 5).  The linear Jacobian rows are cross-checked against torch.autograd of the
# reference-pinned FK above (tests/test_oracle_ik.py).
# ----------------------------------------------------------------------------------------------
def compute_forward_vector(global_t, left_shoulder, right_shoulder, left_hip, right_hip, gaussian_filter_width=20):
    """poselib/poselib/skeleton/skeleton3d.py:542-566.  (T,J,3) fp32 -> (T,3) fp64.  The gaussian is restated from
    scipy.ndimage (un-vendored, unpinned by the reference; goldens from scipy 1.18.1): weights exp(-x^2/(2 s^2))
    normalised over x = -r..r, r = int(4 s + 0.5); edges repeat the end frames (mode='nearest')."""
    p = np.asarray(global_t, dtype=np.float32)
    side = p[:, left_shoulder] - p[:, right_shoulder] + p[:, left_hip] - p[:, right_hip]
    side = side / np.sqrt((side ** 2).sum(axis=-1))[..., None]
    fwd = np.stack([-side[:, 2].astype(np.float64), np.zeros(len(side)), side[:, 0].astype(np.float64)], axis=-1)
    s = float(gaussian_filter_width)
    r = int(4.0 * s + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (s * s) * x ** 2)
    w = w / w.sum()
    T = fwd.shape[0]
    idx = np.clip(np.arange(T)[:, None] + x[None, :], 0, T - 1)        # (T, 2r+1)
    sm = np.einsum("tkc,k->tc", fwd[idx], w)
    return sm / np.sqrt((sm ** 2).sum(axis=-1))[..., None]


def _axis_quat(theta, k):
    """(.., ) angles about coordinate axis k -> (.., 4); plain (sin, cos) without renormalising."""
    q = torch.zeros(theta.shape + (4,), dtype=theta.dtype)
    q[..., k] = torch.sin(0.5 * theta)
    q[..., 3] = torch.cos(0.5 * theta)
    return q


def geometric_jacobian(angles, root_t, root_q, parents, offsets, dof_axis, lower, upper, clip_angles, links):
    """J (B, K, 6, D): for link k and hinge i (joint i = DOF i-1), a_i = R_parent(i) e_i,
    J_v = a_i x (p_k - p_i), J_w = a_i if i is an ancestor-or-self of k, else 0.  Evaluated at the
    clipped angles (straight-through clamp => no limit term)."""
    B = angles.shape[0]
    D = len(dof_axis)
    gq, gt = hu_forward_kinematics(angles.reshape(B, D, 1), root_t, root_q.reshape(B, 1, 4), parents, offsets,
                                   dof_axis, lower, upper, clip_angles)
    eye = torch.eye(3)
    J = torch.zeros(B, len(links), 6, D)
    for n, k in enumerate(links):
        j = k
        while j > 0:
            p = parents[j]
            a = quat_rotate(gq[:, p], eye[dof_axis[j - 1]].expand(B, 3))
            J[:, n, 0:3, j - 1] = torch.cross(a, gt[:, k] - gt[:, j], dim=-1)
            J[:, n, 3:6, j - 1] = a
            j = p
    return J


ARM_AXIS = [1, 0, 2, 1, 0, 1, 2]          # Hu_DOF_AXIS[11:18] == Hu_DOF_AXIS[20:27]


def ik_refine_arm(theta0, p_sh, offs, lower, upper, pe_t, pw_t, qw_t, iters, damping, rot_weight, active_set=False,
                  residual_history=None):
    """DESIGN.md section 5.  theta0 (B,7) warm start (already clamped); offs (9,3) offsets of the 7 arm
    hinges + 2 gripper links; targets in the robot root frame.  Each step:
      FK of the 7-hinge chain; e = [pe* - p_elbow; pw* - p_wrist; w_o * rotvec(qw* qw^-1)];
      dtheta = (J^T J + lambda^2 I)^-1 J^T e;  theta <- clamp(theta + dtheta, lower, upper).
    Exactly `iters` steps, no early exit.  active_set: a hinge sitting on a limit while its gradient component
    (J^T e)_i pushes further out is frozen for that step (row / column removed from the system).
    The arithmetic runs in theta0's dtype: float32 is the spec the kernels implement, float64 the ground truth both are
    measured against.  residual_history (a list) receives ||e|| before every step."""
    B = theta0.shape[0]
    dt = theta0.dtype                      # float32 = the spec the kernels implement; float64 = ground truth for it
    th = theta0.clone()
    eye = torch.eye(3, dtype=dt)
    lo = torch.tensor(lower, dtype=torch.float32).to(dt)        # the limits are fp32 constants in both
    hi = torch.tensor(upper, dtype=torch.float32).to(dt)
    p_sh, offs, pe_t, pw_t, qw_t = (x.to(dt) for x in (p_sh, offs, pe_t, pw_t, qw_t))
    for _ in range(iters):
        G = quat_identity((B,)).to(dt)
        p = p_sh.expand(B, 3)
        ax, pc = [], []
        for c in range(7):
            ax.append(quat_rotate(G, eye[ARM_AXIS[c]].expand(B, 3)))
            pc.append(p)
            G = quat_normalize(quat_mul(G, _axis_quat(th[:, c], ARM_AXIS[c])))
            if c < 6:
                p = p + quat_rotate(G, offs[c + 1].expand(B, 3))
        qe = quat_normalize(quat_mul(qw_t, quat_conjugate(G)))
        n = qe[:, :3].norm(dim=-1)
        sc = torch.where(n > 1e-8, 2 * torch.atan2(n, qe[:, 3]) / n.clamp(min=1e-30), torch.full_like(n, 2.0))
        e = torch.cat([pe_t - pc[3], pw_t - pc[6], rot_weight * sc.unsqueeze(-1) * qe[:, :3]], dim=-1)   # (B,9)
        if residual_history is not None:
            residual_history.append((e * e).sum(-1).sqrt())
        Jm = torch.zeros(B, 9, 7, dtype=dt)
        for c in range(7):
            if c < 3:
                Jm[:, 0:3, c] = torch.cross(ax[c], pc[3] - pc[c], dim=-1)
            if c < 6:
                Jm[:, 3:6, c] = torch.cross(ax[c], pc[6] - pc[c], dim=-1)
            Jm[:, 6:9, c] = rot_weight * ax[c]
        g = (Jm.transpose(1, 2) @ e.unsqueeze(-1))
        if active_set:
            blocked = ((th >= hi) & (g.squeeze(-1) > 0)) | ((th <= lo) & (g.squeeze(-1) < 0))
            m = (~blocked).to(dt)
            Jm = Jm * m.unsqueeze(1)
            g = g * m.unsqueeze(-1)
        A = Jm.transpose(1, 2) @ Jm + (damping * damping) * torch.eye(7, dtype=dt)
        L = torch.linalg.cholesky(A)
        d = torch.cholesky_solve(g, L).squeeze(-1)
        th = torch.minimum(torch.maximum(th + d, lo), hi)
    return th


def arm_chain(th, p_sh, offs):
    """FK of the 7-hinge arm chain at angles th (B,7): hinge positions pc[0..6] and the final orientation."""
    B = th.shape[0]
    G = quat_identity((B,)).to(th.dtype)
    p = p_sh.to(th.dtype).expand(B, 3)
    offs = offs.to(th.dtype)
    pc = []
    for c in range(7):
        pc.append(p)
        G = quat_normalize(quat_mul(G, _axis_quat(th[:, c], ARM_AXIS[c])))
        if c < 6:
            p = p + quat_rotate(G, offs[c + 1].expand(B, 3))
    return pc, G


def refine_pos_dof(dof, sk, clamp=True, ik_iters=10, damping=0.1, rot_weight=0.2):
    """Limit-aware refinement of the position path (builder-specified; no reference
    counterpart): per arm, the 7 closed-form hinge angles are clamped to the Hu v5 limits and, with ik_iters > 0,
    pulled by damped-least-squares steps towards the elbow / wrist positions and wrist orientation of the
    UNCLAMPED closed-form pose (active-set steps).  dof (B,30) from retarget_full_body_pos; returns the refined copy."""
    T = torch.from_numpy
    rob_par = sk["hu_v5_zero_pose/parents"].tolist()
    rob_off = T(sk["hu_v5_zero_pose/offsets"])
    pos = torch.zeros(31, 3)
    for j in range(1, 31):
        pos[j] = rob_off[j] + pos[rob_par[j]]
    out = dof.clone()
    lo, hi = torch.tensor(HU_V5_DOF_LOWER), torch.tensor(HU_V5_DOF_UPPER)
    for first in (12, 21):
        d0 = first - 1
        th_u = dof[:, d0:d0 + 7]
        th = torch.minimum(torch.maximum(th_u, lo[d0:d0 + 7]), hi[d0:d0 + 7]) if (clamp or ik_iters > 0) else th_u
        if ik_iters > 0:
            pc, G = arm_chain(th_u, pos[first], rob_off[first:first + 9])
            th = ik_refine_arm(th, pos[first], rob_off[first:first + 9], HU_V5_DOF_LOWER[d0:d0 + 7], HU_V5_DOF_UPPER[d0:d0 + 7],
                               pc[3], pc[6], G, ik_iters, damping, rot_weight, active_set=True)
        out[:, d0:d0 + 7] = th
    return out


def body_quat_pipeline(raw_gq, sk, clamp=True, ik_iters=10, damping=0.1, rot_weight=0.2, pre_transformed=False, active_set=False,
                       ik_dtype=torch.float32, residual_history=None):
    """The fused config-3q pipeline: a24 -> a21 -> a30 (+a16, a17) -> [limits -> IK] -> FK.
    Returns robot_local_q (B,31,4), dof (B,30), link_pos (B,31,3).  With clamp=False, ik_iters=0 the
    first two are exactly the reference's retarget_from_pose outputs.  ik_dtype=torch.float64 runs the refinement (targets,
    FK of the chain, Jacobian, Cholesky) in float64 from the SAME fp32 closed-form warm start: the ground truth the fp32
    spec and the kernels are both measured against; dof then comes back as float64."""
    T = torch.from_numpy
    src_par = sk["vtrdyn_zero_pose/parents"].tolist()
    rob_par = sk["hu_v5_zero_pose/parents"].tolist()
    rob_off = T(sk["hu_v5_zero_pose/offsets"])
    zq = raw_gq if pre_transformed else zero_pose_transform(raw_gq, T(sk["t2z/vtrdyn"]))
    rl, dof = retarget_body_quat(zq, src_par)
    B = zq.shape[0]
    do_ik = ik_iters > 0
    if clamp or do_ik:
        lo = torch.tensor(HU_V5_DOF_LOWER)
        hi = torch.tensor(HU_V5_DOF_UPPER)
        dof = dof.clone().to(ik_dtype) if do_ik else dof.clone()
        # zero-pose positions of the robot (identity rotations): p_j = off_j + p_parent
        pos = torch.zeros(31, 3)
        for j in range(1, 31):
            pos[j] = rob_off[j] + pos[rob_par[j]]
        for side, (first, sj) in enumerate([(12, [10, 17, 18, 19, 20]), (21, [10, 13, 14, 15, 16])]):
            d0 = first - 1
            th = torch.minimum(torch.maximum(dof[:, d0:d0 + 7], lo[d0:d0 + 7].to(dof.dtype)), hi[d0:d0 + 7].to(dof.dtype))
            if do_ik:
                zz = zq.to(ik_dtype)
                Tc = quat_conjugate(zz[:, sj[0]])
                Ru = quat_normalize(quat_mul(Tc, zz[:, sj[2]]))
                Rf = quat_normalize(quat_mul(Tc, zz[:, sj[3]]))
                Rh = quat_normalize(quat_mul(Tc, zz[:, sj[4]]))
                pp = pos.to(ik_dtype)
                p_sh = pp[first]
                pe_t = p_sh + quat_rotate(Ru, (pp[first + 3] - pp[first]).expand(B, 3))
                pw_t = pe_t + quat_rotate(Rf, (pp[first + 6] - pp[first + 3]).expand(B, 3))
                th = ik_refine_arm(th, p_sh, rob_off[first:first + 9], HU_V5_DOF_LOWER[d0:d0 + 7],
                                   HU_V5_DOF_UPPER[d0:d0 + 7], pe_t, pw_t, Rh, ik_iters, damping, rot_weight, active_set,
                                   residual_history=residual_history)
            dof[:, d0:d0 + 7] = th
            for c in range(7):
                rl[:, first + c] = _axis_quat(th[:, c].float(), ARM_AXIS[c])
    _, link_pos = cal_forward_kinematics(rl, torch.zeros(B, 3), rob_par, rob_off)
    return rl, dof, link_pos


def ik_residual(dof, zq, sk, rot_weight=0.2):
    """Residual norm of the IK objective at `dof` (for the 'refinement decreases the residual' test)."""
    T = torch.from_numpy
    rob_par = sk["hu_v5_zero_pose/parents"].tolist()
    rob_off = T(sk["hu_v5_zero_pose/offsets"])
    B = dof.shape[0]
    pos = torch.zeros(31, 3)
    for j in range(1, 31):
        pos[j] = rob_off[j] + pos[rob_par[j]]
    ax = torch.eye(3)[HU_V5_DOF_AXIS]
    lq = quat_from_angle_axis(dof.reshape(-1), ax.repeat(B, 1, 1).reshape(-1, 3)).reshape(B, 30, 4)
    lq = torch.cat([quat_identity((B, 1)), lq], dim=1)
    gq, gt = cal_forward_kinematics(lq, torch.zeros(B, 3), rob_par, rob_off)
    tot = torch.zeros(B)
    for first, sj in [(12, [10, 17, 18, 19, 20]), (21, [10, 13, 14, 15, 16])]:
        Tc = quat_conjugate(zq[:, sj[0]])
        Ru = quat_normalize(quat_mul(Tc, zq[:, sj[2]]))
        Rf = quat_normalize(quat_mul(Tc, zq[:, sj[3]]))
        Rh = quat_normalize(quat_mul(Tc, zq[:, sj[4]]))
        pe_t = pos[first] + quat_rotate(Ru, (pos[first + 3] - pos[first]).expand(B, 3))
        pw_t = pe_t + quat_rotate(Rf, (pos[first + 6] - pos[first + 3]).expand(B, 3))
        qe = quat_normalize(quat_mul(Rh, quat_conjugate(gq[:, first + 6])))
        n = qe[:, :3].norm(dim=-1)
        ang = 2 * torch.atan2(n, qe[:, 3])
        tot = tot + ((pe_t - gt[:, first + 3]) ** 2).sum(-1) + ((pw_t - gt[:, first + 6]) ** 2).sum(-1) + (rot_weight * ang) ** 2
    return tot.sqrt()


def synth_clip_3q(L, seed=0, scale=0.5, sk=None):
    """SURVEY.md 8(d) config 3q recipe (same as tools/make_golden.py clip_3q, through the oracle):
    local exp-maps scale*N(0,1) on the vtrdyn T-pose tree -> FK -> raw global quats (L,21,4)."""
    sk = load_skeletons() if sk is None else sk
    g = torch.Generator().manual_seed(seed)
    em = scale * torch.randn(L, 21, 3, generator=g)
    lq = exp_map_to_quat(em)
    gq, _ = cal_forward_kinematics(lq, torch.zeros(L, 3), sk["vtrdyn_t_pose/parents"].tolist(),
                                   torch.from_numpy(sk["vtrdyn_t_pose/offsets"]))
    return gq
