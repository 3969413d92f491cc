"""Drop-in for retarget/spatial_transform/transform3d.py (reference lines cited per function).  Like the
reference module it star-re-exports rotation3d and provides `torch` and `np` to `from ... import *` users
(sim_full_body_teleop.py:19,36,85).  Every function is one launch of the sm_100a element-wise family or
of the Kabsch kernel; the per-frame (unbatched) call shapes of the reference and batched ones both work."""
import copy  # noqa: F401  (re-exported by the reference module)
from typing import Dict  # noqa: F401

import numpy as np
import torch

from . import rotation3d as _r3d
from .rotation3d import *  # noqa: F401,F403
from .rotation3d import run_op, euler_seq_code, _engine


def quat_between_two_vecs(vec1, vec2):
    """transform3d.py:9-21: shortest-arc rotation vec1 -> vec2 per row; the WHOLE batch becomes identity when
    either operand's largest row norm is <= 1e-6 (the reference's early-out), reproduced with a device max."""
    v1, v2 = torch.as_tensor(vec1), torch.as_tensor(vec2)
    eng = _engine(v1.device)
    early = eng.max_norm3(v1) <= 1e-6 or eng.max_norm3(v2) <= 1e-6
    if early:
        return torch.tensor([[0, 0, 0, 1]] * v1.shape[0], dtype=torch.float32)
    return run_op(_r3d.OP_QUAT_BETWEEN_TWO_VECS, [v1, v2], [1, 1], [(4,)], iparam=0)[0]


def coord_transform(p, order: list = None, dir=None):
    """transform3d.py:24-29: optional axis permutation, then per-axis multiply."""
    if order is None and dir is None:
        return p
    o = [0, 1, 2] if order is None else [int(i) for i in order]
    d = torch.ones(3) if dir is None else torch.as_tensor(dir, dtype=torch.float32)
    if d.dim() == 0:
        d = d.repeat(3)
    return run_op(_r3d.OP_COORD_TRANSFORM, [p, d], [1, 1], [(3,)], iparam=o[0] | (o[1] << 2) | (o[2] << 4))[0]


def cal_joint_quat(zero_pose_local_translation, motion_local_translation):
    """transform3d.py:32-50 (Kabsch): (b,n,3), (b,n,3) -> (b,4).  A = M^T Z, R = U diag(1,1,det) V^T by an
    fp64 Jacobi eigen-solve on the device, then quat_from_rotation_matrix in the reference's order."""
    z, m = torch.as_tensor(zero_pose_local_translation), torch.as_tensor(motion_local_translation)
    out = _engine(m.device).cal_joint_quat(z, m)
    return out if m.device.type == "cuda" else out.to(m.device)


def quat_in_xyz_axis(q, seq: str = 'xyz'):
    """transform3d.py:52-59: SciPy Euler split (fp64 on the device) into three single-axis quaternions,
    returned as fp32 CPU tensors like `torch.Tensor(ndarray)` does (CUDA input stays on the device)."""
    q = torch.as_tensor(np.asarray(q, dtype=np.float32)) if not torch.is_tensor(q) else q
    q1, q2, q3 = run_op(_r3d.OP_EULER_SPLIT, [q], [1], [(4,), (4,), (4,)], iparam=euler_seq_code(seq))
    return q1, q2, q3


def proj_in_plane(v, n):
    """transform3d.py:62-75 (the reference handles one vector; rows broadcast here)."""
    n_t = torch.as_tensor(n)
    if n_t.dim() == 1:
        assert _engine(n_t.device).max_norm3(n_t.reshape(1, 3)) > 1e-6
    return run_op(_r3d.OP_PROJ_IN_PLANE, [v, n], [1, 1], [(3,)])[0]


def radians_between_vecs(v1, v2, n):
    """transform3d.py:78-100"""
    return run_op(_r3d.OP_RADIANS_BETWEEN_VECS, [v1, v2, n], [1, 1, 1], [()])[0]


def exp_map_to_quat(exp_map):
    """transform3d.py:146-150"""
    return _r3d.exp_map_to_quat(exp_map)


def quat_slerp(q0, q1, t):
    """transform3d.py:153-174: t carries a trailing singleton dimension, as in the reference."""
    return run_op(_r3d.OP_QUAT_SLERP, [q0, q1, t], [1, 1, 1], [(4,)])[0]


def quat_to_dof_pos(quat, dof_axis):
    """transform3d.py:177-183: quat (D,4) [or (B,D,4)], dof_axis: List[int] of length D -> (D,) [(B,D)]."""
    ax = torch.as_tensor(np.asarray(dof_axis, dtype=np.float32)).reshape(-1, 1)
    return run_op(_r3d.OP_QUAT_TO_DOF_POS, [quat, ax], [1, 1], [()])[0]
