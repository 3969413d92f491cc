"""Drop-in for poselib/poselib/core/rotation3d.py (reference lines cited per function): the same free
functions, same names and argument meaning, each backed by one launch of the element-wise sm_100a kernel
family behind `hrt_rot_op` (csrc/hrt_ops.cuh).  CPU tensors in -> CPU tensors out like the reference;
CUDA tensors stay on the device.  Leading (batch) dimensions broadcast as in torch.  fp32 throughout.

There is no CPU arithmetic here: slicing / concatenation / allocation is torch plumbing, every number
is produced by the CUDA library (a missing library or GPU raises HrtError)."""
import math
from typing import List, Optional

import numpy as np
import torch

from . import _lib

# op codes: include/hrt_b200.h (HRT_OP_*)
(OP_QUAT_MUL, OP_QUAT_MUL_NORM, OP_QUAT_MUL_THREE, OP_QUAT_MUL_FOUR, OP_QUAT_POS, OP_QUAT_ABS, OP_QUAT_UNIT,
 OP_QUAT_NORMALIZE, OP_QUAT_CONJUGATE, OP_QUAT_ROTATE, OP_QUAT_FROM_ANGLE_AXIS, OP_QUAT_FROM_ROTATION_MATRIX,
 OP_QUAT_ANGLE_AXIS, OP_QUAT_YAW_ROTATION, OP_TRANSFORM_INVERSE, OP_TRANSFORM_MUL, OP_TRANSFORM_APPLY,
 OP_ROT_MATRIX_DET, OP_ROT_MATRIX_FROM_QUATERNION, OP_PROJECT_QUAT_TO_AXIS, OP_EXTRACT_ROTATION_ALONG_AXIS,
 OP_NORMALIZE_ANGLE, OP_QUAT_TO_ANGLE_AXIS, OP_QUAT_TO_EXP_MAP, OP_EXP_MAP_TO_ANGLE_AXIS, OP_EXP_MAP_TO_QUAT,
 OP_ANGLE_AXIS_TO_EXP_MAP, OP_QUAT_BETWEEN_TWO_VECS, OP_PROJ_IN_PLANE, OP_RADIANS_BETWEEN_VECS, OP_QUAT_SLERP,
 OP_QUAT_TO_DOF_POS, OP_EULER_SPLIT, OP_EULER_ANGLES_F64, OP_COORD_TRANSFORM, OP_CAL_SHOULDER_PR,
 OP_CAL_ELBOWP_SHOULDERY) = range(37)

_AX = {"x": 0, "y": 1, "z": 2}


def _engine(device=None):
    from .engine import default_engine
    idx = 0 if device is None or device.type != "cuda" or device.index is None else device.index
    return default_engine(idx)


def euler_seq_code(seq: str) -> int:
    """Sequence code of HRT_OP_EULER_SPLIT: axes as written in bits 0-5, bit 6 = extrinsic (lower case)."""
    if len(seq) != 3 or not (seq.islower() or seq.isupper()) or any(c not in "xyz" for c in seq.lower()):
        raise ValueError(f"Expected axis specification to be a non-empty string of upto 3 characters, got {seq}")
    a = [_AX[c] for c in seq.lower()]
    if a[0] == a[1] or a[1] == a[2]:
        raise ValueError(f"Expected consecutive axes to be different, got {seq}")
    return a[0] | (a[1] << 2) | (a[2] << 4) | (64 if seq.islower() else 0)


def run_op(op, operands, row_dims, out_row_shapes, iparam=0, fparam=0.0, out_dtype=torch.float32):
    """operands[k]: tensor whose last row_dims[k] dimensions form one row; the leading dimensions
    broadcast.  Returns one tensor per entry of out_row_shapes, shaped broadcast_batch + row shape."""
    first = next(t for t in operands if torch.is_tensor(t))
    ret_dev = first.device
    eng = _engine(ret_dev)
    ts, batches = [], []
    for t, rd in zip(operands, row_dims):
        t = torch.as_tensor(t)
        batches.append(tuple(t.shape[:t.dim() - rd]))
        ts.append(t)
    common = tuple(torch.broadcast_shapes(*batches))
    n = int(np.prod(common)) if len(common) else 1
    ins, periods = [], []
    for t, b in zip(ts, batches):
        rows = int(np.prod(b)) if len(b) else 1
        stripped = tuple(b)
        while stripped and stripped[0] == 1:
            stripped = stripped[1:]
        if tuple(b) == common or rows == n:
            per = 0
        elif rows == 1:
            per = 1
        elif len(stripped) <= len(common) and common[len(common) - len(stripped):] == stripped:
            per = rows                                   # a trailing-suffix table, e.g. (J,4) against (B,J,4)
        else:
            t = t.expand(common + tuple(t.shape[len(b):]))
            per = 0
        ins.append(t.to(device=eng.device, dtype=torch.float32).contiguous())
        periods.append(per)
    outs = []
    for rs in out_row_shapes:
        width = int(np.prod(rs)) if len(rs) else 1
        words = width * (2 if out_dtype == torch.float64 else 1)
        outs.append(torch.empty((max(n, 1), words), device=eng.device, dtype=torch.float32))
    if n > 0:
        eng.rot_op(op, n, ins, periods, outs, iparam, fparam)
    res = []
    for o, rs in zip(outs, out_row_shapes):
        if out_dtype == torch.float64:
            o = o.view(torch.float64)
        o = o[:n].reshape(common + tuple(rs))
        res.append(o if ret_dev.type == "cuda" else o.to(ret_dev))
    return res


# ------------------------------------------------------------------------------------------ quaternions
def quat_mul(a, b):
    """rotation3d.py:15-27"""
    return run_op(OP_QUAT_MUL, [a, b], [1, 1], [(4,)])[0]


def quat_pos(x):
    """rotation3d.py:31-38"""
    return run_op(OP_QUAT_POS, [x], [1], [(4,)])[0]


def quat_abs(x):
    """rotation3d.py:42-47"""
    return run_op(OP_QUAT_ABS, [x], [1], [()])[0]


def quat_unit(x):
    """rotation3d.py:51-56"""
    return run_op(OP_QUAT_UNIT, [x], [1], [(4,)])[0]


def quat_conjugate(x):
    """rotation3d.py:60-64"""
    return run_op(OP_QUAT_CONJUGATE, [x], [1], [(4,)])[0]


def quat_real(x):
    """rotation3d.py:68-72 (a view)"""
    return x[..., 3]


def quat_imaginary(x):
    """rotation3d.py:76-80 (a view)"""
    return x[..., :3]


def quat_norm_check(x):
    """rotation3d.py:84-89"""
    n = quat_abs(x)
    assert bool(((n - 1).abs() < 1e-3).all()), "the quaternion is has non-1 norm: {}".format((n - 1).abs())
    assert bool((x[..., 3] >= 0).all()), "the quaternion has negative real part"


def quat_normalize(q):
    """rotation3d.py:93-98"""
    return run_op(OP_QUAT_NORMALIZE, [q], [1], [(4,)])[0]


def quat_from_xyz(xyz):
    """rotation3d.py:102-108.  The reference takes the Frobenius norm of the WHOLE tensor (`xyz.norm()`):
    a degenerate constructor no caller uses; kept for name compatibility, w = 1 - ||xyz||_F."""
    flat = xyz.reshape(-1)
    pad = (-flat.numel()) % 4
    rows = torch.cat([flat, flat.new_zeros(pad)]).reshape(-1, 4)
    parts = quat_abs(rows)                                   # per-row norms on the device ...
    total = quat_abs(torch.cat([parts, parts.new_zeros((-parts.numel()) % 4)]).reshape(-1, 4)) if parts.numel() > 1 else parts
    while total.numel() > 1:                                  # ... folded until one value is left
        total = quat_abs(torch.cat([total, total.new_zeros((-total.numel()) % 4)]).reshape(-1, 4))
    w = (1.0 - total.reshape(())).unsqueeze(-1)
    assert bool((w >= 0).all()), "xyz has its norm greater than 1"
    return torch.cat([xyz, w.to(xyz.device)], dim=-1)


def quat_identity(shape: List[int]):
    """rotation3d.py:112-119: [0,0,0,1] rows on the CPU, like the reference."""
    q = torch.zeros(list(shape) + [4])
    q[..., 3] = 1.0
    return q


def quat_from_angle_axis(angle, axis, degree: bool = False):
    """rotation3d.py:123-143"""
    return run_op(OP_QUAT_FROM_ANGLE_AXIS, [angle, axis], [0, 1], [(4,)], iparam=int(bool(degree)))[0]


def quat_from_rotation_matrix(m):
    """rotation3d.py:147-193"""
    return run_op(OP_QUAT_FROM_ROTATION_MATRIX, [m], [2], [(4,)])[0]


def quat_mul_norm(x, y):
    """rotation3d.py:197-202"""
    return run_op(OP_QUAT_MUL_NORM, [x, y], [1, 1], [(4,)])[0]


def quat_rotate(rot, vec):
    """rotation3d.py:206-211"""
    return run_op(OP_QUAT_ROTATE, [rot, vec], [1, 1], [(3,)])[0]


def quat_inverse(x):
    """rotation3d.py:215-219"""
    return quat_conjugate(x)


def quat_identity_like(x):
    """rotation3d.py:223-227"""
    return quat_identity(list(x.shape[:-1]))


def quat_angle_axis(x):
    """rotation3d.py:231-240.  Like the reference, the normalised axis is ALSO written back into
    x[..., :3] (the reference divides a view of its argument in place)."""
    angle, axis = run_op(OP_QUAT_ANGLE_AXIS, [x], [1], [(), (3,)])
    x[..., :3] = axis.to(x.device)
    return angle, x[..., :3]


def quat_yaw_rotation(x, z_up: bool = True):
    """rotation3d.py:244-261"""
    return run_op(OP_QUAT_YAW_ROTATION, [x], [1], [(4,)], iparam=int(bool(z_up)))[0]


# ------------------------------------------------------------------------------------------ transforms
def transform_from_rotation_translation(r: Optional[torch.Tensor] = None, t: Optional[torch.Tensor] = None):
    """rotation3d.py:265-275"""
    assert r is not None or t is not None, "rotation and translation can't be all None"
    if r is None:
        r = quat_identity(list(t.shape)).to(t.device)
    if t is None:
        t = torch.zeros(list(r.shape) + [3], device=r.device)
    return torch.cat([r, t], dim=-1)


def transform_identity(shape: List[int]):
    """rotation3d.py:279-285"""
    return transform_from_rotation_translation(quat_identity(shape), torch.zeros(list(shape) + [3]))


def transform_rotation(x):
    """rotation3d.py:289-291"""
    return x[..., :4]


def transform_translation(x):
    """rotation3d.py:295-297"""
    return x[..., 4:]


def transform_inverse(x):
    """rotation3d.py:301-306"""
    return run_op(OP_TRANSFORM_INVERSE, [x], [1], [(7,)])[0]


def transform_identity_like(x):
    """rotation3d.py:310-314"""
    return transform_identity(list(x.shape))


def transform_mul(x, y):
    """rotation3d.py:318-326"""
    return run_op(OP_TRANSFORM_MUL, [x, y], [1, 1], [(7,)])[0]


def transform_apply(rot, vec):
    """rotation3d.py:330-335"""
    assert isinstance(vec, torch.Tensor)
    return run_op(OP_TRANSFORM_APPLY, [rot, vec], [1, 1], [(3,)])[0]


# ------------------------------------------------------------------------------------------ matrices
def rot_matrix_det(x):
    """rotation3d.py:339-350"""
    return run_op(OP_ROT_MATRIX_DET, [x], [2], [()])[0]


def rot_matrix_integrity_check(x):
    """rotation3d.py:354-365.  (The reference body calls Tensor methods that do not exist and cannot run;
    this is the check it describes: det = 1 and R R^T = I to 1e-3, via the device kernels.)"""
    det = rot_matrix_det(x)
    assert bool(((det - 1).abs() < 1e-3).all()), "the matrix has non-one determinant"
    q = quat_from_rotation_matrix(x)
    back = rot_matrix_from_quaternion(q)
    assert bool(((back - x.to(back.device)).abs() < 1e-3).all()), "the matrix is not orthogonal"


def rot_matrix_from_quaternion(quaternions: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:399-427"""
    return run_op(OP_ROT_MATRIX_FROM_QUATERNION, [quaternions], [1], [(3, 3)])[0]


def euclidean_to_rotation_matrix(x):
    """rotation3d.py:431-435"""
    return x[..., :3, :3]


def euclidean_integrity_check(x):
    """rotation3d.py:439-442"""
    euclidean_to_rotation_matrix(x)
    assert bool((x[..., 3, :3] == 0).all()), "the last row is illegal"
    assert bool((x[..., 3, 3] == 1).all()), "the last row is illegal"


def euclidean_translation(x):
    """rotation3d.py:446-450"""
    return x[..., :3, 3]


def euclidean_to_transform(transformation_matrix):
    """rotation3d.py:466-473"""
    return transform_from_rotation_translation(
        r=quat_from_rotation_matrix(euclidean_to_rotation_matrix(transformation_matrix)),
        t=euclidean_translation(transformation_matrix))


def euclidean_inverse(x):
    """rotation3d.py:454-462.  (The reference indexes column 4 of a 4x4 matrix and cannot run; this returns
    the inverse rigid transform it describes.)"""
    inv = transform_inverse(euclidean_to_transform(x))
    s = torch.zeros_like(x)
    s[..., :3, :3] = rot_matrix_from_quaternion(inv[..., :4]).to(x.device)
    s[..., :3, 3] = inv[..., 4:].to(x.device)
    s[..., 3, 3] = 1.0
    return s


# ------------------------------------------------------------------------------------------ axis projections
def _project(batch_q, which):
    return run_op(OP_PROJECT_QUAT_TO_AXIS, [batch_q], [1], [(4,)], iparam=which)[0]


def project_quat_to_axis_x(batch_q: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:480-486"""
    return _project(batch_q, 0)


def project_quat_to_axis_y(batch_q: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:489-495"""
    return _project(batch_q, 1)


def project_quat_to_axis_z(batch_q: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:498-504"""
    return _project(batch_q, 2)


def project_quat_to_axis_xy(batch_q: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:507-517"""
    return _project(batch_q, 3)


def project_quat_to_axis_xz(batch_q: torch.Tensor) -> torch.Tensor:
    """rotation3d.py:520-530"""
    return _project(batch_q, 4)


def extract_rotation_along_axis(batch_quat: torch.Tensor, axis: int) -> torch.Tensor:
    """rotation3d.py:535-556"""
    if axis not in (0, 1, 2):
        raise ValueError("Invalid axis. Axis must be 0 (x), 1 (y), or 2 (z).")
    return run_op(OP_EXTRACT_ROTATION_ALONG_AXIS, [batch_quat], [1], [()], iparam=axis)[0]


def quat_mul_four(q1, q2, q3, q4):
    """rotation3d.py:560-567"""
    return run_op(OP_QUAT_MUL_FOUR, [q1, q2, q3, q4], [1, 1, 1, 1], [(4,)])[0]


def quat_mul_three(q1, q2, q3):
    """rotation3d.py:571-577"""
    return run_op(OP_QUAT_MUL_THREE, [q1, q2, q3], [1, 1, 1], [(4,)])[0]


# ------------------------------------------------------------------------------------------ exp-map family
def normalize_angle(x):
    """rotation3d.py:583-584"""
    return run_op(OP_NORMALIZE_ANGLE, [x], [0], [()])[0]


def quat_to_angle_axis(q):
    """rotation3d.py:588-608"""
    angle, axis = run_op(OP_QUAT_TO_ANGLE_AXIS, [q], [1], [(), (3,)])
    return angle, axis


def angle_axis_to_exp_map(angle, axis):
    """rotation3d.py:612-617"""
    return run_op(OP_ANGLE_AXIS_TO_EXP_MAP, [angle, axis], [0, 1], [(3,)])[0]


def quat_to_exp_map(q):
    """rotation3d.py:621-627"""
    return run_op(OP_QUAT_TO_EXP_MAP, [q], [1], [(3,)])[0]


def exp_map_to_angle_axis(exp_map):
    """rotation3d.py:630-646"""
    angle, axis = run_op(OP_EXP_MAP_TO_ANGLE_AXIS, [exp_map], [1], [(), (3,)])
    return angle, axis


def exp_map_to_quat(exp_map):
    """rotation3d.py:649-652"""
    return run_op(OP_EXP_MAP_TO_QUAT, [exp_map], [1], [(4,)])[0]


def quat_to_eular(q):
    """rotation3d.py:659-661: scipy Rotation.from_quat(q).as_euler('xyz', degrees=True) -> float64 ndarray."""
    q = torch.as_tensor(np.asarray(q, dtype=np.float32)) if not torch.is_tensor(q) else q
    out = run_op(OP_EULER_ANGLES_F64, [q], [1], [(3,)], iparam=euler_seq_code("xyz") | 0x80, out_dtype=torch.float64)[0]
    return out.cpu().numpy()


__all__ = [n for n in dir() if n.startswith(("quat_", "transform_", "rot_matrix_", "euclidean_", "project_quat_", "exp_map_"))] + [
    "extract_rotation_along_axis", "normalize_angle", "angle_axis_to_exp_map", "math", "torch", "List", "Optional"]
