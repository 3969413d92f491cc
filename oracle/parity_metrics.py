"""Distance measures between two retarget results.  TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by
tests/ and tools/parity_study.py to fill profiles/parity_r02.json."""
import numpy as np
import torch

from . import retarget_oracle as oc

ARM_HINGES = list(range(11, 18)) + list(range(20, 27))      # dof indices of the 2 x 7 arm hinges (Hu v5)
GRIPPERS = [18, 19, 27, 28]


def geodesic(q1, q2):
    """Rotation angle between unit quaternions (.., 4), float64, well conditioned near 0 (|vector part| of q1^-1 q2)."""
    a, b = q1.double(), q2.double()
    a = a / a.norm(dim=-1, keepdim=True)
    b = b / b.norm(dim=-1, keepdim=True)
    rel = oc.quat_mul(oc.quat_conjugate(a), b)
    v = rel[..., :3].norm(dim=-1)
    return 2.0 * torch.atan2(v, rel[..., 3].abs())


def fk_positions_of_dof(dof, parents, offsets, dof_axis=None):
    """Link positions of the hinge angles (B, D) on the robot tree, float64 (the comparison is about the two angle sets,
    not about FK rounding).  Gripper DOFs are prismatic on the robot: zeroed."""
    dof_axis = oc.HU_V5_DOF_AXIS if dof_axis is None else dof_axis
    B, D = dof.shape
    d = dof.double().clone()
    d[:, GRIPPERS] = 0
    ax = torch.eye(3, dtype=torch.float64)[dof_axis]
    half = 0.5 * d
    q = torch.cat([ax.unsqueeze(0) * half.sin().unsqueeze(-1), half.cos().unsqueeze(-1)], dim=-1)
    lq = torch.cat([torch.tensor([0., 0, 0, 1], dtype=torch.float64).expand(B, 1, 4), q], dim=1)
    J = D + 1
    off = torch.as_tensor(offsets, dtype=torch.float64)
    gq = [lq[:, 0]]
    gt = [torch.zeros(B, 3, dtype=torch.float64)]
    for j in range(1, J):
        p = int(parents[j])
        gq.append(oc.quat_mul(gq[p], lq[:, j]))
        gt.append(oc.quat_rotate(gq[p], off[j].expand(B, 3)) + gt[p])
    return torch.stack(gt, dim=1)


def distance_stats(dof, dof_ref, rl=None, rl_ref=None, robot_parents=None, robot_offsets=None, tol=1e-5):
    """Per-frame worst |d dof|: fraction within tol, p50 / p99 / max; FK link-position distance of the two angle sets
    (p99 / max over frames of the worst link) and geodesic distance of the local rotations (p99 / max over frames of the
    worst joint).  Non-finite frames are counted and left out."""
    dof, dof_ref = torch.as_tensor(dof).cpu(), torch.as_tensor(dof_ref).cpu()
    fin = torch.isfinite(dof).all(dim=-1) & torch.isfinite(dof_ref).all(dim=-1)
    err = (dof[fin].double() - dof_ref[fin].double()).abs().max(dim=-1).values.numpy()
    q = lambda a, p: float(np.quantile(a, p)) if a.size else float("nan")
    out = {"frames": int(dof.shape[0]), "finite_frames": int(fin.sum()), "frac_le_1e-5": float((err <= tol).mean()),
           "dof_p50": q(err, .5), "dof_p99": q(err, .99), "dof_max": float(err.max())}
    if robot_parents is not None:
        pa = fk_positions_of_dof(dof[fin], robot_parents, robot_offsets)
        pb = fk_positions_of_dof(dof_ref[fin], robot_parents, robot_offsets)
        perr = (pa - pb).norm(dim=-1).amax(dim=-1).numpy()
        out.update({"fk_pos_p99_m": q(perr, .99), "fk_pos_max_m": float(perr.max())})
    if rl is not None:
        geo = geodesic(torch.as_tensor(rl).cpu()[fin], torch.as_tensor(rl_ref).cpu()[fin]).amax(dim=-1).numpy()
        out.update({"geodesic_p99": q(geo, .99), "geodesic_max": float(geo.max())})
    return out
