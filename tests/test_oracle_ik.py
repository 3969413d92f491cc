"""Checks of the builder-specified parts of the oracle (Jacobian, IK): no reference exists for them
(parity unpinned), so they are checked for internal consistency against the reference-pinned FK."""
import numpy as np
import torch

from oracle import retarget_oracle as oc

T = torch.from_numpy


def test_jacobian_linear_rows_match_autograd(skeletons):
    sk = skeletons
    parents = sk["hu_zero_pose/parents"].tolist()
    off = T(sk["hu_zero_pose/offsets"])
    g = torch.Generator().manual_seed(3)
    B = 4
    lo, hi = torch.tensor(oc.HU_DOF_LOWER), torch.tensor(oc.HU_DOF_UPPER)
    ang = lo + (hi - lo) * torch.rand(B, 32, generator=g)
    root_t = torch.randn(B, 3, generator=g)
    root_q = oc.quat_normalize(torch.randn(B, 4, generator=g))
    links = [20, 29, 5]
    J = oc.geometric_jacobian(ang, root_t, root_q, parents, off, oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER, True, links)

    def pos(a):
        _, gt = oc.hu_forward_kinematics(a.reshape(1, 32, 1), root_t[:1], root_q[:1].reshape(1, 1, 4), parents, off,
                                         oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER, False)
        return gt[0, links].reshape(-1)
    Ja = torch.autograd.functional.jacobian(pos, ang[0]).reshape(len(links), 3, 32)
    assert torch.allclose(J[0, :, 0:3, :], Ja, atol=2e-6)


def test_ik_refinement_reduces_residual_and_respects_limits(skeletons):
    sk = skeletons
    raw = oc.synth_clip_3q(256, seed=5, sk=sk)
    zq = oc.zero_pose_transform(raw, T(sk["t2z/vtrdyn"]))
    _, dof0, _ = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=0)
    _, dof1, lp = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=10)
    r0, r1 = oc.ik_residual(dof0, zq, sk), oc.ik_residual(dof1, zq, sk)
    assert float((r1 <= r0 + 1e-5).float().mean()) > 0.99
    assert float(r1.mean()) < 0.7 * float(r0.mean())
    lo, hi = torch.tensor(oc.HU_V5_DOF_LOWER), torch.tensor(oc.HU_V5_DOF_UPPER)
    arm = list(range(11, 18)) + list(range(20, 27))     # only the mapped (arm) hinges are limited:
    # the unmapped leg DOFs stay at the reference's 0, which the knee limit [0.0997, 2.618] excludes
    assert bool(((dof1[:, arm] >= lo[arm]) & (dof1[:, arm] <= hi[arm])).all())
    assert torch.isfinite(lp).all()


def test_zero_pose_is_a_fixed_point(skeletons):
    """T-pose sensors -> zero pose -> all-zero angles; the IK targets coincide with the robot's
    own zero pose, so refinement must leave them at zero."""
    sk = skeletons
    raw = oc.quat_identity((2, 21))
    rl, dof, lp = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=10)
    # the raw identity is the sensor T-pose; after re-referencing the arms are NOT at the zero pose,
    # so use pre-transformed identity for the fixed-point statement
    rl, dof, lp = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=10, pre_transformed=True)
    assert float(dof.abs().max()) < 1e-6
    assert np.allclose(lp[0].numpy(), sk["hu_v5_zero_pose/global_translation"], atol=1e-6)
