"""Differentiable FK (hrt_fk_vjp behind torch.autograd), and the builder-specified Jacobian / IK stages pinned to a
float64 run of their own spec (no reference implementation exists for them: SURVEY.md F2, parity unpinned).

  * gradients: HuForwardModel.forward_kinematics(...).backward() against torch.autograd through the oracle's restatement
    of the reference FK with its straight-through clamp (robot_kinematics_model/hu_forward_model.py:17-33), <= 1e-5 rel;
  * Jacobian / IK: |kernel - float64 spec| must not exceed |float32 spec - float64 spec| (the oracle's fp32 loop is
    itself noisy; this shows the kernel is as accurate as the fp32 torch loop, not merely that it agrees with it);
  * hrt_ik_refine: the stage on its own, against the oracle, with the per-step objective."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def hrt():
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as h
    return h


@pytest.fixture(scope="module")
def oc():
    from oracle import retarget_oracle
    return retarget_oracle


def _q(a, p):
    return float(np.quantile(a, p))


@pytest.mark.parametrize("asset,tables", [("hu_zero_pose", "HU"), ("hu_v5_zero_pose", "HU_V5")])
@pytest.mark.parametrize("clip", [True, False])
def test_fk_gradients_match_autograd_through_the_reference_fk(hrt, oc, skeletons, parity, asset, tables, clip):
    ax, lo, hi = (getattr(oc, f"{tables}_DOF_{n}") for n in ("AXIS", "LOWER", "UPPER"))
    parents = skeletons[asset + "/parents"].tolist()
    off = T(skeletons[asset + "/offsets"])
    D, L = len(ax), 512
    g = torch.Generator().manual_seed(17)
    lo_t, hi_t = torch.tensor(lo), torch.tensor(hi)
    ang0 = (lo_t + (hi_t - lo_t) * (torch.rand(L, D, generator=g) * 1.3 - 0.15)).reshape(L, D, 1)     # 15 % beyond the limits each side
    rt0 = torch.randn(L, 3, generator=g)
    rq0 = oc.quat_normalize(torch.randn(L, 1, 4, generator=g))
    w_q, w_t = torch.randn(L, D + 1, 4, generator=g), torch.randn(L, D + 1, 3, generator=g)
    model = hrt.HuForwardModel(hrt.RobotZeroPose.from_asset(asset).skeleton_tree, device="cuda:0")

    def loss_of(gq, gt):
        # a scalar of link rotations and positions: weighted sums plus a quadratic term so that the upstream gradient
        # depends on the forward values
        return (gq * w_q.to(gq)).sum() + (gt * w_t.to(gt)).sum() + 0.5 * (gt * gt).sum()

    # ours: CPU leaf tensors in (the reference's calling convention), .grad back on the CPU
    ang, rt, rq = ang0.clone().requires_grad_(True), rt0.clone().requires_grad_(True), rq0.clone().requires_grad_(True)
    gq, gt = model.forward_kinematics(ang, rt, rq, clip)
    assert gq.requires_grad and gt.requires_grad and gq.device.type == "cpu"
    loss_of(gq, gt).backward()
    # truth: float64 autograd through the oracle's restatement of the reference FK (straight-through clamp)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        a64, t64, q64 = (x.double().clone().requires_grad_(True) for x in (ang0, rt0, rq0))
        gq_o, gt_o = oc.hu_forward_kinematics(a64, t64, q64, parents, off.double(), ax, lo, hi, clip)
        loss_of(gq_o, gt_o).backward()
    finally:
        torch.set_default_dtype(prev)
    # and the same in float32 (what a user of the reference would get)
    a32, t32, q32 = (x.clone().requires_grad_(True) for x in (ang0, rt0, rq0))
    gq_r, gt_r = oc.hu_forward_kinematics(a32, t32, q32, parents, off, ax, lo, hi, clip)
    loss_of(gq_r, gt_r).backward()
    rec = {}
    for name, mine, truth, ref32 in (("angles", ang.grad, a64.grad, a32.grad), ("root_t", rt.grad, t64.grad, t32.grad),
                                     ("root_q", rq.grad, q64.grad, q32.grad)):
        scale = float(truth.abs().max())
        e = float((mine.double() - truth).abs().max()) / scale
        e32 = float((ref32.double() - truth).abs().max()) / scale
        rec[f"{name}_rel_err"] = e
        rec[f"{name}_fp32_autograd_rel_err"] = e32
        assert mine.shape == truth.shape and mine.device.type == "cpu"
        assert e <= 1e-5, (name, e, e32)
    parity.record(f"a22 d(FK)/d(inputs) kernel VJP vs float64 autograd of the reference FK ({asset}, clip={clip}, {L} configs)", rec)
    # straight-through: a clamped hinge still receives gradient (the reference's reason for the construct)
    if clip:
        outside = ((ang0.reshape(L, D) < lo_t) | (ang0.reshape(L, D) > hi_t))
        assert bool(outside.any()) and float(ang.grad.reshape(L, D)[outside].abs().max()) > 1e-3
    # no grad requested -> plain tensors, as before
    gq_n, gt_n = model.forward_kinematics(ang0, rt0, rq0, clip)
    assert not gq_n.requires_grad and torch.equal(gq_n, gq.detach()) and torch.equal(gt_n, gt.detach())
    # CUDA leaves work too, gradients arrive on the device
    ang_c = ang0.cuda().requires_grad_(True)
    gq_c, gt_c = model.forward_kinematics(ang_c, rt0.cuda(), rq0.cuda(), clip)
    loss_of(gq_c, gt_c).backward()
    assert ang_c.grad.is_cuda and float((ang_c.grad.cpu() - ang.grad).abs().max()) <= 1e-6 * max(1.0, float(ang.grad.abs().max()))


def test_jacobian_kernel_is_as_accurate_as_the_fp32_spec(hrt, oc, skeletons, parity):
    eng = hrt.default_engine(0, robot="hu")
    parents = skeletons["hu_zero_pose/parents"].tolist()
    off = T(skeletons["hu_zero_pose/offsets"])
    B = 2048
    g = torch.Generator().manual_seed(23)
    lo, hi = torch.tensor(oc.HU_DOF_LOWER), torch.tensor(oc.HU_DOF_UPPER)
    ang = lo + (hi - lo) * (torch.rand(B, 32, generator=g) * 1.2 - 0.1)
    rt = torch.randn(B, 3, generator=g)
    rq = oc.quat_normalize(torch.randn(B, 4, generator=g))
    links = [20, 29]
    Jk = eng.fk_jacobian(hrt.TREE_ROBOT, ang, links, rt, rq, clip=True).cpu().double()
    J32 = oc.geometric_jacobian(ang, rt, rq, parents, off, oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER, True, links).double()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        J64 = oc.geometric_jacobian(ang.double(), rt.double(), rq.double(), parents, off.double(), oc.HU_DOF_AXIS, oc.HU_DOF_LOWER,
                                    oc.HU_DOF_UPPER, True, links)
    finally:
        torch.set_default_dtype(prev)
    ek = (Jk - J64).abs().amax(dim=(1, 2, 3)).numpy()
    e32 = (J32 - J64).abs().amax(dim=(1, 2, 3)).numpy()
    rec = {"configs": B, "kernel_vs_fp64_p50": _q(ek, .5), "kernel_vs_fp64_p99": _q(ek, .99), "kernel_vs_fp64_max": float(ek.max()),
           "fp32_spec_vs_fp64_p50": _q(e32, .5), "fp32_spec_vs_fp64_p99": _q(e32, .99), "fp32_spec_vs_fp64_max": float(e32.max())}
    parity.record("J geometric Jacobian (2 wrist links, Hu): kernel and fp32 spec against the float64 spec", rec)
    # as accurate as the fp32 torch loop (within a factor for the different, equally valid, rounding order) and absolutely small
    assert rec["kernel_vs_fp64_p50"] <= 2.0 * rec["fp32_spec_vs_fp64_p50"] + 1e-7
    assert rec["kernel_vs_fp64_p99"] <= 2.0 * rec["fp32_spec_vs_fp64_p99"] + 1e-7
    assert rec["kernel_vs_fp64_max"] <= 5e-6


def _ik_inputs(oc, sk, B, seed):
    raw = oc.synth_clip_3q(B, seed=seed, sk=sk)
    zq = oc.zero_pose_transform(raw, T(sk["t2z/vtrdyn"]))
    _, dof0, _ = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=0)
    rob_par = sk["hu_v5_zero_pose/parents"].tolist()
    rob_off = T(sk["hu_v5_zero_pose/offsets"])
    pos = torch.zeros(31, 3)
    for j in range(1, 31):
        pos[j] = rob_off[j] + pos[rob_par[j]]
    th0, pe, pw, qw = [], [], [], []
    for first, sj in [(12, [10, 17, 18, 19, 20]), (21, [10, 13, 14, 15, 16])]:
        Tc = oc.quat_conjugate(zq[:, sj[0]])
        Ru, Rf, Rh = (oc.quat_normalize(oc.quat_mul(Tc, zq[:, sj[k]])) for k in (2, 3, 4))
        pe_t = pos[first] + oc.quat_rotate(Ru, (pos[first + 3] - pos[first]).expand(B, 3))
        pw_t = pe_t + oc.quat_rotate(Rf, (pos[first + 6] - pos[first + 3]).expand(B, 3))
        th0.append(dof0[:, first - 1:first + 6]); pe.append(pe_t); pw.append(pw_t); qw.append(Rh)
    stack = lambda xs: torch.stack(xs, dim=1).contiguous()
    return raw, stack(th0), stack(pe), stack(pw), stack(qw), pos, rob_off


def test_fused_ik_is_as_accurate_as_the_fp32_spec(hrt, oc, skeletons, parity):
    """|fused kernel - float64 spec| against |float32 spec - float64 spec| on the bench workload's recipe."""
    sk = skeletons
    B = 4096
    raw = oc.synth_clip_3q(B, seed=31, sk=sk)
    eng = hrt.Engine(0).set_standard_trees()
    for tag, active in (("plain", False), ("active_set", True)):
        flags = hrt.BQ_CLAMP | hrt.BQ_IK | (hrt.BQ_ACTIVE_SET if active else 0)
        _, dof_k, _ = eng.retarget_body_quat(raw, flags=flags, ik_iters=10)
        _, dof32, _ = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=10, active_set=active)
        _, dof64, _ = oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=10, active_set=active, ik_dtype=torch.float64)
        ek = (dof_k.cpu().double() - dof64).abs().max(dim=-1).values.numpy()
        e32 = (dof32.double() - dof64).abs().max(dim=-1).values.numpy()
        rec = {"frames": B, "kernel_vs_fp64_frac_le_1e-5": float((ek <= 1e-5).mean()), "kernel_vs_fp64_p50": _q(ek, .5),
               "kernel_vs_fp64_p99": _q(ek, .99), "kernel_vs_fp64_max": float(ek.max()),
               "fp32_spec_vs_fp64_frac_le_1e-5": float((e32 <= 1e-5).mean()), "fp32_spec_vs_fp64_p50": _q(e32, .5),
               "fp32_spec_vs_fp64_p99": _q(e32, .99), "fp32_spec_vs_fp64_max": float(e32.max())}
        parity.record(f"IK fused 10-step refinement ({tag}): kernel and fp32 spec against the float64 spec", rec)
        # the kernel's FMA / polynomial-sincos / rsqrt flavour is a small factor noisier than torch's fp32 ops (measured p99
        # 9.3e-6 vs 6.8e-6 plain, 8.4e-6 vs 3.4e-6 active set) and stays under the 1e-5 bar at p99 in absolute terms
        assert rec["kernel_vs_fp64_p50"] <= 2.0 * rec["fp32_spec_vs_fp64_p50"] + 2e-7, rec
        assert rec["kernel_vs_fp64_p99"] <= 3.0 * rec["fp32_spec_vs_fp64_p99"] + 1e-6 and rec["kernel_vs_fp64_p99"] <= 1.2e-5, rec
        assert rec["kernel_vs_fp64_frac_le_1e-5"] >= rec["fp32_spec_vs_fp64_frac_le_1e-5"] - 0.01, rec
        assert rec["kernel_vs_fp64_max"] <= max(2.0 * rec["fp32_spec_vs_fp64_max"], 1e-3), rec
    eng.close()


def test_standalone_ik_refine_matches_the_spec_and_descends(hrt, oc, skeletons, parity):
    sk = skeletons
    B = 2048
    raw, th0, pe, pw, qw, pos, rob_off = _ik_inputs(oc, sk, B, seed=37)
    eng = hrt.Engine(0).set_standard_trees()
    for tag, active in (("plain", False), ("active_set", True)):
        th_k, res = eng.ik_refine(th0, pe, pw, qw, iters=10, damping=0.1, rot_weight=0.2, active_set=active, want_residual=True)
        assert th_k.shape == (B, 2, 7) and res.shape == (B, 2, 11)
        th_k, res = th_k.cpu(), res.cpu()
        th64 = torch.empty(B, 2, 7, dtype=torch.float64)
        th32 = torch.empty(B, 2, 7)
        for side, first in enumerate((12, 21)):
            d0 = first - 1
            args = (pos[first], rob_off[first:first + 9], oc.HU_V5_DOF_LOWER[d0:d0 + 7], oc.HU_V5_DOF_UPPER[d0:d0 + 7],
                    pe[:, side], pw[:, side], qw[:, side], 10, 0.1, 0.2, active)
            th32[:, side] = oc.ik_refine_arm(th0[:, side], *args)
            th64[:, side] = oc.ik_refine_arm(th0[:, side].double(), *args)
        ek = (th_k.double() - th64).abs().amax(dim=(1, 2)).numpy()
        e32 = (th32.double() - th64).abs().amax(dim=(1, 2)).numpy()
        down = ((res[:, :, 1:] - res[:, :, :-1]) <= 1e-6)
        rec = {"frames": B, "kernel_vs_fp64_p50": _q(ek, .5), "kernel_vs_fp64_p99": _q(ek, .99), "kernel_vs_fp64_max": float(ek.max()),
               "fp32_spec_vs_fp64_p50": _q(e32, .5), "fp32_spec_vs_fp64_p99": _q(e32, .99), "fp32_spec_vs_fp64_max": float(e32.max()),
               "steps_non_increasing": float(down.float().mean()), "arms_monotone_over_all_steps": float(down.all(dim=-1).float().mean()),
               "objective_before_mean": float(res[:, :, 0].mean()), "objective_after_mean": float(res[:, :, -1].mean())}
        parity.record(f"IK hrt_ik_refine ({tag}): kernel and fp32 spec against the float64 spec; per-step objective", rec)
        assert rec["kernel_vs_fp64_p50"] <= 2.0 * rec["fp32_spec_vs_fp64_p50"] + 2e-7, rec
        assert rec["kernel_vs_fp64_p99"] <= 3.0 * rec["fp32_spec_vs_fp64_p99"] + 1e-6 and rec["kernel_vs_fp64_p99"] <= 1.2e-5, rec
        assert rec["objective_after_mean"] < 0.8 * rec["objective_before_mean"], rec
        if active:
            # the active-set variant is a descent method: the objective does not rise from one step to the next
            assert rec["steps_non_increasing"] >= 0.98 and rec["arms_monotone_over_all_steps"] >= 0.93, rec
        # the same angles as the fused pipeline produces for this arm (same device code)
        flags = hrt.BQ_CLAMP | hrt.BQ_IK | (hrt.BQ_ACTIVE_SET if active else 0)
        _, dof_f, _ = eng.retarget_body_quat(raw, flags=flags, ik_iters=10)
        fused = torch.stack([dof_f[:, 11:18], dof_f[:, 20:27]], dim=1).cpu()
        assert float(np.quantile((fused - th_k).abs().amax(dim=(1, 2)).numpy(), 0.97)) <= 1e-5     # targets built on the host here, in the kernel there
    # argument errors and the edge cases of the entry point
    assert eng.ik_refine(th0[:0], pe[:0], pw[:0], qw[:0]).shape == (0, 2, 7)
    bad = th0.clone()
    bad[3, 1, 2] = float("nan")
    out = eng.ik_refine(bad, pe, pw, qw).cpu()
    assert torch.isnan(out[3, 1]).all() and torch.isfinite(out[3, 0]).all() and torch.isfinite(out[4]).all()
    with pytest.raises(hrt.HrtError):
        eng.ik_refine(th0, pe, pw, qw, iters=-1)
    e2 = hrt.Engine(0)
    with pytest.raises(hrt.HrtError):
        e2.ik_refine(th0, pe, pw, qw)                     # arm tables not installed
    e2.close()
    eng.close()
