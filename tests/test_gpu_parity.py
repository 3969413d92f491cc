"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against
  (1) the golden fixtures produced by the UNMODIFIED reference (tests/golden, tools/make_golden.py),
  (2) the CPU oracle on the same seeded inputs at sizes the oracle finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes (2^20 frames / 65,536 configs).

Tolerances (BASELINE.json north_star): joint angles <= 1e-5 rad, FK link positions <= 1e-5 m,
tables / pure fp32 arithmetic chains bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T = torch.from_numpy
ANGLE_TOL = 1e-5
POS_TOL = 1e-5


@pytest.fixture(scope="module")
def hrt():
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as h
    return h


@pytest.fixture(scope="module")
def eng(hrt):
    return hrt.default_engine(0)


@pytest.fixture(scope="module")
def eng_hu(hrt):
    return hrt.default_engine(0, robot="hu")


@pytest.fixture(scope="module")
def oc():
    from oracle import retarget_oracle
    return retarget_oracle


@pytest.fixture(scope="module")
def pm():
    from oracle import parity_metrics
    return parity_metrics


# Measured floors (profiles/parity_study_cpu_r02.json): the oracle with an EXACT (float64) SVD in the Kabsch step and every
# other op bit-identical to the reference already sits this far from the reference's goldens, because the reference's
# rotation comes out of MKL's fp32 sgesdd: a32 0.89 of frames within 1e-5 rad, p99 5.1e-5, max 9.1e-5, FK max 1.8e-5 m;
# a29 0.95 / 5.9e-5 / 1.8e-4 / 8.5e-5 m.  Gates against the REFERENCE are those floors plus a margin; gates against the
# oracle with the same exact SVD (the kernel's own arithmetic class) are the north-star bars on every frame.
# STAIR: below ~7e-4 rad the reference's angle read-back 2*acos(w) (rotation3d.py:588-608) can only return 0, 4.9e-4 or 6.9e-4
# (cos(theta/2) rounds to 1 or 1 - 2^-24): two implementations that differ by one ulp in w land on different steps.  No
# "all frames" bound on a raw angle can be tighter than one step; FK of such a frame moves by step x arm length.
STAIR = 7.5e-4


def check_dist(st, frac=None, p99=None, dmax=None, fk_max=None, geo_max=None, fk_p99=None):
    if fk_p99 is not None:
        assert st["fk_pos_p99_m"] <= fk_p99, st
    if frac is not None:
        assert st["frac_le_1e-5"] >= frac, st
    if p99 is not None:
        assert st["dof_p99"] <= p99, st
    if dmax is not None:
        assert st["dof_max"] <= dmax, st
    if fk_max is not None:
        assert st["fk_pos_max_m"] <= fk_max, st
    if geo_max is not None:
        assert st["geodesic_max"] <= geo_max, st


def maxdiff(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


# ------------------------------------------------------------------------------- FK (a20, a22)
def test_fk_angles_vs_reference_golden(hrt, eng_hu, golden):
    g = golden("fk_hu")
    for clip, key in [(True, "clip"), (False, "noclip")]:
        for exact in (False, True):
            gq, gt = eng_hu.fk_angles(hrt.TREE_ROBOT, T(g["angles"]).reshape(256, 32), T(g["root_t"]),
                                      T(g["root_q"]).reshape(256, 4), clip=clip, exact=exact)
            assert maxdiff(gt, g[f"gt_{key}"]) <= POS_TOL
            assert maxdiff(gq, g[f"gq_{key}"]) <= 3e-6
    # zero angles -> zero-pose link positions (SURVEY 7.1 step 3 known answer)
    gq, gt = eng_hu.fk_angles(hrt.TREE_ROBOT, torch.zeros(3, 32), None, None, clip=False)
    from humanoid_real_time_retarget_b200 import robot_config as cfg
    sk = cfg.skeleton_tables()
    zero = np.zeros((33, 3), np.float32)
    for j in range(1, 33):
        zero[j] = sk["hu_zero_pose/offsets"][j] + zero[sk["hu_zero_pose/parents"][j]]
    assert maxdiff(gt[0], zero) <= 1e-6


def test_fk_clip_of_wild_angles_matches_the_reference_arithmetic(hrt, eng_hu, oc, skeletons):
    """hu_forward_model.py:27-33: the forward value of the clamp is (clamp(x) - x) + x, which LOSES the clamp to rounding
    for huge inputs (|x| > ~1e7 rad: the result is a multiple of ulp(x), not the limit).  The kernel's fast path for
    trees with finite limits must notice such angles (warp vote) and still agree with the reference arithmetic."""
    g = torch.Generator().manual_seed(9)
    B = 64
    ang = (torch.rand(B, 32, generator=g) - 0.5) * 2.0
    wild = torch.tensor([1.0e7, -3.0e7, 5.0e7, 1.3e8, -7.7e8, 3.0e4, -2.0e5, 1.0e30])
    for i, w in enumerate(wild):
        ang[i * 3 % B, (7 * i) % 32] = w
        ang[(i * 5 + 1) % B, (11 * i + 3) % 32] = -w
    root_t = torch.zeros(B, 3)
    root_q = torch.zeros(B, 1, 4)
    root_q[..., 3] = 1.0
    gq_o, gt_o = oc.hu_forward_kinematics(ang.reshape(B, 32, 1), root_t, root_q, skeletons["hu_zero_pose/parents"].tolist(),
                                          T(skeletons["hu_zero_pose/offsets"]), oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER, True)
    for exact in (False, True):
        gq, gt = eng_hu.fk_angles(hrt.TREE_ROBOT, ang, root_t, root_q.reshape(B, 4), clip=True, exact=exact)
        assert maxdiff(gt, gt_o) <= POS_TOL and maxdiff(gq, gq_o) <= 3e-6, exact


@pytest.mark.parametrize("name,tree", [("hu_v5_zero_pose", 0), ("vtrdyn_t_pose", 1), ("vtrdyn_full_zero_pose", 2)])
def test_fk_local_quats_and_local_rotation_vs_golden(hrt, eng, golden, name, tree):
    g = golden(f"fk_{name}")
    if name == "vtrdyn_t_pose":       # T-pose offsets differ from the zero-pose tree installed in slot 1
        from humanoid_real_time_retarget_b200 import robot_config as cfg
        sk = cfg.skeleton_tables()
        e = hrt.Engine(0)
        e.set_tree(3, sk["vtrdyn_t_pose/parents"], sk["vtrdyn_t_pose/offsets"])
        tree, engine = 3, e
    else:
        engine = eng
    gq, gt = engine.fk_local_quats(tree, T(g["local_q"]), T(g["root_t"]), exact=True)
    assert np.array_equal(gq.cpu().numpy(), g["gq"]), "exact-mode FK rotations must be bit-exact"
    assert np.array_equal(gt.cpu().numpy(), g["gt"]), "exact-mode FK positions must be bit-exact"
    gq, gt = engine.fk_local_quats(tree, T(g["local_q"]), T(g["root_t"]), exact=False)
    assert maxdiff(gq, g["gq"]) <= 3e-6 and maxdiff(gt, g["gt"]) <= POS_TOL
    lq = engine.local_from_global(tree, T(g["gq"]))
    assert np.array_equal(lq.cpu().numpy(), g["local_back"]), "cal_local_rotation must be bit-exact"
    # poselib SkeletonState FK agrees (tree.quat == identity)
    assert maxdiff(gq, g["sk_global_rotation"]) <= 3e-6 and maxdiff(gt, g["sk_global_translation"]) <= POS_TOL


def test_drop_in_functional_api(hrt, golden, skeletons):
    g = golden("fk_hu_v5_zero_pose")
    parents = skeletons["hu_v5_zero_pose/parents"].tolist()
    off = T(skeletons["hu_v5_zero_pose/offsets"])
    # CPU tensors in -> CPU tensors out, like the reference
    gq, gt = hrt.cal_forward_kinematics(T(g["local_q"]), T(g["root_t"]), parents, off)
    assert gq.device.type == "cpu" and maxdiff(gt, g["gt"]) <= POS_TOL
    lq = hrt.cal_local_rotation(T(g["gq"]), parents)
    assert np.array_equal(lq.numpy(), g["local_back"])
    # HuForwardModel on the 33-joint Hu tree
    gh = golden("fk_hu")

    class Tree:
        local_translation = T(skeletons["hu_zero_pose/offsets"])
        parent_indices = T(skeletons["hu_zero_pose/parents"]).long()
        num_joints = 33
    model = hrt.HuForwardModel(Tree())
    gq, gt = model.forward_kinematics(T(gh["angles"]), T(gh["root_t"]), T(gh["root_q"]), True)
    assert maxdiff(gt, gh["gt_clip"]) <= POS_TOL and gq.shape == (256, 33, 4)


# ------------------------------------------------------------------------------- a24
def test_zero_pose_transform_bit_exact(hrt, eng, golden):
    g = golden("zero_pose_transform")
    assert np.array_equal(eng.zero_pose_transform(hrt.TREE_SOURCE, T(g["q21"])).cpu().numpy(), g["vtrdyn"])
    assert np.array_equal(eng.zero_pose_transform(hrt.TREE_SOURCE_FULL, T(g["q59"])).cpu().numpy(), g["vtrdyn_full"])
    assert np.array_equal(eng.zero_pose_transform(hrt.TREE_SOURCE, T(g["q21"]), 1).cpu().numpy(), g["vtrdyn_broadcast"])
    out = hrt.vtrdyn_zero_pose_transform(T(g["q21"]))
    assert out.device.type == "cpu" and np.array_equal(out.numpy(), g["vtrdyn"])


# ------------------------------------------------------------------------------- fused quaternion path
def test_body_quat_vs_reference_golden(hrt, eng, golden):
    g = golden("body_quat")
    lq, dof, lp = eng.retarget_body_quat(T(g["raw_global_q"]), flags=0)
    assert maxdiff(lq, g["robot_local_q"]) <= 1.2e-7
    assert maxdiff(dof, g["dof_pos"]) <= ANGLE_TOL
    assert maxdiff(lp, g["fk_gt"]) <= POS_TOL
    # pre-transformed input + the reference-named class, one frame at a time like the reference
    lq2, dof2, _ = eng.retarget_body_quat(T(g["zero_pose_q"]), flags=hrt.BQ_PRE_TRANSFORMED)
    assert torch.equal(lq2, lq) and torch.equal(dof2, dof)
    src = hrt.RobotZeroPose.from_asset("vtrdyn_zero_pose")
    tgt = hrt.RobotZeroPose.from_asset("hu_v5_zero_pose")
    solver = hrt.Mocap2HuBodyRetargeter(src, tgt)
    for i in range(8):
        rl_i, dof_i = solver.retarget_from_pose(T(g["zero_pose_q"][i]))
        assert rl_i.shape == (31, 4) and dof_i.shape == (30,) and rl_i.device.type == "cpu"
        assert maxdiff(dof_i, g["dof_pos"][i]) <= ANGLE_TOL
    assert solver.motion_dof_pos.shape == (8, 30)
    assert maxdiff(solver.motion_global_translation, g["fk_gt"][:8]) <= POS_TOL


def test_body_quat_vs_oracle_all_frames(hrt, eng, oc, skeletons):
    B = 50_000
    raw = oc.synth_clip_3q(B, seed=21, sk=skeletons)
    lq, dof, lp = eng.retarget_body_quat(raw, flags=0)
    rl_o, dof_o, lp_o = oc.body_quat_pipeline(raw, skeletons, clamp=False, ik_iters=0)
    err = (dof.cpu() - dof_o).abs().max(dim=-1).values
    assert float(err.max()) <= ANGLE_TOL, f"{int((err > ANGLE_TOL).sum())} of {B} frames exceed {ANGLE_TOL}"
    assert maxdiff(lp, lp_o) <= POS_TOL
    assert maxdiff(lq, rl_o) <= 1.2e-7


def test_body_quat_with_clamp_and_ik_vs_oracle(hrt, eng, oc, skeletons):
    B = 8192
    raw = oc.synth_clip_3q(B, seed=22, sk=skeletons)
    lq, dof, lp = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP)
    rl_o, dof_o, lp_o = oc.body_quat_pipeline(raw, skeletons, clamp=True, ik_iters=0)
    assert maxdiff(dof, dof_o) <= ANGLE_TOL and maxdiff(lp, lp_o) <= POS_TOL
    lq, dof, lp = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK, ik_iters=10, damping=0.1, rot_weight=0.2)
    rl_o, dof_o, lp_o = oc.body_quat_pipeline(raw, skeletons, clamp=True, ik_iters=10, damping=0.1, rot_weight=0.2)
    err = (dof.cpu() - dof_o).abs().max(dim=-1).values
    # Builder-specified stage (parity unpinned by the reference): kernel vs own oracle, conditioning-aware.
    # Ten Gauss-Newton steps amplify rounding on frames whose targets are far out of reach: the oracle moved
    # against ITSELF by 1-ulp input jitter shows p99.9 = 7e-5 rad, max 3e-4 (printed below), so no
    # implementation can promise 1e-5 on every frame.  Acceptance: >= 97 % of frames within 1e-5,
    # p99.9 within 1e-4 (the oracle's own noise floor), every frame within 1e-3.
    gj = torch.Generator().manual_seed(1)
    raw_j = torch.nextafter(raw, raw + torch.sign(torch.randn(raw.shape, generator=gj)))
    _, dof_j, _ = oc.body_quat_pipeline(raw_j, skeletons, clamp=True, ik_iters=10, damping=0.1, rot_weight=0.2)
    self_delta = (dof_j - dof_o).abs().max(dim=-1).values
    within = float((err <= ANGLE_TOL).float().mean())
    print(f"IK parity: {within:.4f} of frames within {ANGLE_TOL}; p99.9 {float(np.quantile(err.numpy(), 0.999)):.2e} "
          f"max {float(err.max()):.2e}; oracle self-delta: {float((self_delta <= ANGLE_TOL).float().mean()):.4f} within, "
          f"p99.9 {float(np.quantile(self_delta.numpy(), 0.999)):.2e} max {float(self_delta.max()):.2e}")
    assert within >= 0.97
    assert float(np.quantile(err.numpy(), 0.999)) <= 1e-4
    assert float(err.max()) <= 1e-3
    assert float(np.quantile((lp.cpu() - lp_o).abs().amax(dim=(1, 2)).numpy(), 0.99)) <= POS_TOL
    zq = oc.zero_pose_transform(raw, T(skeletons["t2z/vtrdyn"]))
    _, dof0, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP)
    r0 = oc.ik_residual(dof0.cpu(), zq, skeletons)
    r1 = oc.ik_residual(dof.cpu(), zq, skeletons)
    assert float(r1.mean()) < 0.7 * float(r0.mean())


def test_body_quat_active_set_refinement(hrt, eng, oc, skeletons):
    """HRT_BQ_ACTIVE_SET: same refinement with blocked hinges frozen; kernel vs oracle, and it must be a descent
    method (the IK objective never increases, and ends lower than without the active set)."""
    B = 8192
    raw = oc.synth_clip_3q(B, seed=23, sk=skeletons)
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    _, dof_a, lp_a = eng.retarget_body_quat(raw, flags=flags | hrt.BQ_ACTIVE_SET, ik_iters=10)
    _, dof_o, lp_o = oc.body_quat_pipeline(raw, skeletons, clamp=True, ik_iters=10, active_set=True)
    err = (dof_a.cpu() - dof_o).abs().max(dim=-1).values
    print(f"active-set IK parity: {float((err <= ANGLE_TOL).float().mean()):.4f} within 1e-5; p99.9 {float(np.quantile(err.numpy(), 0.999)):.2e} max {float(err.max()):.2e}")
    assert float((err <= ANGLE_TOL).float().mean()) >= 0.95 and float(np.quantile(err.numpy(), 0.99)) <= 1e-4
    zq = oc.zero_pose_transform(raw, T(skeletons["t2z/vtrdyn"]))
    _, dof_c, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP)
    _, dof_p, _ = eng.retarget_body_quat(raw, flags=flags, ik_iters=10)
    r_c, r_p, r_a = (oc.ik_residual(d.cpu(), zq, skeletons) for d in (dof_c, dof_p, dof_a))
    assert float((r_a > r_c + 1e-5).float().mean()) <= 0.001
    assert float(r_a.mean()) < float(r_p.mean()) < float(r_c.mean())


def test_packed_ik_variant_matches_scalar(hrt, eng, oc, skeletons):
    """The FFMA2 (two arms per thread) refinement is the same algorithm with a different rounding sequence:
    it must agree with the default kernel like the default kernel agrees with the oracle, on ragged sizes too."""
    for B in (1, 16, 17, 33, 4097):
        raw = oc.synth_clip_3q(B, seed=40 + B, sk=skeletons)
        flags = hrt.BQ_CLAMP | hrt.BQ_IK
        _, d0, p0 = eng.retarget_body_quat(raw, flags=flags, want_local_q=False)
        _, d1, p1 = eng.retarget_body_quat(raw, flags=flags | hrt.BQ_PACKED_IK, want_local_q=False)
        err = (d1 - d0).abs().max(dim=-1).values.cpu().numpy()
        perr = (p1 - p0).abs().amax(dim=(1, 2)).cpu().numpy()
        # the default kernel carries the chain as a rotation matrix, the packed one as a quaternion: two rounding
        # sequences of one algorithm, which agree like each agrees with the fp32 spec (a handful of sensitive frames
        # differ by more; on the short clips the 97th percentile IS such a frame, so they are gated on the median)
        # (the plain clamped iteration is not a contraction on every frame: the fp32 spec itself is up to 3.6e-3 from the
        # float64 run of the same spec on its worst frame of 4096, profiles/parity_r02.json)
        assert float(err.max()) <= 5e-3
        if B >= 1000:
            assert float(np.quantile(err, 0.97)) <= ANGLE_TOL and float(np.quantile(perr, 0.97)) <= POS_TOL
            assert float(np.quantile(err, 0.999)) <= 1e-4
        else:
            assert float(np.median(err)) <= ANGLE_TOL and float(np.median(perr)) <= POS_TOL
        rest = [i for i in range(30) if i not in list(range(11, 18)) + list(range(20, 27))]
        assert float(d1[:, rest].abs().max()) == 0.0


def test_body_quat_cta_shapes_agree(hrt, eng, oc, skeletons):
    """Calls without the refinement run the instantiation with the IK loop compiled out, at 28 warps per CTA or 8 (short
    clips); with the refinement 16 / 8.  Same device code for the closed form, so the same bits whichever shape runs and
    whichever outputs are requested (the local rotations go straight to HBM, the rest through staged images)."""
    B = 40_000
    raw = oc.synth_clip_3q(B, seed=21, sk=skeletons).cuda()
    lq12, dof12, lp12 = eng.retarget_body_quat(raw, flags=0)                                     # all three outputs
    _, dof28, lp28 = eng.retarget_body_quat(raw, flags=0, want_local_q=False)                   # 28 warps
    _, dof8, lp8 = eng.retarget_body_quat(raw[:4000], flags=0, want_local_q=False)              # 8 warps (short clip)
    assert torch.equal(dof28, dof12) and torch.equal(lp28, lp12)
    assert torch.equal(dof8, dof28[:4000]) and torch.equal(lp8, lp28[:4000])
    _, dofc, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP, want_local_q=False)            # clamp only: no-IK instantiation too
    _, dofi, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK, ik_iters=0, want_local_q=False)   # IK instantiation, 0 steps
    assert (dofc - dofi).abs().max().item() <= 1e-6
    # and against the oracle on a sample
    idx = torch.arange(0, B, 97)
    _, dof_o, lp_o = oc.body_quat_pipeline(raw[idx].cpu(), skeletons, clamp=False, ik_iters=0)
    assert maxdiff(dof28[idx], dof_o) <= 2e-6 and maxdiff(lp28[idx], lp_o) <= POS_TOL


def test_ragged_clip_lengths_around_the_cta_shape_thresholds(hrt, eng, oc, skeletons):
    """Clip lengths just below / above the point where the host switches to the large CTAs (one small CTA per SM = 148 * 8
    warps * 16 frames = 18,944 frames), none a multiple of the 16-frame warp group: every length must reproduce the
    frames of the longest run bit for bit, and nothing may be written past the end of the outputs."""
    n_max = 18_944 + 16 * 28 * 3 + 5
    raw = oc.synth_clip_3q(n_max, seed=33, sk=skeletons).cuda()
    _, dof_ref, lp_ref = eng.retarget_body_quat(raw, flags=0, want_local_q=False)
    g = torch.Generator(device="cuda").manual_seed(8)
    em = 0.4 * torch.randn(n_max, 59, 3, device="cuda", generator=g)
    root = torch.zeros(n_max, 3, device="cuda")
    root[:, 2] = 1.0
    _, gt = eng.fk_local_quats(hrt.TREE_SOURCE_FULL, hrt.rotation3d.exp_map_to_quat(em), root, exact=True)
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
    pdof_ref = torch.empty(n_max, 30, device="cuda")
    eng.retarget_full_body_pos(body, lh, rh, out=(None, pdof_ref, None))
    for n in (18_943, 18_944, 18_945, 18_961, 18_944 + 16 * 28 + 7, n_max - 1):
        guard = 4
        dof = torch.full((n + guard, 30), 7.0, device="cuda")
        lp = torch.full((n + guard, 31, 3), 7.0, device="cuda")
        eng.retarget_body_quat(raw[:n], flags=0, out=(None, dof[:n], lp[:n]))
        assert torch.equal(dof[:n], dof_ref[:n]) and torch.equal(lp[:n], lp_ref[:n]), n
        assert bool((dof[n:] == 7.0).all()) and bool((lp[n:] == 7.0).all()), n
        pdof = torch.full((n + guard, 30), 7.0, device="cuda")
        eng.retarget_full_body_pos(body[:n], lh[:n], rh[:n], out=(None, pdof[:n], None))
        same = (pdof[:n] == pdof_ref[:n]) | (torch.isnan(pdof[:n]) & torch.isnan(pdof_ref[:n]))
        assert bool(same.all()) and bool((pdof[n:] == 7.0).all()), n


def test_body_quat_edge_cases(hrt, eng, oc, skeletons):
    # empty, single, ragged (not a multiple of the 16-frame warp group)
    for B in (0, 1, 15, 17, 33):
        raw = oc.synth_clip_3q(max(B, 1), seed=B, sk=skeletons)[:B]
        lq, dof, lp = eng.retarget_body_quat(raw, flags=0)
        assert lq.shape == (B, 31, 4) and dof.shape == (B, 30) and lp.shape == (B, 31, 3)
        if B:
            _, dof_o, lp_o = oc.body_quat_pipeline(raw, skeletons, clamp=False, ik_iters=0)
            assert maxdiff(dof, dof_o) <= ANGLE_TOL and maxdiff(lp, lp_o) <= POS_TOL
    # T-pose sensors (all identity) and the pre-transformed zero pose -> all-zero angles (inv. 1)
    ident = oc.quat_identity((4, 21))
    _, dof, lp = eng.retarget_body_quat(ident, flags=hrt.BQ_PRE_TRANSFORMED | hrt.BQ_CLAMP | hrt.BQ_IK)
    assert float(dof.abs().max()) <= 1e-6
    assert maxdiff(lp[0], skeletons["hu_v5_zero_pose/global_translation"]) <= 1e-6
    # NaN in -> NaN out for the affected arm only, never a crash
    raw = oc.synth_clip_3q(32, seed=3, sk=skeletons)
    raw[5, 18] = float("nan")
    _, dof, _ = eng.retarget_body_quat(raw, flags=0)
    assert torch.isfinite(dof[:5]).all() and torch.isfinite(dof[6:]).all()
    # The reference's own angle read-back masks NaN to angle 0 (|sin| > 1e-5 is False for NaN, rotation3d.py:601-605), so the
    # closed form reports 0 for the poisoned arm.  The refinement must not turn NaN targets into clamped garbage: NaN out.
    _, dof_c, _ = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP, want_local_q=False)
    assert torch.isfinite(dof_c).all() and float(dof_c[5, 11:16].abs().max()) == 0.0
    for fl in (hrt.BQ_CLAMP | hrt.BQ_IK, hrt.BQ_CLAMP | hrt.BQ_IK | hrt.BQ_PACKED_IK, hrt.BQ_CLAMP | hrt.BQ_IK | hrt.BQ_ACTIVE_SET):
        _, dof_l, lp_l = eng.retarget_body_quat(raw, flags=fl, want_local_q=False)
        assert torch.isfinite(dof_l[:5]).all() and torch.isfinite(dof_l[6:]).all() and torch.isnan(dof_l[5, 11:18]).all()
        assert torch.isfinite(dof_l[5, 20:27]).all() and torch.isnan(lp_l[5, 13:19]).any()
    # quaternion double cover: -q gives the same answer
    raw = oc.synth_clip_3q(64, seed=4, sk=skeletons)
    _, d1, _ = eng.retarget_body_quat(raw, flags=0)
    _, d2, _ = eng.retarget_body_quat(-raw, flags=0)
    assert maxdiff(d1, d2) <= 1e-6


def test_error_behaviour(hrt, eng):
    import ctypes as C
    e = hrt.Engine(0)
    buf = torch.zeros(64, device="cuda")
    p = C.c_void_p(buf.data_ptr())
    assert e.lib.hrt_fk_local_quats(e._h, 0, 1, p, None, p, None, 0, None) == -2      # tree not installed
    assert b"not installed" in e.lib.hrt_last_error_string()
    assert e.lib.hrt_retarget_body_quat(e._h, 1, p, 0, 0, 0.1, 0.2, None, None, None, None) == -2
    with pytest.raises(hrt.HrtError):
        e.set_tree(0, [0, -1], np.zeros((2, 3)))          # root must come first
    with pytest.raises(hrt.HrtError):
        e.set_tree(0, [-1] + [0] * 70, np.zeros((71, 3)))  # too many joints
    # misaligned device pointer is refused, not silently mis-read
    buf = torch.zeros(21 * 4 * 4 + 1, device="cuda")
    import ctypes as C
    rc = eng.lib.hrt_retarget_body_quat(eng._h, 4, C.c_void_p(buf.data_ptr() + 4), 0, 0, 0.1, 0.2, None, None, None, None)
    assert rc == -4


# ------------------------------------------------------------------------------- Jacobian
def test_jacobian_vs_oracle(hrt, eng_hu, oc, skeletons):
    g = torch.Generator().manual_seed(9)
    B = 200
    lo, hi = torch.tensor(oc.HU_DOF_LOWER), torch.tensor(oc.HU_DOF_UPPER)
    ang = lo + (hi - lo) * (torch.rand(B, 32, generator=g) * 1.2 - 0.1)
    root_t = torch.randn(B, 3, generator=g)
    root_q = oc.quat_normalize(torch.randn(B, 4, generator=g))
    links = [20, 29, 5, 32]
    J = eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, links, root_t, root_q, clip=True)
    Jo = oc.geometric_jacobian(ang, root_t, root_q, skeletons["hu_zero_pose/parents"].tolist(),
                               T(skeletons["hu_zero_pose/offsets"]), oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER,
                               True, links)
    assert J.shape == (B, 4, 6, 32)
    assert maxdiff(J, Jo) <= 5e-6
    # chains longer than 8 joints (gripper links: torso + 7 arm hinges + finger) take the 16-step instantiation, the wrists
    # alone the 8-step one; K = 1, 2 and 3 exercise the three lane layouts
    for links in ([21, 31], [20, 29], [22], [6, 21, 12]):
        J = eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, links, root_t, root_q, clip=True)
        Jo = oc.geometric_jacobian(ang, root_t, root_q, skeletons["hu_zero_pose/parents"].tolist(),
                                   T(skeletons["hu_zero_pose/offsets"]), oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER,
                                   True, links)
        assert maxdiff(J, Jo) <= 5e-6, links


# ------------------------------------------------------------------------------- host / streaming calls
def _same_refinement(a, b):
    """Two template instantiations of the fused kernel (16-warp / 8-warp / one-warp server) run the same algorithm but
    need not contract the FMA-fast refinement identically: agreement like kernel-vs-oracle, not bit equality."""
    err = (torch.as_tensor(a).cpu().float() - torch.as_tensor(b).cpu().float()).abs().reshape(len(a), -1).max(dim=-1).values.numpy()
    return float(np.quantile(err, 0.97)) <= ANGLE_TOL and float(err.max()) <= 1e-3


def test_host_call_and_streaming_match_device_call(hrt, eng, oc, skeletons):
    B = 70_000                                   # > one 65,536-frame pipeline chunk, ragged tail
    raw = oc.synth_clip_3q(B, seed=31, sk=skeletons)
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    lq, dof, lp = eng.retarget_body_quat(raw, flags=flags)
    h_in = raw.pin_memory()
    h_lq = torch.empty(B, 31, 4).pin_memory()
    h_dof = torch.empty(B, 30).pin_memory()
    h_lp = torch.empty(B, 31, 3).pin_memory()
    eng.retarget_body_quat_host(h_in, flags=flags, out_local_q=h_lq, out_dof=h_dof, out_link_pos=h_lp)
    # same kernel instantiation, same outputs: the chunked host pipeline must be bit-identical to the device call
    assert torch.equal(h_dof, dof.cpu()) and torch.equal(h_lp, lp.cpu()) and torch.equal(h_lq, lq.cpu())
    # the host call moves only the joint range the solver reads (vtrdyn joints 10-20) across PCIe: the other joints of
    # the caller's tensor are never looked at, whatever they hold
    # (dof-only calls: strided copy of the range; with link positions the whole rows travel, D2H being the bound)
    assert eng.host_input_bytes_per_frame(out_dof=True) == 11 * 16
    assert eng.host_input_bytes_per_frame(out_dof=True, out_link_pos=True) == 21 * 16
    h_in2 = raw.clone()
    h_in2[:, :10] = float("nan")
    h_in2 = h_in2.pin_memory()
    h_dof_n = torch.empty(B, 30).pin_memory()
    h_lp_n = torch.empty(B, 31, 3).pin_memory()
    eng.retarget_body_quat_host(h_in2, flags=flags, out_dof=h_dof_n, out_link_pos=h_lp_n)
    _, dof_n, lp_n = eng.retarget_body_quat(raw, flags=flags, want_local_q=False)
    assert torch.equal(h_dof_n, dof_n.cpu()) and torch.equal(h_lp_n, lp_n.cpu())
    h_dof_n.zero_()
    eng.retarget_body_quat_host(h_in2, flags=flags, out_dof=h_dof_n)
    _, dof_n, _ = eng.retarget_body_quat(raw, flags=flags, want_local_q=False, want_link_pos=False)
    assert torch.equal(h_dof_n, dof_n.cpu())
    # pageable host memory also works (dof only -> the 16-warp instantiation; compared like for like)
    h_dof2 = torch.empty(B, 30)
    eng.retarget_body_quat_host(raw, flags=flags, out_dof=h_dof2)
    _, dof2, _ = eng.retarget_body_quat(raw, flags=flags, want_local_q=False, want_link_pos=False)
    assert torch.equal(h_dof2, dof2.cpu()) and _same_refinement(h_dof2, dof)
    h_dof3 = torch.empty(B, 30)
    eng.retarget_body_quat_host(raw, flags=flags | hrt.BQ_PACKED_IK, out_dof=h_dof3)
    _, dof3, _ = eng.retarget_body_quat(raw, flags=flags | hrt.BQ_PACKED_IK, want_local_q=False, want_link_pos=False)
    assert torch.equal(h_dof3, dof3.cpu())
    # the closed form (exact-order arithmetic) IS bit-identical across instantiations
    lq_c, dof_c, _ = eng.retarget_body_quat(raw, flags=0)
    _, dof_c2, _ = eng.retarget_body_quat(raw, flags=0, want_local_q=False, want_link_pos=False)
    assert torch.equal(dof_c, dof_c2)
    # streaming: one frame at a time through the mapped mailboxes
    raw_np = raw.numpy()
    o_dof, o_lp, o_lq = np.empty(30, np.float32), np.empty((31, 3), np.float32), np.empty((31, 4), np.float32)
    import time
    for persistent in (False, True):             # launch per frame / resident server kernel (survives its idle time-out)
        got_dof, got_lp = [], []
        eng.stream_open(flags=flags, persistent=persistent)
        for i in range(64):
            eng.stream_frame(raw_np[i], o_lq, o_dof, o_lp)
            got_dof.append(o_dof.copy())
            got_lp.append(o_lp.copy())
            if persistent and i == 30:
                time.sleep(0.06)
        eng.stream_close()
        assert _same_refinement(np.stack(got_dof), dof[:64]) and float(np.abs(np.stack(got_lp) - lp[:64].cpu().numpy()).max()) <= 1e-3
        eng.stream_open(flags=0, persistent=persistent)
        for i in range(16):
            eng.stream_frame(raw_np[i], o_lq, o_dof, None)
            assert np.array_equal(o_dof, dof_c[i].cpu().numpy()) and np.array_equal(o_lq, lq_c[i].cpu().numpy())
        eng.stream_close()


# ------------------------------------------------------------------------------- full-size properties
def test_full_size_properties(hrt, eng, eng_hu, oc, skeletons):
    # config 2 size: 65,536 Hu configs.  FK -> global-to-local -> FK is the identity on rotations.
    g = torch.Generator().manual_seed(0)
    L = 65_536
    lo, hi = torch.tensor(oc.HU_DOF_LOWER), torch.tensor(oc.HU_DOF_UPPER)
    ang = (lo + (hi - lo) * (torch.rand(L, 32, generator=g) * 1.2 - 0.1)).cuda()
    gq, gt = eng_hu.fk_angles(hrt.TREE_ROBOT, ang, clip=True)
    assert float((gq.norm(dim=-1) - 1).abs().max()) <= 2e-6 and bool((gq[..., 3] >= 0).all())
    lq = eng_hu.local_from_global(hrt.TREE_ROBOT, gq)
    gq2, gt2 = eng_hu.fk_local_quats(hrt.TREE_ROBOT, lq)
    assert maxdiff(gq2, gq) <= 5e-6 and maxdiff(gt2, gt) <= POS_TOL
    # clipped angles: clamping twice changes nothing (idempotence), and matches FK of pre-clamped input
    gq3, gt3 = eng_hu.fk_angles(hrt.TREE_ROBOT, torch.minimum(torch.maximum(ang, lo.cuda()), hi.cuda()), clip=False)
    assert maxdiff(gt3, gt) <= 2e-6
    # bone lengths are preserved by FK
    par = skeletons["hu_zero_pose/parents"]
    off = T(skeletons["hu_zero_pose/offsets"]).cuda()
    seg = (gt[:, 1:] - gt[:, par[1:]]).norm(dim=-1)
    assert float((seg - off[1:].norm(dim=-1)).abs().max()) <= 2e-6
    # config 3 size: 2^20 frames through the fused pipeline with IK; angles inside limits, finite,
    # link positions consistent with an independent FK of the returned angles
    B = 1 << 20
    raw = oc.synth_clip_3q(B, seed=0, sk=skeletons).cuda()
    lq, dof, lp = eng.retarget_body_quat(raw, flags=hrt.BQ_CLAMP | hrt.BQ_IK)
    assert torch.isfinite(dof).all() and torch.isfinite(lp).all()
    arm = list(range(11, 18)) + list(range(20, 27))
    lo5, hi5 = torch.tensor(oc.HU_V5_DOF_LOWER).cuda(), torch.tensor(oc.HU_V5_DOF_UPPER).cuda()
    assert bool(((dof[:, arm] >= lo5[arm]) & (dof[:, arm] <= hi5[arm])).all())
    rest = [i for i in range(30) if i not in arm]
    assert float(dof[:, rest].abs().max()) == 0.0
    _, gt5 = eng.fk_angles(hrt.TREE_ROBOT, dof, clip=False)
    assert maxdiff(gt5, lp) <= POS_TOL
    # a checksum of a strided sample against the oracle
    idx = torch.arange(0, B, 257)
    _, dof_o, lp_o = oc.body_quat_pipeline(raw[idx].cpu(), skeletons, clamp=True, ik_iters=10)
    err = (dof[idx].cpu() - dof_o).abs().max(dim=-1).values
    assert float(np.quantile(err.numpy(), 0.98)) <= ANGLE_TOL and float(err.max()) <= 1e-3


def test_full_size_properties_position_path(hrt, eng, oc, skeletons, pm, parity):
    """Config 3p size: 2^20 frames through the position solver (16-warp CTAs).  Size-independent properties: frames are
    independent (any sub-range run on its own gives the same bits), the gripper DOFs take only their defined values, the
    limit-aware refinement keeps every arm hinge inside its limits, and a strided sample agrees with the oracle."""
    B = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(5)
    em = 0.4 * torch.randn(B, 59, 3, device="cuda", generator=g)
    lq = hrt.rotation3d.exp_map_to_quat(em)
    root = torch.zeros(B, 3, device="cuda")
    root[:, 2] = 1.0
    _, gt = eng.fk_local_quats(hrt.TREE_SOURCE_FULL, lq, root, exact=True)            # the clip, built on the device
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
    del gt, lq, em
    dof = torch.empty(B, 30, device="cuda")
    eng.retarget_full_body_pos(body, lh, rh, out=(None, dof, None))
    fin = torch.isfinite(dof).all(dim=-1)
    assert float(fin.float().mean()) > 0.9995
    for lo, hi in ((0, 1 << 16), (B - (1 << 16) - 48, B - 48), (123_456, 123_456 + 30_000)):
        part = torch.empty(hi - lo, 30, device="cuda")
        eng.retarget_full_body_pos(body[lo:hi], lh[lo:hi], rh[lo:hi], out=(None, part, None))
        same = (part == dof[lo:hi]) | (torch.isnan(part) & torch.isnan(dof[lo:hi]))
        assert bool(same.all()), (lo, hi)
    small = torch.empty(4096, 30, device="cuda")                                      # 8-warp CTAs
    eng.retarget_full_body_pos(body[5000:9096], lh[5000:9096], rh[5000:9096], out=(None, small, None))
    assert bool(((small == dof[5000:9096]) | torch.isnan(small)).all())
    grip = dof[fin][:, [18, 19, 27, 28]].abs()
    assert float(grip.max()) <= 0.044 + 1e-7                                          # full_body_pos_retargeter.py:199-215
    rest = [i for i in range(30) if i not in list(range(11, 20)) + list(range(20, 29))]
    assert float(dof[fin][:, rest].abs().max()) == 0.0
    # limits + refinement on the whole clip
    dof_r = torch.empty(B, 30, device="cuda")
    eng.retarget_full_body_pos(body, lh, rh, out=(None, dof_r, None), flags=hrt.POS_CLAMP | hrt.POS_IK)
    arm = list(range(11, 18)) + list(range(20, 27))
    lo5, hi5 = torch.tensor(oc.HU_V5_DOF_LOWER).cuda(), torch.tensor(oc.HU_V5_DOF_UPPER).cuda()
    ok = torch.isfinite(dof_r).all(dim=-1)
    assert float(ok.float().mean()) > 0.9995
    assert bool(((dof_r[ok][:, arm] >= lo5[arm]) & (dof_r[ok][:, arm] <= hi5[arm])).all())
    # strided sample against the oracle
    idx = torch.arange(0, B, 509)
    off = T(skeletons["vtrdyn_full_zero_pose/offsets"])
    zgt = T(skeletons["vtrdyn_full_zero_pose/global_translation"])
    with oc.exact_kabsch():
        _, dof_o, _ = oc.retarget_full_body_pos(body[idx].cpu(), lh[idx].cpu(), rh[idx].cpu(), off, zgt, True)
    st = pm.distance_stats(dof[idx], dof_o, robot_parents=skeletons["hu_v5_zero_pose/parents"].tolist(),
                           robot_offsets=skeletons["hu_v5_zero_pose/offsets"])
    parity.record("a32 full_body_pos 2^20 frames (strided sample of 2061) kernel vs exact-SVD oracle", st)
    check_dist(st, frac=0.97, p99=3e-5, dmax=2e-3, fk_max=2e-4)


# ------------------------------------------------------------------------------- position-input paths
def test_primitives_kabsch_vs_reference_golden(hrt, eng, golden):
    """cal_joint_quat through the full_body_pos kernel is covered below; here the known-answer arm of
    retarget/rotation_test.py:96-152 through the upper-body solver: zero pose in -> zero angles."""
    g = golden("upper_body")
    lq, dof = eng.retarget_upper_body(T(g["zero_pose_in"]))
    assert maxdiff(dof, g["zero_pose_dof"]) <= 2e-6
    assert float(dof.abs().max()) <= 1e-3


def test_full_body_pos_vs_reference_golden(hrt, eng, oc, golden, skeletons, pm, parity):
    g = golden("full_body_pos")
    body, lh, rh = T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"])
    lq, dof, bq = eng.retarget_full_body_pos(body, lh, rh)
    rp, ro = skeletons["hu_v5_zero_pose/parents"].tolist(), skeletons["hu_v5_zero_pose/offsets"]
    sd = g["self_delta"]
    # (1) against the UNMODIFIED reference's outputs: bounded by the MKL-SVD floor (see check_dist above)
    st = pm.distance_stats(dof, g["dof_pos"], lq, g["robot_local_q"], rp, ro)
    parity.record("a32 VtrdynFullBodyPosRetargeter kernel vs reference (256 golden frames)", st,
                  reference_self_delta_frac_le_1e5=float((sd <= ANGLE_TOL).mean()), reference_self_delta_p99=float(np.quantile(sd, .99)),
                  reference_self_delta_max=float(sd.max()),
                  conditioned_subset_coverage=float((sd < 2e-6).mean()))
    check_dist(st, frac=0.87, p99=7e-5, dmax=1.5e-4, fk_max=3e-5, geo_max=1.5e-4)    # measured .898 / 5.1e-5 / 9.1e-5 / 1.8e-5 / 9.1e-5 = the floor
    # (2) against the oracle with the exact SVD (the kernel's arithmetic class): north-star bars on every frame
    with oc.exact_kabsch():
        rl_x, dof_x, bq_x = oc.retarget_full_body_pos(body, lh, rh, T(skeletons["vtrdyn_full_zero_pose/offsets"]),
                                                      T(skeletons["vtrdyn_full_zero_pose/global_translation"]), True)
    st = pm.distance_stats(dof, dof_x, lq, rl_x, rp, ro)
    parity.record("a32 VtrdynFullBodyPosRetargeter kernel vs exact-SVD oracle (256 golden frames)", st)
    check_dist(st, frac=0.97, p99=2e-5, dmax=4e-5, fk_max=POS_TOL, geo_max=4e-5)     # measured .984 / 1.4e-5 / 2.3e-5 / 6.2e-6 / 2.3e-5
    # conditioned subsets (frames the reference itself moves little on under 1-ulp input jitter): coverage asserted
    e_all = (dof.cpu() - T(g["dof_pos"])).abs().max(dim=-1).values.numpy()
    for thr, min_cov, min_frac in ((2e-6, 0.12, 1.0), (5e-6, 0.60, 0.96), (1e-5, 0.80, 0.93)):
        cond = sd < thr
        e_c = e_all[cond]
        parity.record(f"a32 conditioned subset (reference self-delta < {thr:g}) kernel vs reference",
                      {"frames": int(cond.sum()), "coverage": float(cond.mean()), "frac_le_1e-5": float((e_c <= ANGLE_TOL).mean()),
                       "dof_max": float(e_c.max())})
        assert float(cond.mean()) >= min_cov and float((e_c <= ANGLE_TOL).mean()) >= min_frac, (thr, float(cond.mean()))
    # the published torso / wrist quaternions: Kabsch outputs, same floor (kabsch_fp64_vs_reference_quat: max 1.2e-5)
    st_q = {"bq_max_abs": maxdiff(bq, g["body_global_q"]), "bq_vs_exact_svd_max_abs": maxdiff(bq, bq_x)}
    parity.record("a32 body_global_rotation (torso + wrist Kabsch quaternions)", st_q)
    # vs the reference the near-planar 5-point wrist fits carry MKL's fp32 SVD noise (measured 7.4e-5); vs the exact SVD: 1 ulp
    assert st_q["bq_max_abs"] <= 2e-4 and st_q["bq_vs_exact_svd_max_abs"] <= 5e-7
    # binary gripper variant
    e2 = hrt.Engine(0).set_standard_trees()
    from humanoid_real_time_retarget_b200 import robot_config as cfg
    e2.configure_pos(hrt.engine.POS_FULL_BODY_POS, hrt.TREE_SOURCE_FULL, hrt.TREE_ROBOT,
                     cfg.skeleton_tables()["vtrdyn_full_zero_pose/global_translation"], False)
    _, dofb, _ = e2.retarget_full_body_pos(T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"]))
    assert maxdiff(dofb[:, [18, 19, 27, 28]], g["dof_pos_binary"][:, [18, 19, 27, 28]]) <= 1e-7
    # reference-named class, one frame at a time, CPU tensors in / out
    src = hrt.RobotZeroPose.from_asset("vtrdyn_full_zero_pose")
    tgt = hrt.RobotZeroPose.from_asset("hu_v5_zero_pose")
    solver = hrt.VtrdynFullBodyPosRetargeter(src, tgt, precise_gripper=True)
    for i in range(4):
        rl_i, dof_i, bq_i = solver.retarget(T(g["body_t"][i]), T(g["lhand_t"][i]), T(g["rhand_t"][i]))
        assert rl_i.shape == (31, 4) and dof_i.shape == (30,) and bq_i.shape == (59, 4) and dof_i.device.type == "cpu"
        assert torch.equal(dof_i, dof[i].cpu())


def test_full_body_pos_vs_oracle_large(hrt, eng, oc, skeletons, pm, parity):
    """20,000 frames of the SURVEY 8(d) config-3p clip: kernel vs the oracle as the reference computes it (MKL fp32 SVD;
    the oracle reproduces the per-frame reference to 1.2e-7 on every golden frame) and vs the oracle with the exact SVD."""
    B = 20_000
    g = torch.Generator().manual_seed(0)
    em = 0.4 * torch.randn(B, 59, 3, generator=g)
    lq = oc.exp_map_to_quat(em)
    root = torch.zeros(B, 3)
    root[:, 2] = 1.0
    _, gt = oc.cal_forward_kinematics(lq, root, skeletons["vtrdyn_full_zero_pose/parents"].tolist(),
                                      T(skeletons["vtrdyn_full_zero_pose/offsets"]))
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
    off = T(skeletons["vtrdyn_full_zero_pose/offsets"])
    zgt = T(skeletons["vtrdyn_full_zero_pose/global_translation"])
    rp, ro = skeletons["hu_v5_zero_pose/parents"].tolist(), skeletons["hu_v5_zero_pose/offsets"]
    rl_o, dof_o, bq_o = oc.retarget_full_body_pos(body, lh, rh, off, zgt, True)
    with oc.exact_kabsch():
        rl_x, dof_x, bq_x = oc.retarget_full_body_pos(body, lh, rh, off, zgt, True)
    # the reference's own sensitivity: the oracle against itself under 1-ulp input jitter (one probe, like the goldens)
    gj = torch.Generator().manual_seed(1)
    jit = lambda x: torch.nextafter(x, x + torch.sign(torch.randn(x.shape, generator=gj)))
    _, dof_j, _ = oc.retarget_full_body_pos(jit(body), jit(lh), jit(rh), off, zgt, True)
    st_self = pm.distance_stats(dof_j, dof_o, robot_parents=rp, robot_offsets=ro)
    parity.record("a32 oracle (MKL SVD) vs itself under 1-ulp input jitter (20k frames)", st_self)
    st_floor = pm.distance_stats(dof_x, dof_o, rl_x, rl_o, rp, ro)
    parity.record("a32 exact-SVD oracle vs oracle as the reference computes it (20k frames) = MKL sgesdd floor", st_floor)
    lq_k, dof_k, bq_k = eng.retarget_full_body_pos(body, lh, rh)
    st_ref = pm.distance_stats(dof_k, dof_o, lq_k, rl_o, rp, ro)
    parity.record("a32 kernel vs oracle as the reference computes it (20k frames)", st_ref)
    st_x = pm.distance_stats(dof_k, dof_x, lq_k, rl_x, rp, ro)
    parity.record("a32 kernel vs exact-SVD oracle (20k frames)", st_x)
    assert st_ref["finite_frames"] >= 0.999 * B
    # against the reference's arithmetic the kernel may not be further away than the exact-SVD restatement is (+ margin)
    assert st_ref["frac_le_1e-5"] >= st_floor["frac_le_1e-5"] - 0.02, (st_ref, st_floor)
    assert st_ref["dof_p99"] <= 1.5 * st_floor["dof_p99"] + 1e-5, (st_ref, st_floor)
    assert st_ref["fk_pos_p99_m"] <= 1.5 * st_floor["fk_pos_p99_m"] + 2e-6, (st_ref, st_floor)
    # against its own arithmetic class: the north-star bars, p99 on raw angles (2*acos(w) read-back stairs remain) and
    # FK link positions / geodesic on all frames except the degenerate tail (fits with sigma_2 ~ 0: < 0.1 % of random poses)
    check_dist(st_x, frac=0.97, p99=3e-5)
    assert st_x["fk_pos_p99_m"] <= POS_TOL, st_x
    assert st_x["geodesic_p99"] <= 3e-5, st_x


def test_pos_path_cta_shapes_agree(hrt, eng, oc, skeletons, pm, parity):
    """The position kernels are instantiated for 8 / 12 / 16 warps per CTA and the host picks by what fits in shared memory
    and by clip length (short clips: 8).  Same device code, so the same angles, whichever shape runs and whichever outputs
    are requested (local rotations and body quaternions are written straight to HBM, only dof_pos is staged)."""
    B = 40_000
    g = torch.Generator().manual_seed(3)
    em = 0.4 * torch.randn(B, 59, 3, generator=g)
    root = torch.zeros(B, 3)
    root[:, 2] = 1.0
    _, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, skeletons["vtrdyn_full_zero_pose/parents"].tolist(),
                                      T(skeletons["vtrdyn_full_zero_pose/offsets"]))
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous().cuda(), gt[:, 14:34].contiguous().cuda(), gt[:, 39:59].contiguous().cuda()
    dev = body.device
    lq8, dof8, bq8 = eng.retarget_full_body_pos(body, lh, rh)                                            # 8 warps
    dof16 = torch.empty(B, 30, device=dev)
    eng.retarget_full_body_pos(body, lh, rh, out=(None, dof16, None))                                    # 16 warps
    dof12, bq12 = torch.empty(B, 30, device=dev), torch.empty(B, 59, 4, device=dev)
    eng.retarget_full_body_pos(body, lh, rh, out=(None, dof12, bq12))                                    # 12 warps
    ok = torch.isfinite(dof8).all(dim=-1)
    assert float(ok.float().mean()) > 0.999
    for name, d in (("16 warps", dof16), ("12 warps", dof12)):
        diff = (d - dof8)[ok].abs().max().item()
        print(f"[pos cta shapes] {name} vs 8 warps: max |d dof| = {diff:.3e}")
        assert diff <= 1e-6, name
        assert torch.equal(torch.isfinite(d).all(dim=-1), ok)
    assert (bq12 - bq8)[ok].abs().max().item() <= 1e-6
    # short clip (8 warps) = the same frames of the long one
    dof_s = torch.empty(1000, 30, device=dev)
    eng.retarget_full_body_pos(body[:1000], lh[:1000], rh[:1000], out=(None, dof_s, None))
    assert (dof_s - dof16[:1000])[ok[:1000]].abs().max().item() <= 1e-6
    # upper-body path, long clip (16 warps) against its oracle on a sample
    _, dof_u = eng.retarget_upper_body(body)
    with oc.exact_kabsch():
        rl_o, dof_o = oc.retarget_upper_body(body[:4096].cpu(), T(skeletons["vtrdyn_zero_pose/offsets"]))
    st = pm.distance_stats(dof_u[:4096], dof_o, robot_parents=skeletons["hu_v5_zero_pose/parents"].tolist(),
                           robot_offsets=skeletons["hu_v5_zero_pose/offsets"])
    parity.record("a29 upper_body long clip (16-warp CTAs, 4096-frame sample) kernel vs exact-SVD oracle", st)
    assert st["finite_frames"] >= 0.999 * 4096
    check_dist(st, frac=0.97, p99=3e-5)


def test_full_body_pos_limits_and_refinement_vs_oracle(hrt, eng, oc, skeletons, golden):
    """Position path + joint limits + fused limit-aware refinement (builder-specified: kernel vs own oracle, the
    refinement applied by the oracle to the KERNEL's closed-form angles so that only the refinement is compared)."""
    g = golden("full_body_pos")
    body, lh, rh = T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"])
    lq0, dof0, _ = eng.retarget_full_body_pos(body, lh, rh)
    dof0 = dof0.cpu()
    # clamp only
    lq1, dof1, _ = eng.retarget_full_body_pos(body, lh, rh, flags=hrt.POS_CLAMP)
    ref1 = oc.refine_pos_dof(dof0, skeletons, clamp=True, ik_iters=0)
    assert torch.equal(dof1.cpu(), ref1)
    arm = list(range(11, 18)) + list(range(20, 27))
    lo5, hi5 = torch.tensor(oc.HU_V5_DOF_LOWER), torch.tensor(oc.HU_V5_DOF_UPPER)
    assert bool(((dof1.cpu()[:, arm] >= lo5[arm]) & (dof1.cpu()[:, arm] <= hi5[arm])).all())
    assert torch.equal(dof1.cpu()[:, [18, 19, 27, 28]], dof0[:, [18, 19, 27, 28]])            # gripper DOFs untouched
    # clamp + 10 refinement steps
    lq2, dof2, _ = eng.retarget_full_body_pos(body, lh, rh, flags=hrt.POS_CLAMP | hrt.POS_IK, ik_iters=10)
    ref2 = oc.refine_pos_dof(dof0, skeletons, clamp=True, ik_iters=10)
    err = (dof2.cpu() - ref2).abs().max(dim=-1).values
    print(f"pos refinement vs oracle: within 1e-5 {float((err <= ANGLE_TOL).float().mean()):.4f}; p99 {float(np.quantile(err.numpy(), 0.99)):.2e} max {float(err.max()):.2e}")
    assert float((err <= ANGLE_TOL).float().mean()) >= 0.97 and float(err.max()) <= 1e-3
    assert bool(((dof2.cpu()[:, arm] >= lo5[arm]) & (dof2.cpu()[:, arm] <= hi5[arm])).all())
    # a frame whose closed form is inside the limits is a fixed point of the refinement
    inside = ((dof0[:, arm] >= lo5[arm]) & (dof0[:, arm] <= hi5[arm])).all(dim=-1)
    if bool(inside.any()):
        assert float((dof2.cpu()[inside] - dof0[inside]).abs().max()) <= 2e-6
    # where limits bite, the refinement brings the wrist closer to the unclamped pose than clamping alone
    _, gt_u = eng.fk_angles(hrt.TREE_ROBOT, dof0, clip=False)
    _, gt_c = eng.fk_angles(hrt.TREE_ROBOT, dof1, clip=False)
    _, gt_r = eng.fk_angles(hrt.TREE_ROBOT, dof2, clip=False)
    wr = [18, 27]
    d_c = (gt_c[:, wr] - gt_u[:, wr]).norm(dim=-1).sum(-1)
    d_r = (gt_r[:, wr] - gt_u[:, wr]).norm(dim=-1).sum(-1)
    bite = ~inside.cuda()
    assert bool(bite.any()) and float(d_r[bite].mean()) < 0.9 * float(d_c[bite].mean())
    # published local rotations are those of the refined angles
    ax = torch.eye(3)[oc.HU_V5_DOF_AXIS]
    q = oc.quat_from_angle_axis(dof2.cpu()[:, 11:18].reshape(-1), ax[11:18].repeat(dof2.shape[0], 1))
    assert maxdiff(lq2[:, 12:19].reshape(-1, 4), q) <= 2e-6


def test_upper_body_and_full_body_vs_reference_golden(hrt, eng, oc, golden, skeletons, pm, parity):
    rp, ro = skeletons["hu_v5_zero_pose/parents"].tolist(), skeletons["hu_v5_zero_pose/offsets"]
    g = golden("upper_body")
    lq, dof = eng.retarget_upper_body(T(g["global_t"]))
    st = pm.distance_stats(dof, g["dof_pos"], lq, g["robot_local_q"], rp, ro)
    parity.record("a29 HuUpperBodyFromMocapRetarget kernel vs reference (256 golden frames)", st)
    check_dist(st, frac=0.92, p99=1.5e-4, dmax=STAIR, fk_max=2.5e-4, geo_max=STAIR)      # measured .949 / 9.7e-5 / 2.7e-4 / 1.1e-4 / 4.9e-4; MKL-SVD floor .95 / 5.9e-5 / 1.8e-4 / 8.5e-5 / 4.9e-4
    with oc.exact_kabsch():
        rl_x, dof_x = oc.retarget_upper_body(T(g["global_t"]), T(skeletons["vtrdyn_zero_pose/offsets"]))
    st = pm.distance_stats(dof, dof_x, lq, rl_x, rp, ro)
    parity.record("a29 HuUpperBodyFromMocapRetarget kernel vs exact-SVD oracle (256 golden frames)", st)
    check_dist(st, frac=0.98, p99=2e-5, dmax=STAIR, fk_p99=POS_TOL, geo_max=STAIR)      # measured .992 / 7.6e-6 / 2.7e-4 (one frame on another stair step)
    used = [11, 12, 13, 14, 20, 21, 22, 23]
    rest = [i for i in range(30) if i not in used]
    assert float(dof[:, rest].abs().max()) == 0.0          # 8 of 30 DOFs are ever non-zero
    # a31 has no Kabsch step (measured parent / wrist quaternions): every frame to the north-star bars against the reference
    g = golden("full_body")
    lq, dof = eng.retarget_full_body(T(g["body_q"]), T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"]))
    st = pm.distance_stats(dof, g["dof_pos"], lq, g["robot_local_q"], rp, ro)
    parity.record("a31 VtrdynFullBodyRetargeter kernel vs reference (256 golden frames, no Kabsch on this path)", st)
    check_dist(st, frac=0.99, p99=ANGLE_TOL, dmax=3e-5, fk_max=POS_TOL, geo_max=3e-5)
    assert maxdiff(dof[:, [18, 19, 27, 28]], g["dof_pos"][:, [18, 19, 27, 28]]) <= 1e-7
    src = hrt.RobotZeroPose.from_asset("vtrdyn_zero_pose")
    tgt = hrt.RobotZeroPose.from_asset("hu_v5_zero_pose")
    s = hrt.HuUpperBodyFromMocapRetarget(src, tgt)
    rl_i, dof_i = s.retarget_from_global_translation(T(golden("upper_body")["global_t"][3]))
    assert rl_i.shape == (31, 4) and dof_i.shape == (30,)
