"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

    python tools/make_golden.py

Every array stored here is an input to, or an output of, the reference's own code imported
from /root/reference through tools/ref_shim.py (torch CPU fp32 + scipy).  The oracle
(oracle/retarget_oracle.py) and the CUDA path are both checked against these fixtures.
Versions used are recorded in tests/golden/MANIFEST.json.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
FULL2BODY = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]


def save(name, **arrs):
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrs.items()))


def rand_quat(n, gen):
    q = torch.randn(n, 4, generator=gen)
    return q / q.norm(dim=-1, keepdim=True)


def clip_3q(ref, L, seed=0, scale=0.5):
    """SURVEY.md 8(d) config 3q recipe: local exp-maps scale*N(0,1) on the vtrdyn T-pose tree."""
    t_pose = ref_shim.load_asset(ref, "asset/t_pose/vtrdyn_t_pose.pkl")
    tree = t_pose.skeleton_tree
    g = torch.Generator().manual_seed(seed)
    em = scale * torch.randn(L, tree.num_joints, 3, generator=g)
    lq = ref.r3d.exp_map_to_quat(em)
    gq, _ = ref.rkm.cal_forward_kinematics(lq, torch.zeros(L, 3), tree.parent_indices.tolist(),
                                           tree.local_translation)
    return gq


def clip_3p(ref, L, seed=0, scale=0.4):
    """SURVEY.md 8(d) config 3p recipe on the vtrdyn_full zero-pose tree, root at (0,0,1)."""
    zp = ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_full_zero_pose.pkl")
    tree = zp.skeleton_tree
    g = torch.Generator().manual_seed(seed)
    em = scale * torch.randn(L, tree.num_joints, 3, generator=g)
    lq = ref.r3d.exp_map_to_quat(em)
    root = torch.zeros(L, 3)
    root[:, 2] = 1.0
    gq, gt = ref.rkm.cal_forward_kinematics(lq, root, tree.parent_indices.tolist(), tree.local_translation)
    return gq, gt


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    r3d, t3d, rkm = ref.r3d, ref.t3d, ref.rkm
    g = torch.Generator().manual_seed(1234)

    # ---- 1. rotation3d ops ---------------------------------------------------------------
    n = 512
    qa, qb = rand_quat(n, g), rand_quat(n, g)
    qa[::7] *= -1.0
    v = torch.randn(n, 3, generator=g)
    raw = torch.randn(n, 4, generator=g) * 3
    ang = (torch.rand(n, generator=g) - 0.5) * 4 * np.pi
    ax = torch.randn(n, 3, generator=g)
    em = torch.randn(n, 3, generator=g) * 0.8
    em[:8] *= 1e-4
    small = torch.tensor([0., 1e-6, 1e-5, 1e-4, 3e-4, 7e-4, 1e-3, 1e-2])
    qsmall = r3d.quat_from_angle_axis(small, torch.tensor([[0., 1., 0.]]).repeat(8, 1))
    # rotation matrices: random + special (identity, pi about axes, 120deg about (1,1,1))
    R = r3d.rot_matrix_from_quaternion(rand_quat(n, g))
    special = torch.tensor([
        [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
        [[-1, 0, 0], [0, -1, 0], [0, 0, 1]],
        [[1, 0, 0], [0, -1, 0], [0, 0, -1]],
        [[-1, 0, 0], [0, 1, 0], [0, 0, -1]],
        [[0, -1, 0], [1, 0, 0], [0, 0, 1]],
        [[0, 0, 1], [1, 0, 0], [0, 1, 0]],
        [[1, 0, 0], [0, 0, -1], [0, 1, 0]],   # test_rotation.py:32-37 -> [0.7071,0,0,0.7071]
    ], dtype=torch.float32)
    R = torch.cat([special, R])
    save("rotation_ops",
         qa=qa, qb=qb, v=v, raw=raw, ang=ang, ax=ax, em=em, R=R, qsmall=qsmall,
         quat_mul=r3d.quat_mul(qa, qb), quat_mul_norm=r3d.quat_mul_norm(qa, qb),
         quat_normalize=r3d.quat_normalize(raw), quat_rotate=r3d.quat_rotate(qa, v),
         quat_from_angle_axis=r3d.quat_from_angle_axis(ang, ax),
         quat_from_rotation_matrix=r3d.quat_from_rotation_matrix(R),
         quat_to_exp_map=r3d.quat_to_exp_map(r3d.quat_normalize(qa)),
         quat_to_exp_map_small=r3d.quat_to_exp_map(qsmall),
         exp_map_to_quat=r3d.exp_map_to_quat(em),
         quat_mul_three=r3d.quat_mul_three(qa, qb, qa), quat_mul_four=r3d.quat_mul_four(qa, qb, qb, qa))

    # ---- 2. FK / local rotation ----------------------------------------------------------
    hu = ref_shim.load_asset(ref, "asset/zero_pose/hu_zero_pose.pkl")
    tree = hu.skeleton_tree
    model = ref.hfm.HuForwardModel(tree, device="cpu")
    L = 256
    lo, hi = ref.hu_cfg.Hu_DOF_LOWER, ref.hu_cfg.Hu_DOF_UPPER
    u = torch.rand(L, 32, generator=g) * 1.2 - 0.1          # widened 10% so the clamp is exercised
    angles = (lo + (hi - lo) * u).reshape(L, 32, 1)
    angles[0] = 0
    root_t = torch.randn(L, 3, generator=g)
    root_q = rand_quat(L, g).reshape(L, 1, 4)
    root_q = r3d.quat_normalize(root_q)
    root_t[0] = 0
    root_q[0] = torch.tensor([0., 0, 0, 1])
    gq_c, gt_c = model.forward_kinematics(angles.clone(), root_t, root_q, True)
    gq_n, gt_n = model.forward_kinematics(angles.clone(), root_t, root_q, False)
    save("fk_hu", angles=angles, root_t=root_t, root_q=root_q, gq_clip=gq_c, gt_clip=gt_c, gq_noclip=gq_n, gt_noclip=gt_n)

    for name, rel in [("hu_v5_zero_pose", "asset/hu_pose/hu_v5_zero_pose.pkl"),
                      ("vtrdyn_t_pose", "asset/t_pose/vtrdyn_t_pose.pkl"),
                      ("vtrdyn_full_zero_pose", "asset/zero_pose/vtrdyn_full_zero_pose.pkl")]:
        st = ref_shim.load_asset(ref, rel)
        tr = st.skeleton_tree
        J = tr.num_joints
        lq = r3d.quat_normalize(torch.randn(128, J, 4, generator=g))
        lq[0] = torch.tensor([0., 0, 0, 1])
        rt = torch.randn(128, 3, generator=g)
        gq, gt = rkm.cal_forward_kinematics(lq, rt, tr.parent_indices.tolist(), tr.local_translation)
        back = rkm.cal_local_rotation(gq, tr.parent_indices.tolist())
        # poselib SkeletonState FK for the same input (skeleton3d.py:402-425, :460-484)
        ss = ref.sk3d.SkeletonState.from_rotation_and_root_translation(tr, lq, rt, is_local=True)
        save(f"fk_{name}", local_q=lq, root_t=rt, gq=gq, gt=gt, local_back=back,
             sk_global_rotation=ss.global_rotation, sk_global_translation=ss.global_translation)

    # ---- 3. zero-pose transform ----------------------------------------------------------
    pm = ref.parse_mocap
    q21 = r3d.quat_normalize(torch.randn(128, 21, 4, generator=g))
    q59 = r3d.quat_normalize(torch.randn(64, 59, 4, generator=g))
    save("zero_pose_transform", q21=q21, q59=q59,
         vtrdyn=pm.vtrdyn_zero_pose_transform(q21), vtrdyn_full=pm.vtrdyn_full_zero_pose_transform(q59),
         vtrdyn_broadcast=pm.vtrdyn_broadcast_zero_pose_transform(q21))

    # ---- 4. Euler split ------------------------------------------------------------------
    qe = rand_quat(768, g)
    # gimbal / near-gimbal cases: second angle +-pi/2 (asymmetric sequences)
    import warnings
    from scipy.spatial.transform import Rotation as sRot
    for i, d in enumerate([0.0, 1e-9, 1e-8, 5e-8, 1e-7, 2e-7, 1e-6, 1e-4]):
        for s, sgn in enumerate([1.0, -1.0]):
            e = np.array([0.3, sgn * (np.pi / 2 - d), -0.7])
            for k, seq in enumerate(["XYZ", "YXZ", "ZYX"]):
                qe[(i * 2 + s) * 3 + k] = torch.from_numpy(sRot.from_euler(seq, e).as_quat().astype(np.float32))
    qe[60] = torch.tensor([0., 0, 0, 1])
    qe[61] = torch.tensor([0., 0, 0, -1])
    qe[62] = torch.tensor([1., 0, 0, 0])
    qe[63] = torch.tensor([0., 1, 0, 0])
    qe[64] = torch.tensor([0., 0, 1, 0])
    qe[65] = torch.tensor([0.5, 0.5, 0.5, 0.5])
    eul = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for seq in ["XYZ", "YXZ", "ZYX"]:
            eul[f"angles_{seq}"] = sRot.from_quat(qe.numpy()).as_euler(seq)
            # scipy 1.18's single-axis from_euler rejects batches: the reference only works per quat
            trip = [t3d.quat_in_xyz_axis(qe[i], seq) for i in range(qe.shape[0])]
            for n in range(3):
                eul[f"q{n + 1}_{seq}"] = torch.stack([t[n] for t in trip])
    save("euler", q=qe, **eul)

    # ---- 5. quaternion path (config 3q): zero-pose transform + Mocap2HuBodyRetargeter ------
    src21 = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl"))
    tgt = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/hu_pose/hu_v5_zero_pose.pkl"))
    L = 384
    raw_gq = clip_3q(ref, L)
    raw_gq[0] = torch.tensor([0., 0, 0, 1])             # sensor T-pose frame
    zq = pm.vtrdyn_zero_pose_transform(raw_gq)
    solver = ref.solvers.Mocap2HuBodyRetargeter(src21, tgt)
    rls, dofs, rlt, doft = [], [], [], []
    import warnings as _w
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        for i in range(L):
            rl, dof = solver.retarget_from_pose(zq[i])
            rls.append(rl); dofs.append(dof)
            rl, dof = solver.retarget_test(zq[i])
            rlt.append(rl); doft.append(dof)
    rl3q = torch.stack(rls)
    fk_gq, fk_gt = rkm.cal_forward_kinematics(rl3q, torch.zeros(L, 3), tgt.parent_indices.tolist(), tgt.local_translation)
    save("body_quat", raw_global_q=raw_gq, zero_pose_q=zq, robot_local_q=rl3q, dof_pos=torch.stack(dofs),
         robot_local_q_test=torch.stack(rlt), dof_pos_test=torch.stack(doft), fk_gq=fk_gq, fk_gt=fk_gt)

    # ---- 6. position paths (a29, a31, a32) -------------------------------------------------
    src59 = rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_full_zero_pose.pkl"))
    L = 256
    gq59, gt59 = clip_3p(ref, L)
    body_t, lh, rh = gt59[:, FULL2BODY], gt59[:, 14:34], gt59[:, 39:59]
    body_q = gq59[:, FULL2BODY]

    def run_pos(precise, bt, l, r):
        s = ref.solvers.VtrdynFullBodyPosRetargeter(src59, tgt, precise_gripper=precise)
        a, b, c = [], [], []
        with _w.catch_warnings():
            _w.simplefilter("ignore")
            for i in range(bt.shape[0]):
                x, y, z = s.retarget(bt[i], l[i], r[i])
                a.append(x); b.append(y.clone()); c.append(z)
        return torch.stack(a), torch.stack(b), torch.stack(c)

    rl_p, dof_p, bq_p = run_pos(True, body_t, lh, rh)
    _, dof_b, _ = run_pos(False, body_t, lh, rh)
    # conditioning probe: the reference against itself under 1-ulp input jitter
    def jitter(x):
        return torch.nextafter(x, x + torch.sign(torch.randn(x.shape, generator=g)))
    _, dof_j, _ = run_pos(True, jitter(body_t), jitter(lh), jitter(rh))
    self_delta = (dof_j - dof_p).abs().max(dim=-1).values
    save("full_body_pos", body_t=body_t, lhand_t=lh, rhand_t=rh, robot_local_q=rl_p, dof_pos=dof_p,
         dof_pos_binary=dof_b, body_global_q=bq_p, self_delta=self_delta)

    s = ref.solvers.VtrdynFullBodyRetargeter(src59, tgt)
    a, b = [], []
    with _w.catch_warnings():
        _w.simplefilter("ignore")
        for i in range(L):
            x, y = s.retarget(body_q[i], body_t[i], None, lh[i], None, rh[i])
            a.append(x); b.append(y.clone())
    save("full_body", body_q=body_q, body_t=body_t, lhand_t=lh, rhand_t=rh, robot_local_q=torch.stack(a), dof_pos=torch.stack(b))

    # upper body: positions of a 21-joint vtrdyn skeleton; the solver flips x,y itself (:41)
    zp21 = ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl")
    tr21 = zp21.skeleton_tree
    gg = torch.Generator().manual_seed(7)
    lq = r3d.exp_map_to_quat(0.4 * torch.randn(L, 21, 3, generator=gg))
    root = torch.zeros(L, 3); root[:, 2] = 1.0
    _, gt21 = rkm.cal_forward_kinematics(lq, root, tr21.parent_indices.tolist(), tr21.local_translation)
    gt21 = gt21 * torch.tensor([-1., -1., 1.])            # so that the solver's own flip undoes it
    s = ref.solvers.HuUpperBodyFromMocapRetarget(src21, tgt)
    a, b = [], []
    for i in range(L):
        x, y = s.retarget_from_global_translation(gt21[i])
        a.append(x); b.append(y)
    zero_in = (zp21.global_translation * torch.tensor([-1., -1., 1.])).unsqueeze(0)
    zx, zy = s.retarget_from_global_translation(zero_in[0])
    save("upper_body", global_t=gt21, robot_local_q=torch.stack(a), dof_pos=torch.stack(b),
         zero_pose_in=zero_in, zero_pose_dof=zy.unsqueeze(0), zero_pose_rl=zx.unsqueeze(0))

    # ---- 7. cal_joint_quat / shoulder / elbow primitives ------------------------------------
    off59 = src59.local_translation
    Rq = rand_quat(L, g)
    Z3 = off59[[11, 36, 34]].unsqueeze(0).repeat(L, 1, 1)
    M3 = r3d.quat_rotate(Rq.unsqueeze(1), Z3) + 0.002 * torch.randn(L, 3, 3, generator=g)
    Z5 = off59[[16, 20, 24, 28, 32]].unsqueeze(0).repeat(L, 1, 1)
    M5 = r3d.quat_rotate(Rq.unsqueeze(1), Z5) + 0.002 * torch.randn(L, 5, 3, generator=g)
    kj3 = torch.cat([t3d.cal_joint_quat(Z3[i:i + 1], M3[i:i + 1]) for i in range(L)])
    kj5 = torch.cat([t3d.cal_joint_quat(Z5[i:i + 1], M5[i:i + 1]) for i in range(L)])
    v1 = torch.randn(L, 3, generator=g) * 0.3
    pq = rand_quat(L, g)
    pr = [ref.fbp.cal_shoulderPR(v1[i], off59[13], pq[i]) for i in range(L)]
    ey = [ref.fbp.cal_elbowP_and_shoulderY(v1[i], off59[14], pq[i]) for i in range(L)]
    save("primitives", Z3=Z3, M3=M3, Z5=Z5, M5=M5, kabsch3=kj3, kabsch5=kj5, v1=v1, parent_q=pq,
         v0_upper=off59[13], v0_lower=off59[14],
         sh_pitch=torch.stack([p[0] for p in pr]), sh_roll=torch.stack([p[1] for p in pr]),
         sh_yaw=torch.stack([p[0] for p in ey]), el_pitch=torch.stack([p[1] for p in ey]))

    import scipy
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"torch": torch.__version__, "scipy": scipy.__version__, "numpy": np.__version__,
                   "generator": "tools/make_golden.py", "reference": ref.root}, f, indent=1)


if __name__ == "__main__":
    main()
