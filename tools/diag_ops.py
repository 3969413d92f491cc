"""Diagnostic: max |diff| and #mismatching rows of each element-wise op against the golden vectors."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanoid_real_time_retarget_b200 as hrt
T = torch.from_numpy
def G(n):
    z = np.load(os.path.join(ROOT, "tests", "golden", n + ".npz")); return {k: z[k] for k in z.files}
def rep(name, a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    d = np.abs(a.astype(np.float64) - np.asarray(b).astype(np.float64))
    bad = np.argwhere(d.reshape(d.shape[0], -1).max(-1) > 0)[:, 0] if d.ndim else []
    print(f"{name:36s} max {d.max():.3e}  rows differing {len(bad)}/{d.shape[0] if d.ndim else 1}  first {list(bad[:5])}")
    return bad
r, t = hrt.rotation3d, hrt.transform3d
g1, g2 = G("rotation_ops"), G("rotation_ops2")
bad = rep("quat_from_rotation_matrix", r.quat_from_rotation_matrix(T(g1["R"])), g1["quat_from_rotation_matrix"])
for i in bad[:4]:
    print(i, g1["R"][i].tolist(), r.quat_from_rotation_matrix(T(g1["R"][i:i+1])).tolist(), g1["quat_from_rotation_matrix"][i].tolist())
v, w, nn = T(g2["v"]), T(g2["w"]), T(g2["nn"])
bad = rep("quat_between_two_vecs", t.quat_between_two_vecs(v, w), g2["quat_between_two_vecs"])
rep("proj_in_plane", t.proj_in_plane(v, nn), g2["proj_in_plane"])
rep("radians_between", t.radians_between_vecs(v, w, nn), g2["radians_between_vecs"])
rep("coord_transform", t.coord_transform(v, order=[2, 0, 1], dir=torch.Tensor([-1, 1, -1])), g2["coord_transform"])
rep("quat_slerp", t.quat_slerp(r.quat_normalize(T(g2["qa"])), r.quat_normalize(T(g2["qb"])), T(g2["tt"])), g2["quat_slerp"])
rep("quat_to_dof_pos", t.quat_to_dof_pos(T(g2["q30"]), g2["dof_axis30"].tolist()), g2["quat_to_dof_pos"])
for seq in ["xyz", "zyx", "XYZ", "YXZ", "ZYX", "ZXZ", "yzy", "XZY"]:
    qs = t.quat_in_xyz_axis(T(g2["qa"]), seq)
    for m in range(3):
        rep(f"euler {seq} {m}", qs[m], g2[f"euler_{seq}_{m+1}"])
rep("quat_to_eular", r.quat_to_eular(g2["qa"]), g2["quat_to_eular"])
zero = torch.randn(384, 4, 3, generator=torch.Generator().manual_seed(5)); mot = torch.randn(384, 4, 3, generator=torch.Generator().manual_seed(6))
rep("kabsch n4", t.cal_joint_quat(zero, mot), g2["cal_joint_quat_n4"])
gp = G("primitives")
rep("kabsch3", t.cal_joint_quat(T(gp["Z3"]), T(gp["M3"])), gp["kabsch3"])
rep("kabsch5", t.cal_joint_quat(T(gp["Z5"]), T(gp["M5"])), gp["kabsch5"])
ge = G("euler")
for seq in ["XYZ", "YXZ", "ZYX"]:
    qs = t.quat_in_xyz_axis(T(ge["q"]), seq)
    for m in range(3):
        rep(f"euler.npz {seq} {m}", qs[m], ge[f"q{m+1}_{seq}"])
