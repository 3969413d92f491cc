"""Extract the numeric content of the reference's pickled SkeletonState assets
(asset/**/*.pkl, SURVEY.md section 2 "(data)") into one .npz that travels with the repo.

Run in the dev container only:  python tools/extract_assets.py
Writes humanoid_real_time_retarget_b200/data/skeletons.npz with, per asset <name>:
  <name>/node_names, /parents (int32), /offsets (J,3 f32 = tree local_translation),
  /local_rotation (J,4), /global_rotation (J,4), /global_translation (J,3), /root_translation (3,)
plus the import-time T-pose -> zero-pose tables of retarget/utils/parse_mocap.py:71-78,97-104:
  t2z/vtrdyn (21,4), t2z/vtrdyn_full (59,4)
and the zero poses those modules rebuild (offsets after rebuild_pose_by_local_rotation).
The md5 of every source pickle is stored so tests can cite BASELINE.md.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402

ASSETS = {
    "hu_v5_zero_pose": "asset/hu_pose/hu_v5_zero_pose.pkl",
    "hu_zero_pose": "asset/zero_pose/hu_zero_pose.pkl",
    "new_hu_zero_pose": "asset/zero_pose/new_hu_zero_pose.pkl",
    "hu_start_pose": "asset/start_pose/hu_start_pose.pkl",
    "vtrdyn_t_pose": "asset/t_pose/vtrdyn_t_pose.pkl",
    "vtrdyn_zero_pose": "asset/zero_pose/vtrdyn_zero_pose.pkl",
    "vtrdyn_full_t_pose": "asset/t_pose/vtrdyn_full_t_pose.pkl",
    "vtrdyn_full_zero_pose": "asset/zero_pose/vtrdyn_full_zero_pose.pkl",
    "noitom_t_pose": "asset/t_pose/noitom_t_pose.pkl",
    "noitom_zero_pose": "asset/zero_pose/noitom_zero_pose.pkl",
    "smpl_t_pose": "asset/t_pose/smpl_t_pose.pkl",
}


def main():
    ref = ref_shim.load()
    out = {}
    for name, rel in ASSETS.items():
        st = ref_shim.load_asset(ref, rel)
        tree = st.skeleton_tree
        with open(os.path.join(ref.root, rel), "rb") as f:
            md5 = hashlib.md5(f.read()).hexdigest()
        out[f"{name}/md5"] = np.array(md5)
        out[f"{name}/node_names"] = np.array(list(tree.node_names))
        out[f"{name}/parents"] = tree.parent_indices.numpy().astype(np.int32)
        out[f"{name}/offsets"] = tree.local_translation.numpy().astype(np.float32)
        out[f"{name}/tree_quat"] = tree.quat.numpy().astype(np.float32) if hasattr(tree, "quat") else None
        out[f"{name}/local_rotation"] = st.local_rotation.numpy().astype(np.float32)
        out[f"{name}/global_rotation"] = st.global_rotation.numpy().astype(np.float32)
        out[f"{name}/global_translation"] = st.global_translation.numpy().astype(np.float32)
        out[f"{name}/root_translation"] = st.root_translation.numpy().astype(np.float32)
        if out[f"{name}/tree_quat"] is None:
            del out[f"{name}/tree_quat"]
    pm = ref.parse_mocap
    out["t2z/vtrdyn"] = pm.vtrdyn_t2zero_pose_transform_quat.numpy().astype(np.float32)
    out["t2z/vtrdyn_full"] = pm.vtrdyn_full_t2zero_pose_transform_quat.numpy().astype(np.float32)
    # the RobotZeroPose objects parse_mocap rebuilds in place (offsets of the derived zero pose)
    out["rebuilt/vtrdyn/offsets"] = pm.vtrdyn_zero_pose.local_translation.numpy().astype(np.float32)
    out["rebuilt/vtrdyn_full/offsets"] = pm.vtrdyn_full_zero_pose.local_translation.numpy().astype(np.float32)
    dst = os.path.join(os.path.dirname(__file__), "..", "humanoid_real_time_retarget_b200", "data", "skeletons.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.normpath(dst), len(out), "arrays")
    for name in ASSETS:
        print(name, out[f"{name}/parents"].shape[0], str(out[f"{name}/md5"]))


if __name__ == "__main__":
    main()
