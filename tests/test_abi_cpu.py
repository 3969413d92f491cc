"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/hrt_b200.h declares; the host tables are bit-exact; the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from humanoid_real_time_retarget_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "hrt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hrt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    from humanoid_real_time_retarget_b200 import EXPORTED_SYMBOLS
    assert declared == set(EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(built, name) is not None
    assert built.hrt_abi_version() == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "humanoid_real_time_retarget_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built):
    import humanoid_real_time_retarget_b200 as hrt
    with pytest.raises(hrt.HrtError):
        hrt.Engine(0)
    with pytest.raises(hrt.HrtError):
        hrt.cal_local_rotation(torch.zeros(2, 21, 4), hrt.robot_config.vtrdyn_parent_indices)
    # the C ABI itself also refuses: no device -> HRT_E_NO_DEVICE, not a silent CPU path
    import ctypes as C
    h = C.c_void_p()
    assert built.hrt_ctx_create(0, C.byref(h)) == -5
    assert b"no CPU path" in built.hrt_last_error_string()


def test_tables_bit_exact(skeletons):
    from humanoid_real_time_retarget_b200 import robot_config as cfg
    from oracle import retarget_oracle as oc
    assert cfg.Hu_v5_DOF_AXIS == oc.HU_V5_DOF_AXIS == [2, 0, 1, 1, 1, 2, 0, 1, 1, 1, 2, 1, 0, 2, 1, 0, 1, 2, 1, 1,
                                                       1, 0, 2, 1, 0, 1, 2, 1, 1, 2]
    assert cfg.Hu_DOF_AXIS == oc.HU_DOF_AXIS and len(cfg.Hu_DOF_AXIS) == 32
    assert cfg.Hu_DOF_LOWER == oc.HU_DOF_LOWER and cfg.Hu_DOF_UPPER == oc.HU_DOF_UPPER
    assert len(cfg.Hu_v5_DOF_LOWER) == 30 == len(cfg.Hu_v5_DOF_UPPER)
    assert cfg.vtrdyn_parent_indices == skeletons["vtrdyn_zero_pose/parents"].tolist()
    assert cfg.vtrdyn_parent_indices == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 11, 10, 13, 14, 15, 10, 17, 18, 19]
    assert [str(s) for s in skeletons["vtrdyn_zero_pose/node_names"]] == cfg.VTRDYN_JOINT_NAMES
    # SURVEY appendix A
    assert skeletons["hu_v5_zero_pose/parents"].tolist() == [-1, 0, 1, 2, 3, 4, 0, 6, 7, 8, 9, 0, 11, 12, 13, 14, 15, 16,
                                                             17, 18, 18, 11, 21, 22, 23, 24, 25, 26, 27, 27, 11]
    assert cfg.BODY_23_TO_21 == [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]
    # T2Z is non-identity only at the arm joints
    t2z = skeletons["t2z/vtrdyn"]
    ident = np.array([0, 0, 0, 1], np.float32)
    nonid = [j for j in range(21) if not np.array_equal(t2z[j], ident)]
    assert nonid == [14, 15, 16, 18, 19, 20]
    # asset md5s recorded in BASELINE.md
    assert str(skeletons["hu_v5_zero_pose/md5"]) == "7a3627cd3fa4cb2d3fa5307d86f8d25c"
    assert str(skeletons["vtrdyn_zero_pose/md5"]) == "c3e53235468e32f790fa8f7db17fbd1b"


def test_rot_op_table_matches_header(built):
    """Every HRT_OP_* of include/hrt_b200.h is known to the library, in the same order as the Python codes."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "hrt_b200.h")).read()
    names = re.findall(r"^\s+(HRT_OP_[A-Z0-9_]+)\b", hdr, flags=re.M)
    assert names[-1] == "HRT_OP_COUNT"
    from humanoid_real_time_retarget_b200 import rotation3d as r3d
    for code, name in enumerate(names[:-1]):
        assert getattr(r3d, name[4:]) == code
        ni, no = C.c_int(), C.c_int()
        wi, wo = (C.c_int * 4)(), (C.c_int * 3)()
        assert built.hrt_rot_op_info(code, C.byref(ni), wi, C.byref(no), wo) == 0
        assert 1 <= ni.value <= 4 and 1 <= no.value <= 3 and all(1 <= w <= 9 for w in list(wi)[:ni.value])
    assert built.hrt_rot_op_info(len(names) - 1, C.byref(ni), wi, C.byref(no), wo) == -1


def test_compat_shims_and_asset_pickles():
    """The reference's import paths resolve to this package; its asset pickles (when the reference tree is
    present, i.e. in the dev container) load into our classes with the reference's private attribute names."""
    import pickle
    import sys
    import humanoid_real_time_retarget_b200 as hrt
    hrt.enable_compat()
    for mod in ("retarget", "poselib", "robot_kinematics_model"):
        assert mod not in sys.modules or hrt.COMPAT_PATH in getattr(sys.modules[mod], "__file__", hrt.COMPAT_PATH)
    import poselib.poselib.core.rotation3d as r3d
    import poselib.poselib.skeleton.skeleton3d as sk3d
    import retarget.spatial_transform.transform3d as t3d
    from retarget.torch_ext import to_torch, to_numpy                     # noqa: F401
    from retarget.retarget_solver import (HuUpperBodyFromMocapRetarget, Mocap2HuBodyRetargeter,   # noqa: F401
                                          VtrdynFullBodyPosRetargeter, VtrdynFullBodyRetargeter)
    from retarget.utils.parse_mocap import vtrdyn_full_zero_pose_transform  # noqa: F401
    from retarget.main import RetargetHuV5fromMocap                       # noqa: F401
    from robot_kinematics_model import RobotZeroPose, cal_forward_kinematics, cal_local_rotation   # noqa: F401
    from robot_kinematics_model.hu_forward_model import HuForwardModel    # noqa: F401
    from retarget.robot_config.Hu import Hu_DOF_AXIS, Hu_DOF_LOWER        # noqa: F401
    assert sk3d.SkeletonState is hrt.SkeletonState and r3d.quat_mul is hrt.rotation3d.quat_mul
    assert t3d.torch is torch and t3d.np is np and len(Hu_DOF_AXIS) == 32 and Hu_DOF_LOWER.shape == (32,)
    from retarget.robot_config.Hu_v5 import VTRDYN2HU_JOINT_MAPPING, SMPL2HU_JOINT_MAPPING
    assert len(VTRDYN2HU_JOINT_MAPPING) == 15 and VTRDYN2HU_JOINT_MAPPING['Hips'] == 'pelvis_link' and len(SMPL2HU_JOINT_MAPPING) == 15
    expected = ["quat_mul", "quat_pos", "quat_abs", "quat_unit", "quat_conjugate", "quat_real", "quat_imaginary", "quat_norm_check",
                "quat_normalize", "quat_from_xyz", "quat_identity", "quat_from_angle_axis", "quat_from_rotation_matrix", "quat_mul_norm",
                "quat_rotate", "quat_inverse", "quat_identity_like", "quat_angle_axis", "quat_yaw_rotation",
                "transform_from_rotation_translation", "transform_identity", "transform_rotation", "transform_translation",
                "transform_inverse", "transform_identity_like", "transform_mul", "transform_apply", "rot_matrix_det",
                "rot_matrix_integrity_check", "rot_matrix_from_quaternion", "euclidean_to_rotation_matrix", "euclidean_integrity_check",
                "euclidean_translation", "euclidean_inverse", "euclidean_to_transform", "project_quat_to_axis_x", "project_quat_to_axis_y",
                "project_quat_to_axis_z", "project_quat_to_axis_xy", "project_quat_to_axis_xz", "extract_rotation_along_axis",
                "quat_mul_four", "quat_mul_three", "normalize_angle", "quat_to_angle_axis", "angle_axis_to_exp_map", "quat_to_exp_map",
                "exp_map_to_angle_axis", "exp_map_to_quat", "quat_to_eular"]
    for name in expected:
        assert callable(getattr(r3d, name)), name
    for name in ["quat_between_two_vecs", "coord_transform", "cal_joint_quat", "quat_in_xyz_axis", "proj_in_plane",
                 "radians_between_vecs", "exp_map_to_quat", "quat_slerp", "quat_to_dof_pos", "quat_mul"]:
        assert callable(getattr(t3d, name)), name
    ref_assets = "/root/reference/asset"
    if not os.path.isdir(ref_assets):
        pytest.skip("reference assets not present (GPU box)")
    from humanoid_real_time_retarget_b200 import robot_config as cfg
    sk = cfg.skeleton_tables()
    for rel, key in [("zero_pose/vtrdyn_zero_pose.pkl", "vtrdyn_zero_pose"), ("hu_pose/hu_v5_zero_pose.pkl", "hu_v5_zero_pose"),
                     ("zero_pose/vtrdyn_full_zero_pose.pkl", "vtrdyn_full_zero_pose"), ("t_pose/vtrdyn_t_pose.pkl", "vtrdyn_t_pose")]:
        with open(os.path.join(ref_assets, rel), "rb") as f:
            st = pickle.load(f)
        assert type(st) is hrt.SkeletonState and type(st.skeleton_tree) is hrt.SkeletonTree
        assert st.skeleton_tree.parent_indices.tolist() == sk[f"{key}/parents"].tolist()
        assert np.array_equal(st.skeleton_tree.local_translation.numpy(), sk[f"{key}/offsets"])
        assert st.tensor.shape[-1] == st.num_joints * 4 + 3 and st.is_local in (True, False)
        assert np.array_equal(st.root_translation.numpy(), sk[f"{key}/root_translation"])


def test_wire_decoder_round_trip_and_whitelist():
    """mocap_receiver.py:49-59 framing: incremental decode of a byte stream cut at arbitrary places; only numpy arrays pass."""
    import pickle
    from humanoid_real_time_retarget_b200 import WireDecoder
    rng = np.random.default_rng(0)
    frames = [{"body_pos": rng.random((23, 3), dtype=np.float32), "body_quat": rng.random((23, 4), dtype=np.float32),
               "left_hand_pos": rng.random((20, 3), dtype=np.float32), "right_hand_pos": rng.random((20, 3), dtype=np.float32)}
              for _ in range(5)]
    stream = b"".join(WireDecoder.encode(f) for f in frames)
    dec, got = WireDecoder(), []
    cuts = [0, 3, 4, 100, 101, 1500, len(stream) - 1, len(stream)]
    for a, b in zip(cuts, cuts[1:]):
        got += dec.feed(stream[a:b])
    assert len(got) == 5 and all(np.array_equal(g[k], f[k]) for g, f in zip(got, frames) for k in f)

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    bad = pickle.dumps({"body_pos": Evil()})
    with pytest.raises(pickle.UnpicklingError):
        WireDecoder().feed(len(bad).to_bytes(4, "big") + bad)
    with pytest.raises(ValueError):
        WireDecoder().feed((1 << 30).to_bytes(4, "big"))
