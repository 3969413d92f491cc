"""Import shim for the UNMODIFIED Python reference at /root/reference.

Dev-container only (the GPU box has no /root/reference).  Used by
tools/make_golden.py to generate tests/golden/*.npz and by ad-hoc probes.
Recipe follows SURVEY.md section 8(c): stub the absent viz/URDF deps, alias
the stale `motion_convert` imports to their in-tree equivalents, chdir into
the reference before importing modules that open('asset/...') at import time.
Nothing under tests/, bench.py or the package imports this file.
"""
import os
import sys
import types
from unittest import mock

REF = os.environ.get("HRT_REFERENCE", "/root/reference")


def _stub(name):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__getattr__ = lambda attr: mock.MagicMock(name=f"{name}.{attr}")
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace holding the reference modules."""
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference not present at {REF}")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in [
        "urdfpy", "trimesh", "trimesh.primitives",
        "poselib.poselib.visualization", "poselib.poselib.visualization.common",
        "vedo_visualizer", "vedo_visualizer.common",
        "body_visualizer", "body_visualizer.common", "body_visualizer.visualizer",
        "motion_convert", "motion_convert.forward_model", "motion_convert.robot_config",
    ]:
        if name not in sys.modules:
            _stub(name)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import robot_kinematics_model
        import robot_kinematics_model.base_forward_model as bfm
        import retarget.robot_config.Hu as hu_cfg
        sys.modules["motion_convert.forward_model.base_forward_model"] = bfm
        sys.modules["motion_convert.robot_config.Hu"] = hu_cfg
        import robot_kinematics_model.hu_forward_model as hfm
        import retarget.spatial_transform.transform3d as t3d
        import poselib.poselib.core.rotation3d as r3d
        import poselib.poselib.skeleton.skeleton3d as sk3d
        import retarget.retarget_solver as solvers
        import retarget.retarget_solver.full_body_pos_retargeter as fbp
        import retarget.utils.parse_mocap as parse_mocap
        import retarget.robot_config.Hu_v5 as hu_v5_cfg
        import retarget.robot_config.VTRDYN as vtrdyn_cfg
        import retarget.robot_config.VTRDYN_FULL as vtrdyn_full_cfg
    finally:
        os.chdir(cwd)
    return types.SimpleNamespace(
        root=REF, rkm=robot_kinematics_model, bfm=bfm, hfm=hfm, t3d=t3d, r3d=r3d,
        sk3d=sk3d, solvers=solvers, fbp=fbp, parse_mocap=parse_mocap,
        hu_cfg=hu_cfg, hu_v5_cfg=hu_v5_cfg, vtrdyn_cfg=vtrdyn_cfg,
        vtrdyn_full_cfg=vtrdyn_full_cfg,
    )


def load_asset(ref, rel):
    import pickle
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with open(rel, "rb") as f:
            return pickle.load(f)
    finally:
        os.chdir(cwd)
