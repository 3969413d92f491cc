"""B200-native (sm_100a CUDA behind a C ABI) implementation of the per-frame mocap -> humanoid
retarget hot path of shuoshuof/Humanoid-Real-Time-Retarget.  See DESIGN.md / INTEGRATION.md."""
from . import robot_config
from ._lib import HrtError, LIB_PATH, EXPORTED_SYMBOLS
from .engine import (BQ_ACTIVE_SET, BQ_CLAMP, BQ_IK, BQ_PACKED_IK, BQ_PRE_TRANSFORMED, POS_CLAMP, POS_IK, POS_FULL_BODY, POS_FULL_BODY_POS,
                     POS_MAIN, POS_UPPER_BODY, FK_EXACT, TREE_ROBOT, TREE_SOURCE, TREE_SOURCE_FULL,
                     Engine, default_engine)
from .kinematics import (BaseForwardModel, HuForwardModel, RobotZeroPose, cal_forward_kinematics, cal_local_rotation)
from .retarget_solver import (cal_elbowP_and_shoulderY, cal_shoulderPR, BaseHumanoidRetargeter, HuUpperBodyFromMocapRetarget, Mocap2HuBodyRetargeter,
                              VtrdynFullBodyPosRetargeter, VtrdynFullBodyRetargeter, to_numpy, to_torch,
                              vtrdyn_broadcast_zero_pose_transform, vtrdyn_full_zero_pose_transform,
                              vtrdyn_zero_pose_transform)

from . import rotation3d, transform3d, skeleton3d
from .skeleton3d import MotionDICT, SkeletonMotion, SkeletonState, SkeletonTree
from .retarget_main import Retarget, RetargetHuV5fromMocap
from .teleop import TeleopSession, WireDecoder

import os as _os
COMPAT_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "compat")


def enable_compat():
    """Put the import shims on sys.path: `poselib.poselib.core.rotation3d`, `poselib.poselib.skeleton.skeleton3d`,
    `retarget.spatial_transform.transform3d`, `retarget.torch_ext`, `retarget.retarget_solver`, `retarget.utils.parse_mocap`,
    `retarget.main`, `robot_kinematics_model` then resolve to this package, so the reference's callers
    (sim_full_body_teleop.py:15-23 ...) and its asset pickles work unchanged."""
    import sys
    if COMPAT_PATH not in sys.path:
        sys.path.insert(0, COMPAT_PATH)


__all__ = [
    "rotation3d", "transform3d", "skeleton3d", "SkeletonTree", "SkeletonState", "SkeletonMotion", "MotionDICT",
    "Retarget", "RetargetHuV5fromMocap", "TeleopSession", "WireDecoder", "enable_compat", "COMPAT_PATH",
    "Engine", "default_engine", "HrtError", "robot_config",
    "cal_forward_kinematics", "cal_local_rotation", "RobotZeroPose", "BaseForwardModel", "HuForwardModel",
    "BaseHumanoidRetargeter", "Mocap2HuBodyRetargeter", "HuUpperBodyFromMocapRetarget", "VtrdynFullBodyRetargeter",
    "VtrdynFullBodyPosRetargeter", "to_numpy", "to_torch",
    "vtrdyn_zero_pose_transform", "vtrdyn_full_zero_pose_transform", "vtrdyn_broadcast_zero_pose_transform",
]
