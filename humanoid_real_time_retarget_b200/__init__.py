"""B200-native (sm_100a CUDA behind a C ABI) implementation of the per-frame mocap -> humanoid
retarget hot path of shuoshuof/Humanoid-Real-Time-Retarget.  See DESIGN.md / INTEGRATION.md."""
from . import robot_config
from ._lib import HrtError, LIB_PATH, EXPORTED_SYMBOLS
from .engine import (BQ_CLAMP, BQ_IK, BQ_PRE_TRANSFORMED, FK_EXACT, TREE_ROBOT, TREE_SOURCE, TREE_SOURCE_FULL,
                     Engine, default_engine)
from .kinematics import (BaseForwardModel, HuForwardModel, RobotZeroPose, cal_forward_kinematics, cal_local_rotation)
from .retarget_solver import (BaseHumanoidRetargeter, HuUpperBodyFromMocapRetarget, Mocap2HuBodyRetargeter,
                              VtrdynFullBodyPosRetargeter, VtrdynFullBodyRetargeter, to_numpy, to_torch,
                              vtrdyn_broadcast_zero_pose_transform, vtrdyn_full_zero_pose_transform,
                              vtrdyn_zero_pose_transform)

__all__ = [
    "Engine", "default_engine", "HrtError", "robot_config",
    "cal_forward_kinematics", "cal_local_rotation", "RobotZeroPose", "BaseForwardModel", "HuForwardModel",
    "BaseHumanoidRetargeter", "Mocap2HuBodyRetargeter", "HuUpperBodyFromMocapRetarget", "VtrdynFullBodyRetargeter",
    "VtrdynFullBodyPosRetargeter", "to_numpy", "to_torch",
    "vtrdyn_zero_pose_transform", "vtrdyn_full_zero_pose_transform", "vtrdyn_broadcast_zero_pose_transform",
]
