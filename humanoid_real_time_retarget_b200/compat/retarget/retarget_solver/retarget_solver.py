"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.retarget_solver import *  # noqa: F401,F403
from humanoid_real_time_retarget_b200.retarget_solver import (BaseHumanoidRetargeter, HuUpperBodyFromMocapRetarget,  # noqa: F401
                                                              Mocap2HuBodyRetargeter, VtrdynFullBodyPosRetargeter,
                                                              VtrdynFullBodyRetargeter, cal_elbowP_and_shoulderY, cal_shoulderPR)
