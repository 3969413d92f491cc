"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.retarget_solver import (vtrdyn_broadcast_zero_pose_transform,  # noqa: F401
                                                              vtrdyn_full_zero_pose_transform, vtrdyn_zero_pose_transform)
