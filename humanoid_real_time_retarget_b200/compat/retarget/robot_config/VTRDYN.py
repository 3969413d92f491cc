"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.robot_config import (VTRDYN_CONNECTIONS, VTRDYN_JOINT_NAMES,  # noqa: F401
                                                           vtrdyn_parent_indices)
