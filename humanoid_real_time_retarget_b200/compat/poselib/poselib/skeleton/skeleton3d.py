"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.skeleton3d import (MotionDICT, Serializable, SkeletonMotion, SkeletonState,  # noqa: F401
                                                         SkeletonTree, tensor_to_dict)
