"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from .tensor_utils import *  # noqa: F401,F403
from .rotation3d import *  # noqa: F401,F403
from humanoid_real_time_retarget_b200.skeleton3d import Serializable  # noqa: F401
