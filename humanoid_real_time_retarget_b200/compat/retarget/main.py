"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.retarget_main import Retarget, RetargetHuV5fromMocap  # noqa: F401
