"""Import shim: the reference module path, served by humanoid_real_time_retarget_b200 (see enable_compat)."""
from humanoid_real_time_retarget_b200.retarget_solver import (HuUpperBodyFromMocapRetarget, Mocap2HuBodyRetargeter,  # noqa: F401
                                                              VtrdynFullBodyPosRetargeter, VtrdynFullBodyRetargeter)

__all__ = ['HuUpperBodyFromMocapRetarget', 'Mocap2HuBodyRetargeter', 'VtrdynFullBodyRetargeter', 'VtrdynFullBodyPosRetargeter']
