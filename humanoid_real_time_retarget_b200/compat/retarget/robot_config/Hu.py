"""Import shim: retarget/robot_config/Hu.py:4-25 (axes as a list, limits as tensors), served by humanoid_real_time_retarget_b200."""
import torch

from humanoid_real_time_retarget_b200 import robot_config as _cfg

Hu_DOF_AXIS = list(_cfg.Hu_DOF_AXIS)
Hu_DOF_LOWER = torch.Tensor(_cfg.Hu_DOF_LOWER)
Hu_DOF_UPPER = torch.Tensor(_cfg.Hu_DOF_UPPER)

_maps = _cfg.joint_mappings("Hu")
SMPL2HU_JOINT_MAPPING = _maps["SMPL2HU_JOINT_MAPPING"]
NOITOM2HU_JOINT_MAPPING = _maps["NOITOM2HU_JOINT_MAPPING"]
VTRDYN2HU_JOINT_MAPPING = _maps["VTRDYN2HU_JOINT_MAPPING"]
VTRDYN_LITE2HU_JOINT_MAPPING = _maps["VTRDYN_LITE2HU_JOINT_MAPPING"]
