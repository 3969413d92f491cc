"""Import shim: retarget/robot_config/Hu_v5.py:12-33 (30 axes; the limits there repeat the 32-entry Hu table)."""
import torch

from humanoid_real_time_retarget_b200 import robot_config as _cfg

Hu_DOF_AXIS = list(_cfg.Hu_v5_DOF_AXIS)
Hu_DOF_LOWER = torch.Tensor(_cfg.Hu_DOF_LOWER)
Hu_DOF_UPPER = torch.Tensor(_cfg.Hu_DOF_UPPER)

_maps = _cfg.joint_mappings("Hu_v5")
SMPL2HU_JOINT_MAPPING = _maps["SMPL2HU_JOINT_MAPPING"]
NOITOM2HU_JOINT_MAPPING = _maps["NOITOM2HU_JOINT_MAPPING"]
VTRDYN2HU_JOINT_MAPPING = _maps["VTRDYN2HU_JOINT_MAPPING"]
VTRDYN_LITE2HU_JOINT_MAPPING = _maps["VTRDYN_LITE2HU_JOINT_MAPPING"]
