"""Import shim: retarget/robot_config/Hu.py:4-25 (axes as a list, limits as tensors), served by humanoid_real_time_retarget_b200."""
import torch

from humanoid_real_time_retarget_b200 import robot_config as _cfg

Hu_DOF_AXIS = list(_cfg.Hu_DOF_AXIS)
Hu_DOF_LOWER = torch.Tensor(_cfg.Hu_DOF_LOWER)
Hu_DOF_UPPER = torch.Tensor(_cfg.Hu_DOF_UPPER)
