"""Static tables of the retarget path, restated from the reference's robot_config (bit-exact;
tests/test_tables.py compares them with values extracted from the reference).

  Hu / Hu v5 per-DOF hinge axes and limits   retarget/robot_config/Hu.py:4-25, Hu_v5.py:12-33
  vtrdyn parents and joint names             retarget/robot_config/VTRDYN.py:2-48
  wire -> solver index remaps                sim_full_body_teleop.py:109-112,
                                             retarget/retarget_solver/full_body_pos_retargeter.py:320-323
"""
import os

import numpy as np

Hu_v5_DOF_AXIS = [
    2, 0, 1, 1, 1,
    2, 0, 1, 1, 1,
    2,
    1, 0, 2, 1, 0, 1, 2, 1, 1,
    1, 0, 2, 1, 0, 1, 2, 1, 1,
    2]

Hu_DOF_AXIS = [
    2, 0, 1, 1, 1, 0,
    2, 0, 1, 1, 1, 0,
    2,
    1, 0, 2, 1, 0, 1, 2, 1, 1,
    1, 0, 2, 1, 0, 1, 2, 1, 1,
    2]

Hu_DOF_LOWER = [
    -0.1745, -0.3491, -1.5708, 0.0997, -0.6981, -0.3665,
    -0.1745, -0.3491, -1.5708, 0.0997, -0.6981, -0.3665,
    -1.0472,
    -3.1416, 0., -1.5708, 0., -1.5708, -0.785, -0.7854, 0., -0.044,
    -3.1416, -1.5708, -1.5708, 0., -1.5708, -0.785, -0.7854, 0., -0.044,
    -1.]

Hu_DOF_UPPER = [
    0.1745, 0.3491, 0.8727, 2.618, 0.6981, 0.3665,
    0.1745, 0.3491, 0.8727, 2.618, 0.6981, 0.3665,
    1.0472,
    1.0472, 1.5708, 1.5708, 1.5708, 1.5708, 0.785, 0.7854, 0.044, 0.,
    1.0472, 0., 1.5708, 1.5708, 1.5708, 0.785, 0.7854, 0.044, 0.,
    1.]

# The reference ships no consistent 30-entry limit table for Hu v5 (Hu_v5.py:20-33 repeats the
# 32-entry Hu table and nothing reads it).  Ours: the Hu table without the two toe DOFs (5, 11).
Hu_v5_DOF_LOWER = [v for i, v in enumerate(Hu_DOF_LOWER) if i not in (5, 11)]
Hu_v5_DOF_UPPER = [v for i, v in enumerate(Hu_DOF_UPPER) if i not in (5, 11)]

VTRDYN_JOINT_NAMES = [
    'Hips', 'RightUpperLeg', 'RightLowerLeg', 'RightFoot', 'LeftUpperLeg', 'LeftLowerLeg', 'LeftFoot',
    'Spine', 'Spine1', 'Spine2', 'Spine3', 'Neck', 'Head',
    'RightShoulder', 'RightUpperArm', 'RightLowerArm', 'RightHand',
    'LeftShoulder', 'LeftUpperArm', 'LeftLowerArm', 'LeftHand']

VTRDYN_CONNECTIONS = [(0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6),
                      (0, 7), (7, 8), (8, 9), (9, 10), (10, 11), (11, 12),
                      (10, 13), (13, 14), (14, 15), (15, 16),
                      (10, 17), (17, 18), (18, 19), (19, 20)]
vtrdyn_parent_indices = [-1] + [c[0] for c in VTRDYN_CONNECTIONS]

# fused quaternion path wiring (body_retargeter.py:40-73): {torso, shoulder, upper, lower, hand}
VTRDYN_ARM_JOINTS = [[10, 17, 18, 19, 20], [10, 13, 14, 15, 16]]   # left, right
HU_V5_ARM_FIRST = [12, 21]                                          # shoulder-pitch links

# wire layout -> solver layout
BODY_23_TO_21 = [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]
HAND_20_REORDER = [0, 4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 18, 19, 12, 13, 14, 15, 1, 2, 3]
FULL59_TO_BODY21 = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "skeletons.npz")
_cache = None


def skeleton_tables():
    """Numeric content of the reference's pickled assets (tools/extract_assets.py)."""
    global _cache
    if _cache is None:
        z = np.load(_DATA, allow_pickle=False)
        _cache = {k: z[k] for k in z.files}
    return _cache


def joint_mappings(robot="Hu_v5"):
    """The reference's source-joint -> robot-link name tables (retarget/robot_config/Hu.py:27-105, Hu_v5.py:35-113),
    extracted verbatim by tools/extract_tables.py: {'SMPL2HU_JOINT_MAPPING': {...}, 'NOITOM2HU_...', 'VTRDYN2HU_...',
    'VTRDYN_LITE2HU_...'}; the argument of SkeletonState.retarget_to."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "joint_mappings.json")) as f:
        return json.load(f)[robot]
