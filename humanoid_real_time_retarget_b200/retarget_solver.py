"""Drop-in for the retarget solvers and their helpers.

  BaseHumanoidRetargeter            retarget/retarget_solver/base_retargeter.py:15-58
  Mocap2HuBodyRetargeter            retarget/retarget_solver/body_retargeter.py:30-99
  VtrdynFullBodyPosRetargeter       retarget/retarget_solver/full_body_pos_retargeter.py:17-217
  HuUpperBodyFromMocapRetarget      retarget/retarget_solver/retarget_solver.py:27-99
  VtrdynFullBodyRetargeter          retarget/retarget_solver/full_body_retargeter.py:15-177
  vtrdyn_zero_pose_transform & co.  retarget/utils/parse_mocap.py:81-89,106-114,126-134

Every solver instance owns its own hrt_ctx, built from the zero poses handed to its constructor (the reference reads
every offset from those objects); nothing here touches a process-wide engine, so instances with different skeletons or
different `precise_gripper` never see each other.
"""
import numpy as np
import torch

from . import robot_config as cfg
from .engine import (BQ_CLAMP, BQ_IK, BQ_PRE_TRANSFORMED, POS_FULL_BODY, POS_FULL_BODY_POS, POS_UPPER_BODY, TREE_ROBOT, TREE_SOURCE,
                     TREE_SOURCE_FULL, Engine, engine_from_zero_poses)
from .kinematics import RobotZeroPose, cal_forward_kinematics


def to_numpy(tensor):
    """retarget/torch_ext.py:10-14."""
    return tensor.cpu().numpy() if torch.is_tensor(tensor) else tensor


def to_torch(tensor):
    """retarget/torch_ext.py:17-21."""
    return tensor if torch.is_tensor(tensor) else torch.from_numpy(tensor).to(torch.float32)


_bundled = {}


def _bundled_engine(device):
    """Private context holding the BUNDLED vtrdyn tables for the module-level transforms below: the reference builds their
    T2Z table at import from its own asset files (parse_mocap.py:65-78,91-104), not from a caller's zero pose."""
    if device not in _bundled:
        _bundled[device] = Engine(device).set_standard_trees()
    return _bundled[device]


def _zpt(global_rotation, tree, variant):
    g = to_torch(global_rotation)
    eng = _bundled_engine(g.device.index or 0 if g.is_cuda else 0)
    return eng.zero_pose_transform(tree, g, variant).reshape(g.shape).to(g.device)


def vtrdyn_zero_pose_transform(global_rotation):
    return _zpt(global_rotation, TREE_SOURCE, 0)


def vtrdyn_full_zero_pose_transform(global_rotation):
    return _zpt(global_rotation, TREE_SOURCE_FULL, 0)


def vtrdyn_broadcast_zero_pose_transform(global_rotation):
    return _zpt(global_rotation, TREE_SOURCE, 1)


def _arm_solve(op, v1, v0, parent_global_rotation):
    from . import rotation3d as r3d
    v1, v0, pq = to_torch(v1), to_torch(v0), to_torch(parent_global_rotation)
    a, b = r3d.run_op(op, [v1, v0, pq], [1, 1, 1], [(4,), (4,)])
    if v1.dim() == 1 and a.dim() == 2 and a.shape[0] == 1:      # the reference squeezes the (1,4) parent's batch axis
        a, b = a[0], b[0]
    return a, b


def cal_shoulderPR(v1, v0, parent_global_rotation):
    """retarget/retarget_solver/retarget_solver.py:127-158 (= full_body_pos_retargeter.py:247-278): shoulder pitch about y and
    roll about x that carry the zero-pose bone v0 onto the measured bone v1 (world frame).  One kernel launch, the same
    device code as the fused position solver; rows broadcast."""
    from . import rotation3d as r3d
    return _arm_solve(r3d.OP_CAL_SHOULDER_PR, v1, v0, parent_global_rotation)


def cal_elbowP_and_shoulderY(v1, v0, parent_global_rotation):
    """retarget_solver.py:103-125 (= full_body_pos_retargeter.py:221-243): shoulder yaw about z and elbow pitch about y."""
    from . import rotation3d as r3d
    return _arm_solve(r3d.OP_CAL_ELBOWP_SHOULDERY, v1, v0, parent_global_rotation)


class BaseHumanoidRetargeter:
    def __init__(self, source_zero_pose: RobotZeroPose, target_zero_pose: RobotZeroPose):
        self.source_zero_pose = source_zero_pose
        self.target_zero_pose = target_zero_pose
        self._motion_local_rotation = []
        self._motion_dof_pos = []

    def _fk(self):
        self._motion_global_rotation, self._motion_global_translation = cal_forward_kinematics(
            motion_local_rotation=self.motion_local_rotation,
            motion_root_translation=torch.zeros((self.motion_length, 3)),
            parent_indices=self.target_zero_pose.parent_indices.tolist(),
            zero_pose_local_translation=self.target_zero_pose.local_translation)

    @property
    def motion_global_rotation(self):
        if not (hasattr(self, '_motion_global_rotation') and len(self._motion_global_rotation) == self.motion_length):
            self._fk()
        return self._motion_global_rotation.clone()

    @property
    def motion_global_translation(self):
        if not (hasattr(self, '_motion_global_translation') and len(self._motion_global_translation) == self.motion_length):
            self._fk()
        return self._motion_global_translation.clone()

    @property
    def motion_local_rotation(self):
        return torch.cat([x.reshape(-1, *x.shape[-2:]) for x in self._motion_local_rotation]).clone()

    @property
    def motion_dof_pos(self):
        return torch.cat([x.reshape(-1, x.shape[-1]) for x in self._motion_dof_pos]).clone()

    @property
    def motion_length(self):
        return sum(x.reshape(-1, *x.shape[-2:]).shape[0] for x in self._motion_local_rotation)


def _require_parents(parents, pairs, what):
    for child, parent in pairs:
        if int(parents[child]) != parent:
            raise ValueError(f"{what}: source joint {child} must hang off joint {parent} (got parent {int(parents[child])}); "
                             "the solver's hard-coded joint indices are those of the vtrdyn skeleton")


class _OwnedEngineRetargeter(BaseHumanoidRetargeter):
    """Builds this instance's own context from the constructor's zero poses."""
    _SOURCE_JOINTS = None

    def __init__(self, mocap_zero_pose, target_zero_pose, device=0):
        super().__init__(mocap_zero_pose, target_zero_pose)
        self._eng, self._src_slot, (self._src_parents, self._src_offsets, self._src_global_t) = engine_from_zero_poses(
            mocap_zero_pose, target_zero_pose, device)
        if self._SOURCE_JOINTS is not None and self._src_parents.shape[0] != self._SOURCE_JOINTS:
            raise ValueError(f"{type(self).__name__} needs a {self._SOURCE_JOINTS}-joint source zero pose, "
                             f"got {self._src_parents.shape[0]} joints")

    def _finish(self, ref, lq, dof, single, record, extra=None):
        lq, dof = lq.to(ref.device), dof.to(ref.device)
        if single:
            lq, dof = lq[0], dof[0]
        if record:
            self._motion_local_rotation.append(lq)
            self._motion_dof_pos.append(dof)
        if extra is None:
            return lq, dof
        extra = extra.to(ref.device)
        return lq, dof, (extra[0] if single else extra)

    def _record(self, lq, dof, record):
        if record:
            self._motion_local_rotation.append(lq)
            self._motion_dof_pos.append(dof)


class Mocap2HuBodyRetargeter(_OwnedEngineRetargeter):
    """retarget_from_pose takes ONE frame (21,4) of zero-pose re-referenced global quats, like the
    reference, or a batch (B,21,4).  Returns new tensors (robot_local_rotation, dof_pos).  The reference reads the
    source zero pose's parent_indices (cal_local_rotation, body_retargeter.py:35) and the target's joint count (:38)."""
    _SOURCE_JOINTS = 21

    def __init__(self, mocap_zero_pose: RobotZeroPose, target_zero_pose: RobotZeroPose, device=0):
        super().__init__(mocap_zero_pose, target_zero_pose, device)
        # local rotations 18, 19 / 14, 15 are taken relative to the source tree's own parents (body_retargeter.py:35,40-53)
        arms = cfg.VTRDYN_ARM_JOINTS
        _require_parents(self._src_parents, [(a[2], a[1]) for a in arms] + [(a[3], a[2]) for a in arms], type(self).__name__)
        self._eng.configure_body_quat(self._src_slot, TREE_ROBOT, arms, cfg.HU_V5_ARM_FIRST)

    def retarget_from_pose(self, source_global_rotation, record=True):
        g = to_torch(source_global_rotation)
        single = g.dim() == 2
        if single and not g.is_cuda and g.dtype == torch.float32:
            # the teleop call: one CPU frame.  It goes through the pinned mailbox of the streaming entry point (one launch,
            # no torch allocation or copy kernels on the way) instead of the batched device path.
            cfg_ = (BQ_PRE_TRANSFORMED, 0, 0.0, 0.0, False)
            if self._eng._bq_stream_cfg != cfg_:
                self._eng.stream_open(flags=BQ_PRE_TRANSFORMED, ik_iters=0, damping=0.0, rot_weight=0.0)
            lq, dof = torch.empty((31, 4)), torch.empty((30,))
            self._eng.stream_frame_tensors(g.contiguous(), lq, dof, None)
            self._record(lq, dof, record)
            return lq, dof
        lq, dof, _ = self._eng.retarget_body_quat(g.reshape(-1, 21, 4), flags=BQ_PRE_TRANSFORMED, want_link_pos=False)
        return self._finish(g, lq, dof, single, record)

    def retarget_clip(self, raw_global_rotation, clamp=True, ik_iters=10, damping=0.1, rot_weight=0.2):
        """Whole pipeline on a (B,21,4) clip of RAW mocap quats: zero-pose transform, mapping, angle
        decomposition, limits, IK refinement, FK.  Returns (robot_local_rotation, dof_pos, link_pos)."""
        g = to_torch(raw_global_rotation)
        flags = (BQ_CLAMP if clamp else 0) | (BQ_IK if ik_iters > 0 else 0)
        out = self._eng.retarget_body_quat(g.reshape(-1, 21, 4), flags=flags, ik_iters=ik_iters, damping=damping,
                                           rot_weight=rot_weight)
        return tuple(o.to(g.device) for o in out)


class VtrdynFullBodyPosRetargeter(_OwnedEngineRetargeter):
    """retarget/retarget_solver/full_body_pos_retargeter.py:17-217.  One frame ((21,3), (20,3), (20,3)) like
    the reference, or a batch with a leading frame axis.  Returns (robot_local_rotation, dof_pos,
    body_global_rotation).  Offsets [11,36,34], [13],[14],[38],[39], [16,20,24,28,32], [41,45,49,53,56] and the global
    translations [18,22,26,30,33],[14] come from `mocap_zero_pose` (:69-107,139,162,184)."""
    _SOURCE_JOINTS = 59

    def __init__(self, mocap_zero_pose, target_zero_pose, precise_gripper=False, device=0, resident=False):
        """resident=True (additive): single CPU frames are served by a resident one-warp kernel polling the pinned mailbox
        (no launch or stream synchronisation per frame, ~28 us instead of ~46 us; it leaves after 20 ms without a frame)."""
        super().__init__(mocap_zero_pose, target_zero_pose, device)
        self.precise_gripper = precise_gripper
        self.resident = bool(resident)
        self._eng.configure_pos(POS_FULL_BODY_POS, self._src_slot, TREE_ROBOT, self._src_global_t, precise_gripper)

    def retarget(self, body_global_translation, left_hand_global_translation, right_hand_global_translation, record=True):
        b = to_torch(body_global_translation)
        single = b.dim() == 2
        if single and not b.is_cuda:
            # the teleop call (sim_full_body_teleop.py:115): one CPU frame through the pinned mailbox of the streaming entry
            # point -- one launch, no torch allocation or copy kernels on the way
            lh, rh = to_torch(left_hand_global_translation), to_torch(right_hand_global_translation)
            if b.dtype == lh.dtype == rh.dtype == torch.float32 and b.shape == (21, 3) and lh.shape == rh.shape == (20, 3):
                cfg_ = (False, self.resident, False, False, True, POS_FULL_BODY_POS)
                if self._eng._pos_stream_cfg != cfg_:
                    self._eng.stream_pos_open(persistent=self.resident, body_gq=True)
                lq, dof, bq = torch.empty((31, 4)), torch.empty((30,)), torch.empty((59, 4))
                self._eng.stream_pos_frame_tensors(b.contiguous(), lh.contiguous(), rh.contiguous(), lq, dof, bq)
                self._record(lq, dof, record)
                return lq, dof, bq
        lq, dof, bq = self._eng.retarget_full_body_pos(b.reshape(-1, 21, 3), to_torch(left_hand_global_translation).reshape(-1, 20, 3),
                                                       to_torch(right_hand_global_translation).reshape(-1, 20, 3))
        return self._finish(b, lq, dof, single, record, bq)


class HuUpperBodyFromMocapRetarget(_OwnedEngineRetargeter):
    """retarget/retarget_solver/retarget_solver.py:27-99: offsets [17,13,11], [19], [15], [20], [16] of the 21-joint
    `mocap_zero_pose` (:50-86)."""
    _SOURCE_JOINTS = 21

    def __init__(self, mocap_zero_pose, target_zero_pose, device=0):
        super().__init__(mocap_zero_pose, target_zero_pose, device)
        self._eng.configure_pos(POS_UPPER_BODY, self._src_slot, TREE_ROBOT)

    def retarget_from_global_translation(self, source_global_translation, record=True):
        b = to_torch(source_global_translation)
        single = b.dim() == 2
        if single and not b.is_cuda and b.dtype == torch.float32 and b.shape == (21, 3):
            # the teleop call (sim_teleop.py:86-108): one CPU frame through the pinned streaming mailbox
            cfg_ = (False, False, False, False, False, POS_UPPER_BODY)
            if self._eng._pos_stream_cfg != cfg_:
                self._eng.stream_pos_open(mode=POS_UPPER_BODY)
            lq, dof = torch.empty((31, 4)), torch.empty((30,))
            self._eng.stream_pos_frame_tensors(b.contiguous(), None, None, lq, dof)
            self._record(lq, dof, record)
            return lq, dof
        lq, dof = self._eng.retarget_upper_body(b.reshape(-1, 21, 3))
        return self._finish(b, lq, dof, single, record)


class VtrdynFullBodyRetargeter(_OwnedEngineRetargeter):
    """retarget/retarget_solver/full_body_retargeter.py:15-177 (the two hand-rotation arguments are unused
    there and here): offsets [13],[14],[38],[39] and the gripper reference [18,22,26,30,33],[24] of the 59-joint
    `mocap_zero_pose` (:59-99,152)."""
    _SOURCE_JOINTS = 59

    def __init__(self, mocap_zero_pose, target_zero_pose, device=0):
        super().__init__(mocap_zero_pose, target_zero_pose, device)
        self._eng.configure_pos(POS_FULL_BODY, self._src_slot, TREE_ROBOT)

    def retarget(self, body_global_rotation, body_global_translation, left_hand_global_rotation,
                 left_hand_global_translation, right_hand_global_rotation, right_hand_global_translation, record=True):
        b = to_torch(body_global_translation)
        single = b.dim() == 2
        if single and not b.is_cuda:
            q, lh, rh = to_torch(body_global_rotation), to_torch(left_hand_global_translation), to_torch(right_hand_global_translation)
            if (b.dtype == q.dtype == lh.dtype == rh.dtype == torch.float32 and b.shape == (21, 3) and q.shape == (21, 4)
                    and lh.shape == rh.shape == (20, 3)):
                cfg_ = (False, False, False, False, False, POS_FULL_BODY)
                if self._eng._pos_stream_cfg != cfg_:
                    self._eng.stream_pos_open(mode=POS_FULL_BODY)
                lq, dof = torch.empty((31, 4)), torch.empty((30,))
                self._eng.stream_pos_frame_tensors(b.contiguous(), lh.contiguous(), rh.contiguous(), lq, dof, None, q.contiguous())
                self._record(lq, dof, record)
                return lq, dof
        lq, dof = self._eng.retarget_full_body(to_torch(body_global_rotation).reshape(-1, 21, 4), b.reshape(-1, 21, 3),
                                               to_torch(left_hand_global_translation).reshape(-1, 20, 3),
                                               to_torch(right_hand_global_translation).reshape(-1, 20, 3))
        return self._finish(b, lq, dof, single, record)
