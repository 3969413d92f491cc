"""Drop-in for robot_kinematics_model/ of the reference: same names, same argument meaning.

  cal_forward_kinematics / cal_local_rotation   robot_kinematics_model/kinematics.py:13,41
  RobotZeroPose                                 robot_kinematics_model/base_robot.py:24-119
  BaseForwardModel / HuForwardModel             base_forward_model.py:7-14, hu_forward_model.py:13-33

Inputs may be CPU or CUDA tensors (the reference works on CPU tensors); results come back on the
device of the input.  All arithmetic runs in the CUDA kernels behind the C ABI.
"""
import copy

import numpy as np
import torch

from . import robot_config as cfg
from .engine import Engine, TREE_ROBOT

_engines = {}


def _tree_engine(parent_indices, offsets, dof_axis=None, lower=None, upper=None, device=0):
    """One context per distinct (tree, tables) signature: the functional API of the reference passes
    the tree on every call, the C ABI installs it once."""
    parents = [int(p) for p in (parent_indices.tolist() if hasattr(parent_indices, "tolist") else parent_indices)]
    off = np.ascontiguousarray(offsets.detach().cpu().numpy() if torch.is_tensor(offsets) else np.asarray(offsets),
                               dtype=np.float32)
    key = (device, tuple(parents), off.tobytes(), None if dof_axis is None else tuple(dof_axis),
           None if lower is None else tuple(float(x) for x in lower), None if upper is None else tuple(float(x) for x in upper))
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(device)
        eng.set_tree(TREE_ROBOT, parents, off, dof_axis, lower, upper)
        _engines[key] = eng
    return eng


def _dev_index(t):
    return t.device.index if t.is_cuda else 0


def cal_forward_kinematics(motion_local_rotation, motion_root_translation, parent_indices, zero_pose_local_translation,
                           exact=False):
    """(L,J,4) local quats + (L,3) root translation -> (L,J,4) global quats, (L,J,3) positions."""
    src_dev = motion_local_rotation.device
    eng = _tree_engine(parent_indices, zero_pose_local_translation, device=_dev_index(motion_local_rotation))
    J = len(parent_indices)
    lead = motion_local_rotation.shape[:-2]
    lq = motion_local_rotation.reshape(-1, J, 4)
    rt = motion_root_translation.reshape(-1, 3).expand(lq.shape[0], 3) if motion_root_translation is not None else None
    gq, gt = eng.fk_local_quats(TREE_ROBOT, lq, rt, exact=exact)
    return gq.reshape(*lead, J, 4).to(src_dev), gt.reshape(*lead, J, 3).to(src_dev)


def cal_local_rotation(motion_global_rotation, parent_indices):
    """(N,J,4) global quats -> (N,J,4) local quats."""
    src_dev = motion_global_rotation.device
    J = len(parent_indices)
    eng = _tree_engine(parent_indices, np.zeros((J, 3), np.float32), device=_dev_index(motion_global_rotation))
    out = eng.local_from_global(TREE_ROBOT, motion_global_rotation.reshape(-1, J, 4))
    return out.reshape(motion_global_rotation.shape).to(src_dev)


class RobotZeroPose:
    """Holder of a zero pose; every tensor property returns a clone, like the reference."""

    def __init__(self, local_translation, global_translation, parent_indices, num_joints, node_names, skeleton_tree=None):
        self._local_translation = torch.as_tensor(local_translation, dtype=torch.float32)
        self._global_translation = torch.as_tensor(global_translation, dtype=torch.float32)
        self._parent_indices = torch.as_tensor(parent_indices).long()
        self._num_joints = int(num_joints)
        self._node_names = list(node_names)
        self._global_rotation = torch.tensor([[0, 0, 0, 1.]] * self._num_joints, dtype=torch.float32)
        self._local_rotation = torch.tensor([[0, 0, 0, 1.]] * self._num_joints, dtype=torch.float32)
        self._skeleton_tree = skeleton_tree

    local_translation = property(lambda self: self._local_translation.clone())
    global_translation = property(lambda self: self._global_translation.clone())
    global_rotation = property(lambda self: self._global_rotation.clone())
    local_rotation = property(lambda self: self._local_rotation.clone())
    parent_indices = property(lambda self: self._parent_indices.clone())
    num_joints = property(lambda self: self._num_joints)
    num_dofs = property(lambda self: self._num_joints - 1)
    node_names = property(lambda self: self._node_names)
    skeleton_tree = property(lambda self: copy.deepcopy(self._skeleton_tree))

    @classmethod
    def from_skeleton_state(cls, skeleton_state):
        tree = skeleton_state.skeleton_tree
        return cls(skeleton_state.local_translation, skeleton_state.global_translation, tree.parent_indices,
                   tree.num_joints, tree.node_names, tree)

    @classmethod
    def from_asset(cls, name):
        """Bundled tables (tools/extract_assets.py), e.g. 'hu_v5_zero_pose', 'vtrdyn_zero_pose'."""
        from .skeleton3d import SkeletonTree
        sk = cfg.skeleton_tables()
        names = [str(s) for s in sk[f"{name}/node_names"]]
        tree = SkeletonTree(names, torch.from_numpy(sk[f"{name}/parents"].astype(np.int64)),
                            torch.from_numpy(sk[f"{name}/offsets"].copy()), torch.from_numpy(sk[f"{name}/tree_quat"].copy()))
        return cls(sk[f"{name}/offsets"], sk[f"{name}/global_translation"], sk[f"{name}/parents"],
                   sk[f"{name}/parents"].shape[0], names, tree)

    @classmethod
    def from_urdf(cls, urdf_path):
        """base_robot.py:68-78 parses a URDF with urdfpy; file readers are out of scope.
        The Hu v5 zero pose that the reference's callers build this way is bundled: from_asset('hu_v5_zero_pose')."""
        if "hu_v5" in str(urdf_path):
            return cls.from_asset("hu_v5_zero_pose")
        raise NotImplementedError("URDF parsing is outside the retarget hot path; use from_asset / from_skeleton_state")

    @classmethod
    def from_dict(cls, robot_dict, is_local=False):
        """base_robot.py:90-96"""
        if is_local:
            robot_dict['global_translation'] = cls.cal_global_translation(robot_dict['local_translation'], robot_dict['parent_indices'])
        else:
            robot_dict['local_translation'] = cls.cal_local_translation(robot_dict['global_translation'], robot_dict['parent_indices'])
        return cls(**robot_dict)

    @staticmethod
    def cal_global_translation(local_translation, parent_indices):
        raise NotImplementedError          # as in the reference (base_robot.py:103-104)

    def get_sk_zero_pose(self):
        from .skeleton3d import SkeletonState
        return SkeletonState.zero_pose(self.skeleton_tree)

    @staticmethod
    def cal_local_translation(global_translation, parent_indices):
        local_translation = global_translation.clone()
        local_translation[1:] -= global_translation[parent_indices[1:]]
        return local_translation

    def rebuild_pose_by_local_rotation(self, local_rotation):
        global_rotation, gt = cal_forward_kinematics(local_rotation, self.global_translation[0], self.parent_indices.tolist(),
                                                     self.local_translation, exact=True)
        self._global_translation = gt
        self._local_translation = self.cal_local_translation(self.global_translation, self.parent_indices)
        if self._skeleton_tree is not None:
            self._skeleton_tree._local_translation = self.local_translation
        return global_rotation


class BaseForwardModel:
    def __init__(self, skeleton_tree, device='cuda:0'):
        self.sk_local_translation = skeleton_tree.local_translation
        self.parent_indices = skeleton_tree.parent_indices
        self.num_joints = int(skeleton_tree.num_joints)
        self.device = device

    def forward_kinematics(self, **kwargs):
        return cal_forward_kinematics(**kwargs, parent_indices=self.parent_indices,
                                      zero_pose_local_translation=self.sk_local_translation)


class _HuFKFunction(torch.autograd.Function):
    """forward = hrt_fk_angles, backward = hrt_fk_vjp (the Jacobian is contracted on the device, never materialised).
    The clamp is straight-through like the reference's `(clamp(x) - x).detach() + x` (hu_forward_model.py:27-33): the
    backward is evaluated at the clipped angles and hands the gradient to the raw angles unchanged."""

    @staticmethod
    def forward(ctx, angles, root_t, root_q, eng, clip, exact):
        L = angles.shape[0]
        gq, gt = eng.fk_angles(TREE_ROBOT, angles.reshape(L, -1), root_t, root_q.reshape(L, 4), clip=clip, exact=exact)
        ctx.save_for_backward(angles, root_t, root_q)
        ctx.eng, ctx.clip = eng, clip
        return gq.to(angles.device), gt.to(angles.device)

    @staticmethod
    def backward(ctx, g_gq, g_gt):
        angles, root_t, root_q = ctx.saved_tensors
        L = angles.shape[0]
        want_root = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        ga, grt, grq = ctx.eng.fk_vjp(TREE_ROBOT, angles.reshape(L, -1), g_gq, g_gt, root_t, root_q.reshape(L, 4), clip=ctx.clip,
                                      want_root=want_root)
        ga = ga.reshape(angles.shape).to(angles.device) if ctx.needs_input_grad[0] else None
        grt = grt.reshape(root_t.shape).to(root_t.device) if ctx.needs_input_grad[1] else None
        grq = grq.reshape(root_q.shape).to(root_q.device) if ctx.needs_input_grad[2] else None
        return ga, grt, grq, None, None, None


class HuForwardModel(BaseForwardModel):
    """Angles -> FK with the straight-through joint-limit clamp's forward value.  The reference wires
    this class to the 33-joint Hu tables (hu_forward_model.py:9); a 31-joint tree selects Hu v5."""

    def __init__(self, skeleton_tree, device='cuda:0'):
        super().__init__(skeleton_tree, device)
        if self.num_joints == 33:
            self._tables = (cfg.Hu_DOF_AXIS, cfg.Hu_DOF_LOWER, cfg.Hu_DOF_UPPER)
        elif self.num_joints == 31:
            self._tables = (cfg.Hu_v5_DOF_AXIS, cfg.Hu_v5_DOF_LOWER, cfg.Hu_v5_DOF_UPPER)
        else:
            raise ValueError("HuForwardModel needs the 33-joint Hu or the 31-joint Hu v5 tree")
        self.joint_rotation_axis = torch.eye(3)[self._tables[0]]
        dev = torch.device(device)
        self._eng = _tree_engine(self.parent_indices, self.sk_local_translation, *self._tables,
                                 device=dev.index or 0)

    def forward_kinematics(self, motion_joint_angles, motion_root_translation, motion_root_rotation, clip_angles,
                           exact=False):
        L = motion_joint_angles.shape[0]
        src_dev = motion_joint_angles.device
        if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad
                                           for t in (motion_joint_angles, motion_root_translation, motion_root_rotation)):
            # optimisation-based callers: autograd flows through FK (hu_forward_model.py:27-33, SURVEY section 4 invariant 4)
            rt = motion_root_translation.to(torch.float32).expand(L, 3)
            return _HuFKFunction.apply(motion_joint_angles.to(torch.float32), rt, motion_root_rotation.to(torch.float32),
                                       self._eng, bool(clip_angles), bool(exact))
        gq, gt = self._eng.fk_angles(TREE_ROBOT, motion_joint_angles.reshape(L, -1), motion_root_translation,
                                     motion_root_rotation.reshape(L, 4), clip=clip_angles, exact=exact)
        return gq.to(src_dev), gt.to(src_dev)

    def jacobian(self, motion_joint_angles, motion_root_translation, motion_root_rotation, clip_angles, links):
        """Geometric Jacobian (B, K, 6, D) of the requested links (ours)."""
        L = motion_joint_angles.shape[0]
        return self._eng.fk_jacobian(TREE_ROBOT, motion_joint_angles.reshape(L, -1), links, motion_root_translation,
                                     None if motion_root_rotation is None else motion_root_rotation.reshape(L, 4),
                                     clip=clip_angles)
