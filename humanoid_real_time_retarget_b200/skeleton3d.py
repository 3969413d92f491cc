"""Drop-in for poselib/poselib/skeleton/skeleton3d.py: SkeletonTree / SkeletonState / SkeletonMotion /
MotionDICT with the reference's attribute names (so the bundled asset pickles load into these classes
through the `compat/` import shims) and the reference's lazily cached properties -- but every rotation
product, forward-kinematics pass, global->local conversion and velocity estimate runs in the sm_100a
kernels behind the C ABI (fk_limb_kernel / local_from_global_kernel / rot_op_kernel / the motion kernels).

Reference lines: SkeletonTree skeleton3d.py:22-263, SkeletonState :266-934, SkeletonMotion :937-1292,
MotionDICT :1295-1314.  File/FBX readers stay out of scope."""
import copy
import json
import os
import xml.etree.ElementTree as ET
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import rotation3d as r3d
from .kinematics import _tree_engine, cal_forward_kinematics, cal_local_rotation
from .engine import TREE_ROBOT


# ---------------------------------------------------------------------------------- (de)serialisation
def tensor_to_dict(x):
    """poselib/poselib/core/tensor_utils.py:34-46"""
    a = x.detach().cpu().numpy()
    return {"arr": a, "context": {"dtype": a.dtype.name}}


def _tensor_from_dict(d):
    return torch.from_numpy(np.asarray(d["arr"]).astype(d["context"]["dtype"]))


class Serializable:
    """poselib/poselib/core/backend/abstract.py: to_file / from_file on top of to_dict / from_dict (.npy, .json)."""

    @classmethod
    def from_file(cls, path, *args, **kwargs):
        if path.endswith(".npy"):
            d = np.load(path, allow_pickle=True).item()
        elif path.endswith(".json"):
            def hook(o):
                if isinstance(o, dict) and "__ndarray__" in o:
                    return np.asarray(o["__ndarray__"], dtype=o["dtype"]).reshape(o["shape"])
                return o
            with open(path) as f:
                d = json.load(f, object_hook=hook)
        else:
            raise AssertionError("failed to load {} from {}".format(cls.__name__, path))
        assert d["__name__"] == cls.__name__, "the file belongs to {}, not {}".format(d["__name__"], cls.__name__)
        return cls.from_dict(d, *args, **kwargs)

    def to_file(self, path: str) -> None:
        folder = os.path.dirname(path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        d = self.to_dict()
        d["__name__"] = self.__class__.__name__
        if path.endswith(".npy"):
            np.save(path, d)
        elif path.endswith(".json"):
            def enc(o):
                if isinstance(o, np.ndarray):
                    return dict(__ndarray__=o.tolist(), dtype=str(o.dtype), shape=o.shape)
                if isinstance(o, np.generic):
                    return o.item()
                raise TypeError(type(o))
            with open(path, "w") as f:
                json.dump(d, f, default=enc, indent=4)


# ---------------------------------------------------------------------------------- SkeletonTree
class SkeletonTree(Serializable):
    """Names, parent indices (int64, -1 = root), zero-pose local translations and a per-node rest rotation
    `quat` (identity by default).  Pickled attributes: _node_names, _parent_indices, _local_translation,
    _quat, _node_indices (skeleton3d.py:89-95)."""

    def __init__(self, node_names, parent_indices, local_translation, quat=None):
        assert len({len(node_names), len(parent_indices), len(local_translation)}) == 1
        self._node_names = node_names
        self._parent_indices = parent_indices.long()
        self._local_translation = local_translation
        self._quat = r3d.quat_identity([len(node_names)]) if quat is None else quat
        self._node_indices = {name: i for i, name in enumerate(node_names)}

    def __len__(self):
        return len(self._node_names)

    def __iter__(self):
        yield from self._node_names

    def __getitem__(self, item):
        return self._node_names[item]

    def __repr__(self):
        pad = lambda s: "\n    ".join(s.split("\n"))
        return ("SkeletonTree(\n    node_names={},\n    parent_indices={},\n    local_translation={}\n)".format(
            pad(repr(self.node_names)), pad(repr(self.parent_indices)), pad(repr(self.local_translation))))

    node_names = property(lambda self: self._node_names)
    parent_indices = property(lambda self: self._parent_indices)
    local_translation = property(lambda self: self._local_translation)
    quat = property(lambda self: self._quat)
    num_joints = property(lambda self: len(self))

    @classmethod
    def from_dict(cls, dict_repr, *args, **kwargs):
        return cls(list(map(str, dict_repr["node_names"])), _tensor_from_dict(dict_repr["parent_indices"]),
                   _tensor_from_dict(dict_repr["local_translation"]),
                   _tensor_from_dict(dict_repr["quat"]) if "quat" in dict_repr else None)

    def to_dict(self):
        return OrderedDict([("node_names", self.node_names), ("parent_indices", tensor_to_dict(self.parent_indices)),
                            ("local_translation", tensor_to_dict(self.local_translation))])

    @classmethod
    def from_mjcf(cls, path: str, load_quat: bool = False) -> "SkeletonTree":
        """skeleton3d.py:157-206: depth-first walk of <worldbody><body>...; MJCF quats are wxyz."""
        world = ET.parse(path).getroot().find("worldbody")
        root = None if world is None else world.find("body")
        if root is None:
            raise ValueError("MJCF parsed incorrectly please verify it.")
        names, parents, offsets, quats = [], [], [], []

        def visit(node, parent):
            me = len(names)
            names.append(node.attrib.get("name"))
            parents.append(parent)
            offsets.append(np.array(node.attrib.get("pos").split(), dtype=float))
            quats.append(np.array(node.attrib.get("quat", "1. 0. 0. 0.").split(), dtype=float))
            for child in node.findall("body"):
                visit(child, me)

        visit(root, -1)
        q = torch.from_numpy(np.array(quats, dtype=np.float32))[:, [1, 2, 3, 0]] if load_quat else None
        return cls(names, torch.from_numpy(np.array(parents, dtype=np.int32)),
                   torch.from_numpy(np.array(offsets, dtype=np.float32)), q)

    def parent_of(self, node_name):
        return self[int(self.parent_indices[self.index(node_name)].item())]

    def index(self, node_name):
        return self._node_indices[node_name]

    def drop_nodes_by_names(self, node_names: List[str], pairwise_translation=None) -> "SkeletonTree":
        """skeleton3d.py:226-259: a kept node re-attaches to its nearest kept ancestor; its offset is the sum
        of the dropped offsets on the way (accumulated IN PLACE on this tree's row, like the reference) or the
        supplied pairwise translation."""
        drop = set(node_names)
        parents = self.parent_indices.numpy()
        kept_names, kept_off, kept_par, new_index = [], [], [], {}
        for j in range(len(self)):
            if self[j] in drop:
                continue
            anc = parents[j]
            off = self.local_translation[j, :]
            if anc != -1:
                while anc != -1 and self[anc] in drop:
                    off += self.local_translation[anc, :]
                    anc = parents[anc]
                assert anc != -1, "the root node cannot be dropped"
                if pairwise_translation is not None:
                    off = pairwise_translation[anc, j, :]
            new_index[self[j]] = len(kept_names)
            kept_names.append(self[j])
            kept_off.append(off.clone())
            kept_par.append(-1 if anc == -1 else new_index[self[anc]])
        off = torch.stack(kept_off).to(self.local_translation.dtype) if kept_off else torch.zeros(0, 3)
        return SkeletonTree(kept_names, torch.tensor(kept_par, dtype=self.parent_indices.dtype), off)

    def keep_nodes_by_names(self, node_names: List[str], pairwise_translation=None) -> "SkeletonTree":
        return self.drop_nodes_by_names([n for n in self if n not in node_names], pairwise_translation)


# ---------------------------------------------------------------------------------- SkeletonState
class SkeletonState(Serializable):
    """`tensor` = [J*4 rotations | 3 root translation] with any leading dims (skeleton3d.py:336-339).  The
    derived quantities are computed on first use by the CUDA kernels and cached under the reference's
    private attribute names."""

    def __init__(self, tensor_backend, skeleton_tree, is_local):
        self._skeleton_tree = skeleton_tree
        self._is_local = is_local
        self.tensor = tensor_backend.clone()

    def __len__(self):
        return self.tensor.shape[0]

    # ---- raw views
    @property
    def rotation(self):
        if not hasattr(self, "_rotation"):
            J = self.num_joints
            self._rotation = self.tensor[..., :J * 4].reshape(*self.tensor.shape[:-1], J, 4)
        return self._rotation

    _local_rotation = property(lambda self: self.rotation if self._is_local else None)
    _global_rotation = property(lambda self: None if self._is_local else self.rotation)
    is_local = property(lambda self: self._is_local)
    invariant_property = property(lambda self: {"skeleton_tree": self.skeleton_tree, "is_local": self.is_local})
    num_joints = property(lambda self: self.skeleton_tree.num_joints)
    skeleton_tree = property(lambda self: self._skeleton_tree)

    @property
    def root_translation(self):
        if not hasattr(self, "_root_translation"):
            J = self.num_joints
            self._root_translation = self.tensor[..., J * 4:J * 4 + 3]
        return self._root_translation

    # ---- forward kinematics (skeleton3d.py:402-425)
    @property
    def global_transformation(self):
        if not hasattr(self, "_global_transformation"):
            tree = self.skeleton_tree
            lq = self.local_rotation
            # every non-root local rotation is pre-multiplied by the node's rest rotation tree.quat[j]
            pre = r3d.quat_mul_norm(tree.quat.to(lq.device), lq)
            pre[..., 0, :] = lq[..., 0, :]
            lead = lq.shape[:-2]
            root = self.root_translation.reshape(-1, 3).expand(max(int(np.prod(lead)), 1), 3) if len(lead) else self.root_translation.reshape(1, 3)
            gq, gt = cal_forward_kinematics(pre.reshape(-1, len(tree), 4), root, tree.parent_indices.tolist(),
                                            tree.local_translation, exact=True)
            self._global_transformation = torch.cat([gq.reshape(*lead, len(tree), 4), gt.reshape(*lead, len(tree), 3)], dim=-1)
        return self._global_transformation

    @property
    def global_rotation(self):
        if self._global_rotation is not None:
            return self._global_rotation
        if not hasattr(self, "_comp_global_rotation"):
            self._comp_global_rotation = r3d.transform_rotation(self.global_transformation)
        return self._comp_global_rotation

    @property
    def global_translation(self):
        if not hasattr(self, "_global_translation"):
            self._global_translation = r3d.transform_translation(self.global_transformation)
        return self._global_translation

    @property
    def global_translation_xy(self):
        out = torch.zeros_like(self.global_translation)
        out[..., 0:2] = self.global_translation[..., 0:2]
        return out

    @property
    def global_translation_xz(self):
        out = torch.zeros_like(self.global_translation)
        out[..., 0:1] = self.global_translation[..., 0:1]
        out[..., 2:3] = self.global_translation[..., 2:3]
        return out

    # ---- global -> local (skeleton3d.py:460-484)
    @property
    def local_rotation(self):
        if self._local_rotation is not None:
            return self._local_rotation
        if not hasattr(self, "_comp_local_rotation"):
            tree = self.skeleton_tree
            g = self.global_rotation
            l = cal_local_rotation(g.reshape(-1, len(tree), 4), tree.parent_indices.tolist()).reshape(g.shape)
            inv_rest = r3d.quat_normalize(r3d.quat_inverse(tree.quat.to(g.device)))
            out = r3d.quat_mul_norm(inv_rest, l)
            out[..., 0, :] = g[..., 0, :]
            self._comp_local_rotation = out
        return self._comp_local_rotation

    @property
    def local_transformation(self):
        if not hasattr(self, "_local_transformation"):
            self._local_transformation = r3d.transform_from_rotation_translation(r=self.local_rotation, t=self.local_translation)
        return self._local_transformation

    @property
    def local_translation(self):
        if not hasattr(self, "_local_translation"):
            tree = self.skeleton_tree
            shape = tuple(self.tensor.shape[:-1]) + (len(tree), 3)
            lt = tree.local_translation.broadcast_to(*shape).clone()
            lt[..., 0, :] = self.root_translation
            self._local_translation = lt
        return self._local_translation

    # ---- root-relative quantities
    @property
    def root_translation_xy(self):
        if not hasattr(self, "_root_translation_xy"):
            self._root_translation_xy = self.global_translation_xy[..., 0, :]
        return self._root_translation_xy

    @property
    def global_root_rotation(self):
        if not hasattr(self, "_global_root_rotation"):
            self._global_root_rotation = self.global_rotation[..., 0, :]
        return self._global_root_rotation

    @property
    def global_root_yaw_rotation(self):
        if not hasattr(self, "_global_root_yaw_rotation"):
            self._global_root_yaw_rotation = r3d.quat_yaw_rotation(self.global_root_rotation)
        return self._global_root_yaw_rotation

    @property
    def local_translation_to_root(self):
        if not hasattr(self, "_local_translation_to_root"):
            self._local_translation_to_root = self.global_translation - self.root_translation.unsqueeze(-2)
        return self._local_translation_to_root

    @property
    def local_rotation_to_root(self):
        return r3d.quat_mul(r3d.quat_inverse(self.global_root_rotation).unsqueeze(-2), self.global_rotation)

    def compute_forward_vector(self, left_shoulder_index, right_shoulder_index, left_hip_index, right_hip_index,
                               gaussian_filter_width=20):
        """skeleton3d.py:542-566: up x (average of the right->left shoulder and hip vectors), smoothed along frames with a
        gaussian of the given width and normalised; (T,3) float64 like the reference (numpy promotes the cross product)."""
        gt = self.global_translation
        eng = r3d._engine(gt.device)
        out = eng.forward_vector(gt, left_shoulder_index, right_shoulder_index, left_hip_index, right_hip_index, gaussian_filter_width)
        return out.to(gt.device)

    # ---- construction
    @staticmethod
    def _to_state_vector(rot, rt):
        lead = rot.shape[:-2]
        return torch.cat([rot.reshape(*lead, -1), rt.broadcast_to(*lead, rt.shape[-1]).reshape(*lead, -1)], dim=-1)

    @classmethod
    def from_dict(cls, dict_repr, *args, **kwargs):
        return cls(SkeletonState._to_state_vector(_tensor_from_dict(dict_repr["rotation"]), _tensor_from_dict(dict_repr["root_translation"])),
                   SkeletonTree.from_dict(dict_repr["skeleton_tree"]), dict_repr["is_local"])

    def to_dict(self):
        return OrderedDict([("rotation", tensor_to_dict(self.rotation)), ("root_translation", tensor_to_dict(self.root_translation)),
                            ("skeleton_tree", self.skeleton_tree.to_dict()), ("is_local", self.is_local)])

    @classmethod
    def from_rotation_and_root_translation(cls, skeleton_tree, r, t, is_local=True):
        """skeleton3d.py:594-617: rotations are normalised (quat_normalize) on the way in."""
        assert r.dim() > 0, "the rotation needs to have at least 1 dimension (dim = {})".format(r.dim)
        return cls(SkeletonState._to_state_vector(r3d.quat_normalize(r), torch.as_tensor(t).to(r.device)),
                   skeleton_tree=skeleton_tree, is_local=is_local)

    @classmethod
    def zero_pose(cls, skeleton_tree):
        """skeleton3d.py:619-634: local rotations = tree.quat, root at the origin."""
        return cls.from_rotation_and_root_translation(skeleton_tree, skeleton_tree.quat,
                                                      torch.zeros(3, dtype=skeleton_tree.local_translation.dtype), is_local=True)

    def local_repr(self):
        if self.is_local:
            return self
        return SkeletonState.from_rotation_and_root_translation(self.skeleton_tree, self.local_rotation, self.root_translation, is_local=True)

    def global_repr(self):
        if not self.is_local:
            return self
        return SkeletonState.from_rotation_and_root_translation(self.skeleton_tree, self.global_rotation, self.root_translation, is_local=False)

    # ---- node dropping / remapping (skeleton3d.py:668-740)
    def _get_pairwise_average_translation(self):
        T = self.global_transformation
        inv = r3d.transform_inverse(T)
        J = len(self.skeleton_tree)
        pair = r3d.transform_mul(inv.unsqueeze(-2), T.unsqueeze(-3))
        return r3d.transform_translation(pair).reshape(-1, J, J, 3).mean(dim=0)

    def _transfer_to(self, new_skeleton_tree: SkeletonTree):
        old = [self.skeleton_tree.index(n) for n in new_skeleton_tree]
        return SkeletonState.from_rotation_and_root_translation(new_skeleton_tree, self.global_rotation[..., old, :],
                                                                self.root_translation, is_local=False)

    def drop_nodes_by_names(self, node_names: List[str], estimate_local_translation_from_states: bool = True) -> "SkeletonState":
        pairwise = self._get_pairwise_average_translation() if estimate_local_translation_from_states else None
        return self._transfer_to(self.skeleton_tree.drop_nodes_by_names(node_names, pairwise))

    def keep_nodes_by_names(self, node_names: List[str], estimate_local_translation_from_states: bool = True) -> "SkeletonState":
        return self.drop_nodes_by_names([n for n in self.skeleton_tree if n not in node_names], estimate_local_translation_from_states)

    def _remapped_to(self, joint_mapping: Dict[str, str], target_skeleton_tree: SkeletonTree):
        inverse = {tgt: src for src, tgt in joint_mapping.items()}
        reduced = target_skeleton_tree.keep_nodes_by_names(list(inverse))
        assert len({len(joint_mapping), len(self.skeleton_tree), len(reduced)}) == 1, \
            "the joint mapping is not consistent with the skeleton trees"
        src_idx = [self.skeleton_tree.index(inverse[n]) for n in reduced]
        return SkeletonState.from_rotation_and_root_translation(reduced, self.local_rotation[..., src_idx, :],
                                                                self.root_translation, is_local=True)

    # ---- naive T-pose-relative retarget (skeleton3d.py:742-889)
    def retarget_to(self, joint_mapping: Dict[str, str], source_tpose_local_rotation, source_tpose_root_translation,
                    target_skeleton_tree: SkeletonTree, target_tpose_local_rotation, target_tpose_root_translation,
                    rotation_to_target_skeleton, scale_to_target_skeleton: float, z_up: bool = True) -> "SkeletonState":
        mk = SkeletonState.from_rotation_and_root_translation
        src_tpose = mk(copy.deepcopy(self.skeleton_tree), source_tpose_local_rotation, source_tpose_root_translation, True)
        tgt_tpose = mk(copy.deepcopy(target_skeleton_tree), target_tpose_local_rotation, target_tpose_root_translation, True)
        # 1. keep the mapped source joints only, then rename them onto the (reduced) target tree
        reduced = self.skeleton_tree.keep_nodes_by_names(list(joint_mapping), self._get_pairwise_average_translation())
        src_tpose = src_tpose._transfer_to(reduced)._remapped_to(joint_mapping, copy.deepcopy(target_skeleton_tree))
        src_state = self._transfer_to(reduced)._remapped_to(joint_mapping, copy.deepcopy(target_skeleton_tree))
        # 2. rotate both into the target's frame of reference
        rot = torch.as_tensor(rotation_to_target_skeleton)

        def aligned(state):
            lr = state.local_rotation.clone()
            lr[..., 0, :] = r3d.quat_mul_norm(rot, state.local_rotation[..., 0, :])
            return mk(state.skeleton_tree, lr, r3d.quat_rotate(rot, state.root_translation), True)

        src_tpose, src_state = aligned(src_tpose), aligned(src_state)
        # 3. root motion in target units
        root_diff = (src_state.root_translation - src_tpose.root_translation) * scale_to_target_skeleton
        # 4. rotation of the source relative to its T-pose, re-applied to the target T-pose
        cur = src_state.skeleton_tree
        tgt_rest = src_state.global_rotation[0, :].clone()
        for i, name in enumerate(cur):
            if name in tgt_tpose.skeleton_tree._node_indices:
                tgt_rest[i, :] = tgt_tpose.global_rotation[tgt_tpose.skeleton_tree.index(name), :]
        diff = r3d.quat_mul_norm(src_state.global_rotation, r3d.quat_inverse(src_tpose.global_rotation))
        new_global = r3d.quat_mul_norm(diff, tgt_rest)
        # 5. unmapped target joints inherit their nearest mapped ancestor
        lead = src_state.global_rotation.shape[:-2]
        out = r3d.quat_identity(list(lead) + [len(target_skeleton_tree)])
        for i, name in enumerate(target_skeleton_tree):
            while name not in cur._node_indices:
                name = target_skeleton_tree.parent_of(name)
            out[:, i, :] = new_global[:, cur.index(name), :]
        return mk(target_skeleton_tree, out, tgt_tpose.root_translation + root_diff, False).local_repr()

    def retarget_to_by_tpose(self, joint_mapping, source_tpose: "SkeletonState", target_tpose: "SkeletonState",
                             rotation_to_target_skeleton, scale_to_target_skeleton: float) -> "SkeletonState":
        assert len(source_tpose.tensor.shape) == 1 and len(target_tpose.tensor.shape) == 1, \
            "the retargeting script currently doesn't support vectorized operations"
        return self.retarget_to(joint_mapping, source_tpose.local_rotation, source_tpose.root_translation,
                                target_tpose.skeleton_tree, target_tpose.local_rotation, target_tpose.root_translation,
                                rotation_to_target_skeleton, scale_to_target_skeleton)


# ---------------------------------------------------------------------------------- SkeletonMotion
class SkeletonMotion(SkeletonState):
    """SkeletonState + per-joint global linear / angular velocity + fps:
    tensor = [J*4 | 3 | J*3 | J*3] (skeleton3d.py:954-964)."""

    def __init__(self, tensor_backend, skeleton_tree, is_local, fps, *args, **kwargs):
        self._fps = fps
        super().__init__(tensor_backend, skeleton_tree, is_local, *args, **kwargs)

    def clone(self):
        return SkeletonMotion(self.tensor.clone(), self.skeleton_tree, self._is_local, self._fps)

    invariant_property = property(lambda self: {"skeleton_tree": self.skeleton_tree, "is_local": self.is_local, "fps": self.fps})
    fps = property(lambda self: self._fps)
    time_delta = property(lambda self: 1.0 / self.fps)
    global_root_velocity = property(lambda self: self.global_velocity[..., 0, :])
    global_root_angular_velocity = property(lambda self: self.global_angular_velocity[..., 0, :])

    @property
    def global_velocity(self):
        J = self.num_joints
        return self.tensor[..., J * 4 + 3:J * 7 + 3].reshape(*self.tensor.shape[:-1], J, 3)

    @property
    def global_angular_velocity(self):
        J = self.num_joints
        return self.tensor[..., J * 7 + 3:J * 10 + 3].reshape(*self.tensor.shape[:-1], J, 3)

    @classmethod
    def from_state_vector_and_velocity(cls, skeleton_tree, state_vector, global_velocity, global_angular_velocity, is_local, fps):
        lead = state_vector.shape[:-1]
        v = torch.cat([state_vector, global_velocity.reshape(*lead, -1).to(state_vector.device),
                       global_angular_velocity.reshape(*lead, -1).to(state_vector.device)], dim=-1)
        return cls(v, skeleton_tree=skeleton_tree, is_local=is_local, fps=fps)

    @classmethod
    def from_skeleton_state(cls, skeleton_state: SkeletonState, fps: int):
        """skeleton3d.py:1026-1049: velocities from finite differences + gaussian smoothing, on the device."""
        assert type(skeleton_state) == SkeletonState, "expected type of {}, got {}".format(SkeletonState, type(skeleton_state))
        vel = SkeletonMotion._compute_velocity(skeleton_state.global_translation, 1 / fps)
        avel = SkeletonMotion._compute_angular_velocity(skeleton_state.global_rotation, 1 / fps)
        return cls.from_state_vector_and_velocity(skeleton_state.skeleton_tree, skeleton_state.tensor, vel, avel,
                                                  skeleton_state.is_local, fps)

    @staticmethod
    def _to_state_vector(rot, rt, vel, avel):
        lead = rot.shape[:-2]
        return torch.cat([SkeletonState._to_state_vector(rot, rt), vel.reshape(*lead, -1), avel.reshape(*lead, -1)], dim=-1)

    @classmethod
    def from_dict(cls, dict_repr, *args, **kwargs):
        t = _tensor_from_dict
        return cls(SkeletonMotion._to_state_vector(t(dict_repr["rotation"]), t(dict_repr["root_translation"]),
                                                   t(dict_repr["global_velocity"]), t(dict_repr["global_angular_velocity"])),
                   skeleton_tree=SkeletonTree.from_dict(dict_repr["skeleton_tree"]), is_local=dict_repr["is_local"], fps=dict_repr["fps"])

    def to_dict(self):
        return OrderedDict([("rotation", tensor_to_dict(self.rotation)), ("root_translation", tensor_to_dict(self.root_translation)),
                            ("global_velocity", tensor_to_dict(self.global_velocity)),
                            ("global_angular_velocity", tensor_to_dict(self.global_angular_velocity)),
                            ("skeleton_tree", self.skeleton_tree.to_dict()), ("is_local", self.is_local), ("fps", self.fps)])

    @classmethod
    def from_fbx(cls, *args, **kwargs):
        raise NotImplementedError("FBX reading is outside the retarget hot path (DESIGN.md section 8)")

    @staticmethod
    def _frames_first(x, width):
        """(..., T, J, width) -> list of (T, J, width) clips + the leading shape"""
        lead = x.shape[:-3]
        return x.reshape(-1, *x.shape[-3:]), lead

    @staticmethod
    def _compute_velocity(p, time_delta, guassian_filter=True):
        """skeleton3d.py:1126-1135: np.gradient along the frame axis (-3) / dt, gaussian_filter1d(sigma=2, 'nearest')."""
        eng = r3d._engine(p.device)
        clips, lead = SkeletonMotion._frames_first(p, 3)
        out = torch.stack([eng.motion_velocity(c, time_delta, guassian_filter) for c in clips]).reshape(p.shape)
        return out.to(p.device)

    @staticmethod
    def _compute_angular_velocity(r, time_delta: float, guassian_filter=True):
        """skeleton3d.py:1137-1146"""
        eng = r3d._engine(r.device)
        clips, lead = SkeletonMotion._frames_first(r, 4)
        out = torch.stack([eng.motion_angular_velocity(c, time_delta, guassian_filter) for c in clips])
        return out.reshape(*r.shape[:-1], 3).to(r.device)

    def crop(self, start: int, end: int, fps: Optional[int] = None):
        """skeleton3d.py:1148-1183: frames [start:end:old_fps/fps] (the frame axis leads)."""
        old = int(self.fps)
        new = old if fps is None else int(fps)
        assert old % new == 0, ("the resampling doesn't support fps with non-integer division from the original fps: "
                                "{} => {}".format(old, new))
        return SkeletonMotion(self.tensor[start:end:old // new].clone(), self.skeleton_tree, self.is_local, new)

    def retarget_to(self, joint_mapping, source_tpose_local_rotation, source_tpose_root_translation, target_skeleton_tree,
                    target_tpose_local_rotation, target_tpose_root_translation, rotation_to_target_skeleton,
                    scale_to_target_skeleton: float, z_up: bool = True) -> "SkeletonMotion":
        state = SkeletonState(self.tensor[..., :self.num_joints * 4 + 3], self.skeleton_tree, self.is_local)
        return SkeletonMotion.from_skeleton_state(
            state.retarget_to(joint_mapping, source_tpose_local_rotation, source_tpose_root_translation, target_skeleton_tree,
                              target_tpose_local_rotation, target_tpose_root_translation, rotation_to_target_skeleton,
                              scale_to_target_skeleton, z_up), self.fps)

    def retarget_to_by_tpose(self, joint_mapping, source_tpose, target_tpose, rotation_to_target_skeleton,
                             scale_to_target_skeleton: float, z_up: bool = True) -> "SkeletonMotion":
        return self.retarget_to(joint_mapping, source_tpose.local_rotation, source_tpose.root_translation,
                                target_tpose.skeleton_tree, target_tpose.local_rotation, target_tpose.root_translation,
                                rotation_to_target_skeleton, scale_to_target_skeleton, z_up)


class MotionDICT:
    """skeleton3d.py:1295-1314: a bare (frames, J, 3) position clip with its tree."""

    def __init__(self, gt, sk_tree, get_state=False) -> None:
        self.global_translation = gt.clone()
        self.skeleton_tree = sk_tree
        if (not get_state) and len(self.global_translation.shape) == 2:
            self.global_translation = self.global_translation[None, ...]

    def clone(self):
        return MotionDICT(self.global_translation.clone(), self.skeleton_tree)

    def __getitem__(self, t):
        return MotionDICT(self.global_translation[t].clone(), self.skeleton_tree, get_state=True)

    def __len__(self):
        return self.global_translation.shape[0]
