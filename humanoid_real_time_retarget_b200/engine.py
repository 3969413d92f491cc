"""Thin Python handle on an hrt_ctx (include/hrt_b200.h).  PyTorch is plumbing only: it owns the
device memory and the streams; every number is produced by the sm_100a kernels behind the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import robot_config as cfg

TREE_ROBOT, TREE_SOURCE, TREE_SOURCE_FULL = 0, 1, 2
FK_EXACT = 1
BQ_CLAMP, BQ_IK, BQ_PRE_TRANSFORMED, BQ_PACKED_IK, BQ_ACTIVE_SET = 1, 2, 4, 8, 16
POS_FULL_BODY_POS, POS_UPPER_BODY, POS_FULL_BODY, POS_MAIN = 0, 1, 2, 3
POS_CLAMP, POS_IK = 1, 2


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _np_ptr(a):
    return C.c_void_p(0 if a is None else a.ctypes.data)


def _f32c(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One hrt_ctx on one GPU.  Not thread-safe; make one per thread / stream of work."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.HrtError("no CUDA device visible: humanoid_real_time_retarget_b200 has no CPU path")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        _lib.check(self.lib.hrt_ctx_create(self.device.index, C.byref(h)))
        self._h = h
        self._trees = {}
        self._pos_stream_cfg = None        # what the position / quaternion streaming mailboxes are currently opened for
        self._bq_stream_cfg = None
        self.sm_count = self.lib.hrt_ctx_sm_count(h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hrt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ trees
    def set_tree(self, tree, parents, offsets, dof_axis=None, lower=None, upper=None, t2z=None):
        parents = np.ascontiguousarray(np.asarray(parents, dtype=np.int32))
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.float32).reshape(-1, 3))
        J = parents.shape[0]
        assert offsets.shape[0] == J
        ax = None if dof_axis is None else np.ascontiguousarray(np.asarray(dof_axis, dtype=np.uint8))
        lo = None if lower is None else np.ascontiguousarray(np.asarray(lower, dtype=np.float32))
        hi = None if upper is None else np.ascontiguousarray(np.asarray(upper, dtype=np.float32))
        tz = None if t2z is None else np.ascontiguousarray(np.asarray(t2z, dtype=np.float32).reshape(J, 4))
        for a in (ax, lo, hi):
            assert a is None or a.shape[0] == J - 1
        self._trees.pop(tree, None)
        self._pos_stream_cfg = self._bq_stream_cfg = None
        _lib.check(self.lib.hrt_set_tree(self._h, tree, J, _np_ptr(parents), _np_ptr(offsets), _np_ptr(ax),
                                         _np_ptr(lo), _np_ptr(hi), _np_ptr(tz)))
        self._trees[tree] = J

    def set_standard_trees(self, robot="hu_v5", precise_gripper=True):
        """Hu v5 (or Hu) robot + vtrdyn / vtrdyn_full sources from the bundled tables, and the fused
        quaternion-path wiring of body_retargeter.py:40-73."""
        sk = cfg.skeleton_tables()
        if robot == "hu_v5":
            self.set_tree(TREE_ROBOT, sk["hu_v5_zero_pose/parents"], sk["hu_v5_zero_pose/offsets"],
                          cfg.Hu_v5_DOF_AXIS, cfg.Hu_v5_DOF_LOWER, cfg.Hu_v5_DOF_UPPER)
        elif robot == "hu":
            self.set_tree(TREE_ROBOT, sk["hu_zero_pose/parents"], sk["hu_zero_pose/offsets"],
                          cfg.Hu_DOF_AXIS, cfg.Hu_DOF_LOWER, cfg.Hu_DOF_UPPER)
        else:
            raise ValueError(robot)
        self.set_tree(TREE_SOURCE, sk["vtrdyn_zero_pose/parents"], sk["vtrdyn_zero_pose/offsets"], t2z=sk["t2z/vtrdyn"])
        self.set_tree(TREE_SOURCE_FULL, sk["vtrdyn_full_zero_pose/parents"], sk["vtrdyn_full_zero_pose/offsets"],
                      t2z=sk["t2z/vtrdyn_full"])
        if robot == "hu_v5":
            self.configure_body_quat(TREE_SOURCE, TREE_ROBOT, cfg.VTRDYN_ARM_JOINTS, cfg.HU_V5_ARM_FIRST)
            self.configure_pos(POS_FULL_BODY_POS, TREE_SOURCE_FULL, TREE_ROBOT, sk["vtrdyn_full_zero_pose/global_translation"], precise_gripper)
            self.configure_pos(POS_UPPER_BODY, TREE_SOURCE, TREE_ROBOT)
            self.configure_pos(POS_FULL_BODY, TREE_SOURCE_FULL, TREE_ROBOT)
            self.configure_pos(POS_MAIN, TREE_SOURCE, TREE_ROBOT)
        return self

    def configure_body_quat(self, src_tree, rob_tree, src_joints, rob_first):
        sj = np.ascontiguousarray(np.asarray(src_joints, dtype=np.int32).reshape(2, 5))
        rf = np.ascontiguousarray(np.asarray(rob_first, dtype=np.int32).reshape(2))
        _lib.check(self.lib.hrt_configure_body_quat(self._h, src_tree, rob_tree, _np_ptr(sj), _np_ptr(rf)))
        self._bq = (self._trees[src_tree], self._trees[rob_tree])
        self._bq_src_joints = sj.copy()
        self._bq_stream_cfg = None         # the C side closed the stream: it held the previous wiring by value

    def configure_pos(self, mode, src_tree, rob_tree, src_global_t=None, precise_gripper=False):
        gt = None if src_global_t is None else np.ascontiguousarray(np.asarray(src_global_t, dtype=np.float32))
        _lib.check(self.lib.hrt_configure_pos(self._h, mode, src_tree, rob_tree, _np_ptr(gt), int(bool(precise_gripper))))
        self._pos_stream_cfg = None        # the C side closed the stream (and stopped a resident server)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ kernels on device tensors
    def fk_local_quats(self, tree, local_q, root_t=None, exact=False, want_q=True, want_t=True):
        J = self._trees[tree]
        local_q = _f32c(local_q, self.device)
        B = local_q.numel() // (J * 4)
        assert local_q.numel() == B * J * 4
        root_t = None if root_t is None else _f32c(root_t, self.device)
        gq = torch.empty((B, J, 4), device=self.device, dtype=torch.float32) if want_q else None
        gt = torch.empty((B, J, 3), device=self.device, dtype=torch.float32) if want_t else None
        _lib.check(self.lib.hrt_fk_local_quats(self._h, tree, B, _ptr(local_q), _ptr(root_t), _ptr(gq), _ptr(gt),
                                               FK_EXACT if exact else 0, self._stream()))
        return gq, gt

    def fk_angles(self, tree, angles, root_t=None, root_q=None, clip=True, exact=False, out=None):
        J = self._trees[tree]
        angles = _f32c(angles, self.device)
        B = angles.numel() // (J - 1)
        assert angles.numel() == B * (J - 1)
        root_t = None if root_t is None else _f32c(root_t, self.device)
        root_q = None if root_q is None else _f32c(root_q, self.device)
        if out is None:
            gq = torch.empty((B, J, 4), device=self.device, dtype=torch.float32)
            gt = torch.empty((B, J, 3), device=self.device, dtype=torch.float32)
        else:
            gq, gt = out
        _lib.check(self.lib.hrt_fk_angles(self._h, tree, B, _ptr(angles), _ptr(root_t), _ptr(root_q), int(bool(clip)),
                                          _ptr(gq), _ptr(gt), FK_EXACT if exact else 0, self._stream()))
        return gq, gt

    def fk_jacobian(self, tree, angles, links, root_t=None, root_q=None, clip=True, out=None):
        J = self._trees[tree]
        angles = _f32c(angles, self.device)
        B = angles.numel() // (J - 1)
        links = np.ascontiguousarray(np.asarray(links, dtype=np.int32))
        K = links.shape[0]
        root_t = None if root_t is None else _f32c(root_t, self.device)
        root_q = None if root_q is None else _f32c(root_q, self.device)
        jac = out if out is not None else torch.empty((B, K, 6, J - 1), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.hrt_fk_jacobian(self._h, tree, B, _ptr(angles), _ptr(root_t), _ptr(root_q), int(bool(clip)),
                                            _np_ptr(links), K, _ptr(jac), self._stream()))
        return jac

    def fk_vjp(self, tree, angles, g_gq=None, g_gt=None, root_t=None, root_q=None, clip=True, want_root=True):
        """Backward of fk_angles: upstream gradients of (global quats, link positions) -> (g_angles, g_root_t, g_root_q)."""
        J = self._trees[tree]
        angles = _f32c(angles, self.device)
        B = angles.numel() // (J - 1)
        g_gq = None if g_gq is None else _f32c(g_gq, self.device)
        g_gt = None if g_gt is None else _f32c(g_gt, self.device)
        root_t = None if root_t is None else _f32c(root_t, self.device)
        root_q = None if root_q is None else _f32c(root_q, self.device)
        ga = torch.empty((B, J - 1), device=self.device, dtype=torch.float32)
        grt = torch.empty((B, 3), device=self.device, dtype=torch.float32) if want_root else None
        grq = torch.empty((B, 4), device=self.device, dtype=torch.float32) if want_root else None
        _lib.check(self.lib.hrt_fk_vjp(self._h, tree, B, _ptr(angles), _ptr(root_t), _ptr(root_q), int(bool(clip)), _ptr(g_gq), _ptr(g_gt),
                                       _ptr(ga), _ptr(grt), _ptr(grq), self._stream()))
        return ga, grt, grq

    def ik_refine(self, theta0, pe_t, pw_t, qw_t, iters=10, damping=0.1, rot_weight=0.2, active_set=False, want_residual=False):
        """Stand-alone DLS refinement of both arms: theta0 (B,2,7), pe_t / pw_t (B,2,3), qw_t (B,2,4) -> theta (B,2,7)
        [, residual (B,2,iters+1)].  Needs configure_body_quat (robot arm tables)."""
        theta0, pe_t, pw_t, qw_t = (_f32c(x, self.device) for x in (theta0, pe_t, pw_t, qw_t))
        B = theta0.numel() // 14
        assert pe_t.numel() == B * 6 and pw_t.numel() == B * 6 and qw_t.numel() == B * 8
        th = torch.empty((B, 2, 7), device=self.device, dtype=torch.float32)
        res = torch.empty((B, 2, iters + 1), device=self.device, dtype=torch.float32) if want_residual else None
        _lib.check(self.lib.hrt_ik_refine(self._h, B, _ptr(theta0), _ptr(pe_t), _ptr(pw_t), _ptr(qw_t), int(iters), float(damping),
                                          float(rot_weight), BQ_ACTIVE_SET if active_set else 0, _ptr(th), _ptr(res), self._stream()))
        return (th, res) if want_residual else th

    def local_from_global(self, tree, global_q):
        J = self._trees[tree]
        global_q = _f32c(global_q, self.device)
        B = global_q.numel() // (J * 4)
        out = torch.empty_like(global_q)
        _lib.check(self.lib.hrt_local_from_global(self._h, tree, B, _ptr(global_q), _ptr(out), self._stream()))
        return out

    def zero_pose_transform(self, tree, global_q, variant=0):
        J = self._trees[tree]
        global_q = _f32c(global_q, self.device)
        B = global_q.numel() // (J * 4)
        out = torch.empty_like(global_q)
        _lib.check(self.lib.hrt_zero_pose_transform(self._h, tree, B, _ptr(global_q), variant, _ptr(out), self._stream()))
        return out

    def retarget_body_quat(self, src_gq, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2,
                           want_local_q=True, want_dof=True, want_link_pos=True, out=None):
        """Fused quaternion path on device tensors.  Returns (robot_local_q, dof_pos, link_pos)."""
        JS, JR = self._bq
        src_gq = _f32c(src_gq, self.device)
        B = src_gq.numel() // (JS * 4)
        assert src_gq.numel() == B * JS * 4
        if out is None:
            lq = torch.empty((B, JR, 4), device=self.device, dtype=torch.float32) if want_local_q else None
            dof = torch.empty((B, JR - 1), device=self.device, dtype=torch.float32) if want_dof else None
            lp = torch.empty((B, JR, 3), device=self.device, dtype=torch.float32) if want_link_pos else None
        else:
            lq, dof, lp = out
        _lib.check(self.lib.hrt_retarget_body_quat(self._h, B, _ptr(src_gq), flags, ik_iters, damping, rot_weight,
                                                   _ptr(lq), _ptr(dof), _ptr(lp), self._stream()))
        return lq, dof, lp

    # ------------------------------------------------------------------ multi-GPU reassembly over NVLink (configs[4])
    def peer_alloc(self, nbytes):
        """Device allocation exportable to the other ranks: (device pointer, 64-byte CUDA IPC handle)."""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _lib.check(self.lib.hrt_peer_alloc(self._h, int(nbytes), C.byref(p), h))
        return p.value, h.raw

    def peer_open(self, handle):
        p = C.c_void_p()
        _lib.check(self.lib.hrt_peer_open(self._h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def peer_close(self, ptr):
        _lib.check(self.lib.hrt_peer_close(self._h, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        _lib.check(self.lib.hrt_peer_free(self._h, C.c_void_p(ptr)))

    def retarget_body_quat_gather(self, src_gq, peer_dof_ptrs, frame0, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2, link_pos=None):
        """The fused quaternion path on this rank's frames, its dof rows stored into EVERY rank's clip-wide buffer
        (peer_dof_ptrs: device pointers, own buffer included) at frame `frame0` onwards."""
        JS, JR = self._bq
        src_gq = _f32c(src_gq, self.device)
        B = src_gq.numel() // (JS * 4)
        arr = (C.c_void_p * len(peer_dof_ptrs))(*peer_dof_ptrs)
        _lib.check(self.lib.hrt_retarget_body_quat_gather(self._h, B, _ptr(src_gq), flags, ik_iters, damping, rot_weight, _ptr(link_pos),
                                                          len(peer_dof_ptrs), arr, int(frame0), self._stream()))

    def retarget_body_quat_multicast(self, src_gq, mc_dof_ptr, frame0, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2, link_pos=None):
        """The fused quaternion path on this rank's frames, its dof rows published ONCE through the NVSwitch multicast
        address `mc_dof_ptr` of the ranks' clip-wide buffers (multimem.st; the switch writes every rank's copy)."""
        JS, JR = self._bq
        src_gq = _f32c(src_gq, self.device)
        B = src_gq.numel() // (JS * 4)
        _lib.check(self.lib.hrt_retarget_body_quat_multicast(self._h, B, _ptr(src_gq), flags, ik_iters, damping, rot_weight, _ptr(link_pos),
                                                             C.c_void_p(int(mc_dof_ptr)), int(frame0), self._stream()))

    def reassembly_layout(self, n_total, shard_frames):
        """Bytes of the symmetric allocation the in-kernel reassembly needs per rank: (staging, flag offset, total, max rounds)."""
        arr = (C.c_int64 * len(shard_frames))(*[int(x) for x in shard_frames])
        st, fo, tot, mr = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int()
        _lib.check(self.lib.hrt_reassembly_layout(self._h, int(n_total), len(shard_frames), arr, C.byref(st), C.byref(fo), C.byref(tot), C.byref(mr)))
        return st.value, fo.value, tot.value, mr.value

    def retarget_body_quat_reassemble(self, src_gq, full_dof, n_total, my_rank, shard_lo, shard_frames, symm_ptr, symm_mc_ptr, epoch,
                                      flags=0, ik_iters=10, damping=0.1, rot_weight=0.2, link_pos=None):
        """The fused quaternion path on this rank's shard with the reassembly of dof_pos INSIDE the kernel: packed hinge angles
        go out through the multicast address `symm_mc_ptr`, the peers' packed rows that have landed in `symm_ptr` are unpacked
        into `full_dof` (n_total, D) between this rank's own rounds."""
        JS, JR = self._bq
        B = int(shard_frames[my_rank])
        src = _f32c(src_gq, self.device) if B else None
        lo = (C.c_int64 * len(shard_lo))(*[int(x) for x in shard_lo])
        nn = (C.c_int64 * len(shard_frames))(*[int(x) for x in shard_frames])
        _lib.check(self.lib.hrt_retarget_body_quat_reassemble(self._h, B, _ptr(src), flags, ik_iters, damping, rot_weight, _ptr(link_pos),
                                                              _ptr(full_dof), int(n_total), len(shard_frames), int(my_rank), lo, nn,
                                                              C.c_void_p(int(symm_ptr)), C.c_void_p(int(symm_mc_ptr)),
                                                              int(epoch) & 0xFFFFFFFF, self._stream()))

    def peer_barrier(self, peer_flag_ptrs, my_rank, epoch):
        arr = (C.c_void_p * len(peer_flag_ptrs))(*peer_flag_ptrs)
        _lib.check(self.lib.hrt_peer_barrier(self._h, len(peer_flag_ptrs), int(my_rank), arr, int(epoch) & 0xFFFFFFFF, self._stream()))

    def _pos_outputs(self, B, want_local_q, want_dof, want_body_gq=False):
        lq = torch.empty((B, 31, 4), device=self.device, dtype=torch.float32) if want_local_q else None
        dof = torch.empty((B, 30), device=self.device, dtype=torch.float32) if want_dof else None
        bq = torch.empty((B, 59, 4), device=self.device, dtype=torch.float32) if want_body_gq else None
        return lq, dof, bq

    def retarget_full_body_pos(self, body_t, lhand_t, rhand_t, want_local_q=True, want_dof=True, want_body_gq=True, out=None,
                               flags=0, ik_iters=10, damping=0.1, rot_weight=0.2):
        """VtrdynFullBodyPosRetargeter.retarget on (B,21,3), (B,20,3), (B,20,3) device tensors.  flags: POS_CLAMP |
        POS_IK add the joint limits / the fused limit-aware refinement (ours)."""
        body_t, lhand_t, rhand_t = (_f32c(x, self.device) for x in (body_t, lhand_t, rhand_t))
        B = body_t.numel() // 63
        assert lhand_t.numel() == B * 60 and rhand_t.numel() == B * 60
        lq, dof, bq = out if out is not None else self._pos_outputs(B, want_local_q, want_dof, want_body_gq)
        if flags:
            _lib.check(self.lib.hrt_retarget_full_body_pos_ex(self._h, B, _ptr(body_t), _ptr(lhand_t), _ptr(rhand_t), flags, ik_iters,
                                                              damping, rot_weight, _ptr(lq), _ptr(dof), _ptr(bq), self._stream()))
        else:
            _lib.check(self.lib.hrt_retarget_full_body_pos(self._h, B, _ptr(body_t), _ptr(lhand_t), _ptr(rhand_t),
                                                           _ptr(lq), _ptr(dof), _ptr(bq), self._stream()))
        return lq, dof, bq

    def retarget_upper_body(self, body_t, want_local_q=True, want_dof=True):
        """HuUpperBodyFromMocapRetarget.retarget_from_global_translation on (B,21,3)."""
        body_t = _f32c(body_t, self.device)
        B = body_t.numel() // 63
        lq, dof, _ = self._pos_outputs(B, want_local_q, want_dof)
        _lib.check(self.lib.hrt_retarget_upper_body(self._h, B, _ptr(body_t), _ptr(lq), _ptr(dof), self._stream()))
        return lq, dof

    def retarget_full_body(self, body_q, body_t, lhand_t, rhand_t, want_local_q=True, want_dof=True):
        """VtrdynFullBodyRetargeter.retarget on (B,21,4), (B,21,3), (B,20,3), (B,20,3)."""
        body_q, body_t, lhand_t, rhand_t = (_f32c(x, self.device) for x in (body_q, body_t, lhand_t, rhand_t))
        B = body_t.numel() // 63
        lq, dof, _ = self._pos_outputs(B, want_local_q, want_dof)
        _lib.check(self.lib.hrt_retarget_full_body(self._h, B, _ptr(body_q), _ptr(body_t), _ptr(lhand_t), _ptr(rhand_t),
                                                   _ptr(lq), _ptr(dof), self._stream()))
        return lq, dof

    def retarget_full_body_pos_wire(self, body23_t, lhand_t, rhand_t, want_local_q=True, want_dof=True, want_body_gq=False):
        """The same solver on the mocap WIRE layout (sim_full_body_teleop.py:109-112): body (B,23,3), hands in
        HandNodes order (B,20,3); the 23->21 / finger reorder happens inside the kernel's index tables."""
        body23_t, lhand_t, rhand_t = (_f32c(x, self.device) for x in (body23_t, lhand_t, rhand_t))
        B = body23_t.numel() // 69
        assert body23_t.numel() == B * 69 and lhand_t.numel() == B * 60 and rhand_t.numel() == B * 60
        lq, dof, bq = self._pos_outputs(B, want_local_q, want_dof, want_body_gq)
        _lib.check(self.lib.hrt_retarget_full_body_pos_wire(self._h, B, _ptr(body23_t), _ptr(lhand_t), _ptr(rhand_t),
                                                            _ptr(lq), _ptr(dof), _ptr(bq), self._stream()))
        return lq, dof, bq

    def retarget_main_arms(self, body_q, body_t, want_local_q=True, want_dof=True):
        """The per-frame arm solves of RetargetHuV5fromMocap (retarget/main.py:201-240), batched."""
        body_q, body_t = _f32c(body_q, self.device), _f32c(body_t, self.device)
        B = body_t.numel() // 63
        assert body_q.numel() == B * 84
        lq, dof, _ = self._pos_outputs(B, want_local_q, want_dof)
        _lib.check(self.lib.hrt_retarget_main_arms(self._h, B, _ptr(body_q), _ptr(body_t), _ptr(lq), _ptr(dof), self._stream()))
        return lq, dof

    # ------------------------------------------------------------------ clip-level stages
    def rescale_motion(self, tree, global_t, dir=None):
        """Retarget.rescale_motion_to_standard_size (retarget/main.py:36-47) [+ coord_transform(dir)]."""
        J = self._trees[tree]
        global_t = _f32c(global_t, self.device)
        B = global_t.numel() // (J * 3)
        out = torch.empty_like(global_t)
        d = None if dir is None else np.ascontiguousarray(np.asarray(dir, dtype=np.float32).reshape(3))
        _lib.check(self.lib.hrt_rescale_motion(self._h, tree, B, _ptr(global_t), _np_ptr(d), _ptr(out), self._stream()))
        return out

    def rebuild_global_rotation(self, tree, global_t, kabsch_joints=(0, 10), kabsch_pts=((4, 1, 7), (17, 13, 11))):
        """Global rotations rebuilt from joint positions (retarget/main.py:116-152)."""
        J = self._trees[tree]
        global_t = _f32c(global_t, self.device)
        B = global_t.numel() // (J * 3)
        kj = np.ascontiguousarray(np.asarray(kabsch_joints, dtype=np.int32))
        kp = np.ascontiguousarray(np.asarray(kabsch_pts, dtype=np.int32).reshape(-1))
        out = torch.empty((B, J, 4), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.hrt_rebuild_global_rotation(self._h, tree, B, _ptr(global_t), kj.shape[0], _np_ptr(kj), _np_ptr(kp),
                                                        _ptr(out), self._stream()))
        return out

    def motion_velocity(self, global_t, dt, gaussian=True):
        """SkeletonMotion._compute_velocity (skeleton3d.py:1126-1135) on (T,J,3)."""
        global_t = _f32c(global_t, self.device)
        T, J = global_t.shape[0], global_t.shape[1]
        out = torch.empty_like(global_t)
        scratch = torch.empty_like(global_t) if gaussian else None
        _lib.check(self.lib.hrt_motion_velocity(self._h, T, J, _ptr(global_t), float(np.float32(dt)), int(bool(gaussian)),
                                                _ptr(scratch), _ptr(out), self._stream()))
        return out

    def motion_angular_velocity(self, global_q, dt, gaussian=True):
        """SkeletonMotion._compute_angular_velocity (skeleton3d.py:1137-1146) on (T,J,4)."""
        global_q = _f32c(global_q, self.device)
        T, J = global_q.shape[0], global_q.shape[1]
        out = torch.empty((T, J, 3), device=self.device, dtype=torch.float32)
        scratch = torch.empty_like(out) if gaussian else None
        _lib.check(self.lib.hrt_motion_angular_velocity(self._h, T, J, _ptr(global_q), float(np.float32(dt)), int(bool(gaussian)),
                                                        _ptr(scratch), _ptr(out), self._stream()))
        return out

    def forward_vector(self, global_t, left_shoulder, right_shoulder, left_hip, right_hip, sigma=20):
        """SkeletonState.compute_forward_vector (skeleton3d.py:542-566) on (T,J,3): (T,3) float64 like the reference."""
        global_t = _f32c(global_t, self.device)
        T, J = global_t.shape[0], global_t.shape[1]
        out = torch.empty((T, 3), device=self.device, dtype=torch.float64)
        scratch = torch.empty_like(out)
        _lib.check(self.lib.hrt_forward_vector(self._h, T, J, _ptr(global_t), int(left_shoulder), int(right_shoulder), int(left_hip),
                                               int(right_hip), float(sigma), _ptr(scratch), _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ element-wise rotation algebra
    def rot_op_info(self, op):
        ni, no = C.c_int(), C.c_int()
        wi, wo = (C.c_int * 4)(), (C.c_int * 3)()
        _lib.check(self.lib.hrt_rot_op_info(op, C.byref(ni), wi, C.byref(no), wo))
        return list(wi)[:ni.value], list(wo)[:no.value]

    def rot_op(self, op, n, ins, periods, outs, iparam=0, fparam=0.0):
        """ins / outs: contiguous fp32 device tensors; periods[k] = 0 (n rows) or the row count that repeats."""
        pin = (C.c_void_p * 4)(*([t.data_ptr() for t in ins] + [0] * (4 - len(ins))))
        per = (C.c_int64 * 4)(*(list(periods) + [0] * (4 - len(periods))))
        pout = (C.c_void_p * 3)(*([t.data_ptr() for t in outs] + [0] * (3 - len(outs))))
        _lib.check(self.lib.hrt_rot_op(self._h, op, n, pin, per, iparam, fparam, pout, self._stream()))

    def max_norm3(self, v):
        v = _f32c(v, self.device)
        out = C.c_float()
        _lib.check(self.lib.hrt_max_norm3(self._h, v.numel() // 3, _ptr(v), C.byref(out), self._stream()))
        return out.value

    def cal_joint_quat(self, zero_t, motion_t):
        """cal_joint_quat (transform3d.py:32-50): zero_t (b|1,n,3), motion_t (b,n,3) -> (b,4)."""
        zero_t, motion_t = _f32c(zero_t, self.device), _f32c(motion_t, self.device)
        b, n = motion_t.shape[0], motion_t.shape[1]
        zb = zero_t.numel() // (n * 3)
        out = torch.empty((b, 4), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.hrt_cal_joint_quat(self._h, b, n, _ptr(zero_t), 0 if zb == b else zb, _ptr(motion_t), _ptr(out),
                                               self._stream()))
        return out

    def stream_pos_open(self, wire_layout=False, persistent=False, clamp=False, ik=False, body_gq=False, mode=POS_FULL_BODY_POS):
        """persistent=True: a resident one-warp server polls the mailbox (no launch / sync per frame);
        clamp / ik: joint limits / + 10 limit-aware refinement steps, as retarget_full_body_pos(flags=...);
        body_gq=True: the mailbox also carries the (59,4) body quaternions (stream_pos_frame_tensors);
        mode: POS_FULL_BODY_POS, POS_UPPER_BODY or POS_FULL_BODY (the last two: launch per frame only)."""
        _lib.check(self.lib.hrt_stream_pos_open(self._h, (1 if wire_layout else 0) | (2 if persistent else 0) |
                                                (4 if clamp else 0) | (8 if ik else 0) | (16 if body_gq else 0) | (int(mode) << 8)))
        self._pos_stream_cfg = (bool(wire_layout), bool(persistent), bool(clamp), bool(ik), bool(body_gq), int(mode))

    def stream_pos_frame(self, body_np, lhand_np, rhand_np, out_local_q=None, out_dof=None):
        """numpy float32 in / out, one frame of the position path."""
        _lib.check(self.lib.hrt_stream_pos_frame(self._h, _np_ptr(body_np), _np_ptr(lhand_np), _np_ptr(rhand_np),
                                                 _np_ptr(out_local_q), _np_ptr(out_dof)))

    def stream_pos_frame_tensors(self, body, lhand, rhand, out_local_q, out_dof, out_body_gq=None, body_q=None):
        """The same on contiguous float32 CPU tensors (data_ptr() is several microseconds cheaper than numpy's ctypes
        view on the per-frame path); out_body_gq needs stream_pos_open(body_gq=True), body_q is mode POS_FULL_BODY's input."""
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.lib.hrt_stream_pos_frame_ex(self._h, p(body), p(lhand), p(rhand), p(body_q), p(out_local_q), p(out_dof),
                                                    p(out_body_gq)))

    def stream_pos_close(self):
        _lib.check(self.lib.hrt_stream_pos_close(self._h))
        self._pos_stream_cfg = None

    # ------------------------------------------------------------------ host-buffer (reference-facing) calls
    def retarget_body_quat_host(self, src_gq, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2,
                                out_local_q=None, out_dof=None, out_link_pos=None):
        """src_gq and the outputs are HOST tensors / arrays (pinned for full speed).  Copies are
        inside the call (chunked, overlapped)."""
        JS, JR = self._bq
        assert src_gq.device.type == "cpu" and src_gq.dtype == torch.float32 and src_gq.is_contiguous()
        B = src_gq.numel() // (JS * 4)
        for t in (out_local_q, out_dof, out_link_pos):
            assert t is None or (t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous())
        _lib.check(self.lib.hrt_retarget_body_quat_host(self._h, B, _ptr(src_gq), flags, ik_iters, damping, rot_weight,
                                                        _ptr(out_local_q), _ptr(out_dof), _ptr(out_link_pos)))
        return out_local_q, out_dof, out_link_pos

    def host_input_bytes_per_frame(self, out_local_q=False, out_dof=True, out_link_pos=False):
        """Bytes of a source frame that hrt_retarget_body_quat_host moves across PCIe for a call with these outputs.  The solver
        reads 9 source joints; when the joint range holding them is at most 3/4 of a row (and at least 128 B) and the
        outputs are the smaller side of the traffic, only that column range is copied (one strided copy per chunk);
        same rule as the C side."""
        JS, JR = self._bq
        w = (int(self._bq_src_joints.max()) - int(self._bq_src_joints.min()) + 1) * 16
        out_b = (JR * 16 if out_local_q else 0) + ((JR - 1) * 4 if out_dof else 0) + (JR * 12 if out_link_pos else 0)
        return w if (w * 4 <= JS * 16 * 3 and w >= 128 and out_b < JS * 16) else JS * 16

    def retarget_full_body_pos_host(self, body_t, lhand_t, rhand_t, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2,
                                    out_local_q=None, out_dof=None):
        """Position path on HOST tensors (pinned for full speed); copies are inside the call (chunked, overlapped)."""
        for t in (body_t, lhand_t, rhand_t, out_local_q, out_dof):
            assert t is None or (t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous())
        B = body_t.numel() // 63
        assert lhand_t.numel() == B * 60 and rhand_t.numel() == B * 60
        _lib.check(self.lib.hrt_retarget_full_body_pos_host(self._h, B, _ptr(body_t), _ptr(lhand_t), _ptr(rhand_t), flags, ik_iters,
                                                            damping, rot_weight, _ptr(out_local_q), _ptr(out_dof)))
        return out_local_q, out_dof

    def stream_open(self, flags=0, ik_iters=10, damping=0.1, rot_weight=0.2, persistent=False):
        """persistent=True: a resident one-warp server polls the mailbox (no launch / sync per frame)."""
        _lib.check(self.lib.hrt_stream_open(self._h, flags | (32 if persistent else 0), ik_iters, damping, rot_weight))
        self._bq_stream_cfg = (int(flags), int(ik_iters), float(damping), float(rot_weight), bool(persistent))

    def stream_frame(self, src_gq_np, out_local_q=None, out_dof=None, out_link_pos=None):
        """numpy float32 in / out, one frame."""
        _lib.check(self.lib.hrt_stream_frame(self._h, _np_ptr(src_gq_np), _np_ptr(out_local_q), _np_ptr(out_dof),
                                             _np_ptr(out_link_pos)))

    def stream_frame_tensors(self, src_gq, out_local_q=None, out_dof=None, out_link_pos=None):
        """The same on contiguous float32 CPU tensors."""
        _lib.check(self.lib.hrt_stream_frame(self._h, src_gq.data_ptr(),
                                             out_local_q.data_ptr() if out_local_q is not None else None,
                                             out_dof.data_ptr() if out_dof is not None else None,
                                             out_link_pos.data_ptr() if out_link_pos is not None else None))

    def stream_close(self):
        _lib.check(self.lib.hrt_stream_close(self._h))
        self._bq_stream_cfg = None


def _zp_arrays(zero_pose):
    """(parents int32 (J,), offsets float32 (J,3), global translations float32 (J,3)) of a RobotZeroPose-like object."""
    def arr(x, dt):
        x = x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        return np.ascontiguousarray(x, dtype=dt)
    parents = arr(zero_pose.parent_indices, np.int32).reshape(-1)
    off = arr(zero_pose.local_translation, np.float32).reshape(-1, 3)
    glob = arr(zero_pose.global_translation, np.float32).reshape(-1, 3)
    if off.shape[0] != parents.shape[0] or glob.shape[0] != parents.shape[0]:
        raise ValueError(f"zero pose tables disagree: {parents.shape[0]} parents, {off.shape[0]} offsets, {glob.shape[0]} positions")
    return parents, off, glob


def engine_from_zero_poses(source_zero_pose, target_zero_pose, device=0):
    """A NEW context whose trees are the zero poses a solver was constructed with -- the reference's solvers read every
    offset from those objects (full_body_pos_retargeter.py:69-107,139,162,184; retarget_solver.py:49-86;
    full_body_retargeter.py:59-99,152; body_retargeter.py:35,38), never from module-level tables.  The robot must be the
    31-joint Hu v5 the reference's solvers hard-code (`Hu_DOF_AXIS` of Hu_v5.py, robot_local_rotation[12..27]); the source
    is the 21-joint vtrdyn or the 59-joint vtrdyn_full skeleton.  Returns (engine, source tree slot, source tables)."""
    sp, so, sg = _zp_arrays(source_zero_pose)
    tp, to, _ = _zp_arrays(target_zero_pose)
    if tp.shape[0] != 31:
        raise ValueError(f"target zero pose has {tp.shape[0]} joints: the retarget solvers write Hu v5 joints 12-29 and "
                         "decompose with the 30-entry Hu_DOF_AXIS (retarget/robot_config/Hu_v5.py:12-18)")
    if sp.shape[0] not in (21, 59):
        raise ValueError(f"source zero pose has {sp.shape[0]} joints: expected the 21-joint vtrdyn or the 59-joint vtrdyn_full skeleton")
    eng = Engine(device)
    eng.set_tree(TREE_ROBOT, tp, to, cfg.Hu_v5_DOF_AXIS, cfg.Hu_v5_DOF_LOWER, cfg.Hu_v5_DOF_UPPER)
    sk = cfg.skeleton_tables()
    if sp.shape[0] == 21:
        slot, t2z = TREE_SOURCE, sk["t2z/vtrdyn"]
    else:
        slot, t2z = TREE_SOURCE_FULL, sk["t2z/vtrdyn_full"]
    # T2Z (parse_mocap.py:71-78,97-104) is built from rotations only: it does not depend on the offsets
    eng.set_tree(slot, sp, so, t2z=t2z)
    return eng, slot, (sp, so, sg)


_default = {}


def default_engine(device=0, robot="hu_v5"):
    key = (device, robot)
    if key not in _default:
        _default[key] = Engine(device).set_standard_trees(robot)
    return _default[key]
