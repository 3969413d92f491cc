"""Drop-in for retarget/main.py:30-279 (`Retarget`, `RetargetHuV5fromMocap`): offline retarget of a whole
position clip (L,21,3) of the vtrdyn mocap skeleton onto Hu v5.

The reference walks the clip frame by frame in Python; here every stage is one batched kernel launch over
the clip (frames are independent except for the velocity smoothing at the very end):
  coord_transform + limb-length rescale   main.py:170-172,36-47    rescale_motion_kernel
  global-rotation rebuild                 main.py:116-152          bone_max_norm_kernel + rebuild_rotation_kernel
  SkeletonState / SkeletonMotion          main.py:154-160          rot_op / local_from_global / fk_limb / motion kernels
  shoulder + elbow solves per frame       main.py:201-240          pos_retarget_kernel<POS_MAIN>
The reference plots the two motions and returns nothing; this returns them."""
import torch

from .engine import POS_MAIN, TREE_ROBOT, engine_from_zero_poses
from .kinematics import RobotZeroPose
from .skeleton3d import SkeletonMotion, SkeletonState


class Retarget:
    def __init__(self, mocap_zero_pose: RobotZeroPose, target_zero_pose: RobotZeroPose):
        self.mocap_zero_pose = mocap_zero_pose
        self.target_zero_pose = target_zero_pose

    def cal_motion_local_rotation(self):
        pass

    @staticmethod
    def rescale_motion_to_standard_size(motion_global_translation, zero_pose: RobotZeroPose):
        """main.py:36-47: every bone rescaled to its zero-pose length, parents before children."""
        from .kinematics import _tree_engine
        g = motion_global_translation
        eng = _tree_engine(zero_pose.parent_indices, zero_pose.local_translation, device=g.device.index or 0 if g.is_cuda else 0)
        return eng.rescale_motion(0, g).reshape(g.shape).to(g.device)


class RetargetHuV5fromMocap(Retarget):
    def __init__(self, mocap_zero_pose: RobotZeroPose, target_zero_pose: RobotZeroPose, device=0):
        super().__init__(mocap_zero_pose, target_zero_pose)
        # this instance's own context, built from ITS zero poses (main.py:37-47,118-152,203-240 read every offset from them)
        self._eng, self._src, (parents, _, _) = engine_from_zero_poses(mocap_zero_pose, target_zero_pose, device)
        if parents.shape[0] != 21:
            raise ValueError(f"RetargetHuV5fromMocap needs the 21-joint vtrdyn zero pose, got {parents.shape[0]} joints")
        self._eng.configure_pos(POS_MAIN, self._src, TREE_ROBOT)

    def _rebuild_with_vtrdyn_zero_pose(self, motion_global_translation, fps=30) -> SkeletonMotion:
        """main.py:116-165"""
        g = motion_global_translation
        gq = self._eng.rebuild_global_rotation(self._src, g).to(g.device)
        state = SkeletonState.from_rotation_and_root_translation(self.mocap_zero_pose.skeleton_tree, gq, g[:, 0, :].clone(),
                                                                 is_local=False)
        motion = SkeletonMotion.from_skeleton_state(state, fps=fps)
        self.rebuild_error = float((motion.global_translation - g).abs().max())
        return motion

    def retarget_from_global_translation(self, global_translation, fps=30):
        """main.py:169-279.  global_translation (L,21,3).  Returns (mocap_motion, retargeted_motion)."""
        g = torch.as_tensor(global_translation, dtype=torch.float32)
        scaled = self._eng.rescale_motion(self._src, g, dir=[-1.0, -1.0, 1.0]).to(g.device)     # main.py:170-172
        mocap_motion = self._rebuild_with_vtrdyn_zero_pose(scaled, fps=fps)
        lq, _ = self._eng.retarget_main_arms(mocap_motion.global_rotation, mocap_motion.global_translation, want_dof=False)
        retargeted_state = SkeletonState.from_rotation_and_root_translation(
            self.target_zero_pose.skeleton_tree, lq.to(g.device), torch.zeros_like(mocap_motion.root_translation), is_local=True)
        return mocap_motion, SkeletonMotion.from_skeleton_state(retargeted_state, fps=fps)
