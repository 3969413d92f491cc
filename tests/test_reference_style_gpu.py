"""The reference's own test scripts, re-expressed against the drop-in modules (same calls, same known answers):
poselib/poselib/core/tests/test_rotation.py:12-56 and poselib/poselib/skeleton/tests/test_skeleton.py:26-39."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ANT_LIKE_MJCF = """<mujoco model="quadruped">
  <worldbody>
    <body name="torso" pos="0 0 0.75">
      <body name="front_left_leg" pos="0 0 0">
        <body name="aux_1" pos="0.2 0.2 0">
          <body name="front_left_foot" pos="0.2 0.2 0" quat="0.9238795 0 0 0.3826834"/>
        </body>
      </body>
      <body name="front_right_leg" pos="0 0 0">
        <body name="aux_2" pos="-0.2 0.2 0">
          <body name="front_right_foot" pos="-0.2 0.2 0"/>
        </body>
      </body>
      <body name="left_back_leg" pos="0 0 0">
        <body name="aux_3" pos="-0.2 -0.2 0">
          <body name="left_back_foot" pos="-0.2 -0.2 0"/>
        </body>
      </body>
    </body>
  </worldbody>
</mujoco>
"""


@pytest.fixture(scope="module")
def compat():
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as hrt
    hrt.enable_compat()
    return hrt


def test_rotation_script(compat):
    from poselib.poselib.core.rotation3d import (euclidean_to_transform, quat_from_angle_axis, quat_from_rotation_matrix,
                                                 quat_inverse, quat_normalize, quat_rotate, transform_apply,
                                                 transform_from_rotation_translation, transform_inverse, transform_mul)
    q = torch.from_numpy(np.array([[0, 1, 2, 3], [-2, 3, -1, 5]], dtype=np.float32))
    r = quat_normalize(q)
    assert np.allclose(r.norm(dim=-1).numpy(), 1.0, atol=1e-6) and bool((r[:, 3] >= 0).all())
    x = torch.from_numpy(np.array([[1, 0, 0], [0, -1, 0]], dtype=np.float32))
    assert np.allclose(quat_rotate(r, x).norm(dim=-1).numpy(), 1.0, atol=1e-6)
    rng = np.random.default_rng(0)
    angle = torch.from_numpy(np.array(rng.random() * 10.0, dtype=np.float32))
    axis = torch.from_numpy(np.array([1, rng.random() * 10.0, rng.random() * 10.0], dtype=np.float32))
    rot = quat_from_angle_axis(angle, axis)
    x = torch.from_numpy(rng.random((5, 6, 3)))                       # float64 in the reference's script too
    y = quat_rotate(quat_inverse(rot), quat_rotate(rot, x))
    assert np.allclose(x.numpy(), y.numpy(), atol=1e-6)               # test_rotation.py:27-30
    m = torch.from_numpy(np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32))
    r = quat_from_rotation_matrix(m)
    assert np.allclose(r.numpy(), [0.70710678, 0, 0, 0.70710678], atol=1e-6)      # :32-33
    t = torch.from_numpy(np.array([0, 1, 0], dtype=np.float32))
    se3 = transform_from_rotation_translation(r=r, t=t)
    assert np.allclose(transform_apply(se3, t).numpy(), [0, 1, 1], atol=1e-6)     # :34-37
    rot = quat_from_angle_axis(torch.from_numpy(np.array([45, -54], dtype=np.float32)),
                               torch.from_numpy(np.array([[1, 0, 0], [0, 1, 0]], dtype=np.float32)), degree=True)
    trans = torch.from_numpy(np.array([[1, 1, 0], [1, 1, 0]], dtype=np.float32))
    transform = transform_from_rotation_translation(r=rot, t=trans)
    ident = transform_mul(transform, transform_inverse(transform))
    gt = np.zeros((2, 7))
    gt[:, 3] = 1.0                                                    # xyzw identity + zero translation (:47-51)
    assert np.allclose(ident.numpy(), gt, atol=1e-6)
    transform2 = torch.from_numpy(np.array([[1, 0, 0, 1], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32))
    transform2 = euclidean_to_transform(transform2)                   # :54-56
    assert np.allclose(transform2.numpy(), [0.70710678, 0, 0, 0.70710678, 1, 0, 0], atol=1e-6)


def test_skeleton_script(compat, tmp_path):
    from poselib.poselib.skeleton.skeleton3d import SkeletonMotion, SkeletonState, SkeletonTree
    path = os.path.join(tmp_path, "quadruped.xml")
    with open(path, "w") as f:
        f.write(ANT_LIKE_MJCF)
    skel_tree = SkeletonTree.from_mjcf(path)
    assert skel_tree.node_names[:4] == ["torso", "front_left_leg", "aux_1", "front_left_foot"]
    assert skel_tree.parent_indices.tolist() == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8]
    assert skel_tree.parent_of("aux_2") == "front_right_leg" and skel_tree.index("left_back_foot") == 9
    skel_tree_rec = SkeletonTree.from_dict(skel_tree.to_dict())
    assert skel_tree_rec.node_names == skel_tree.node_names
    assert torch.equal(skel_tree_rec.local_translation, skel_tree.local_translation)
    skel_state = SkeletonState.zero_pose(skeleton_tree=skel_tree)
    gt = skel_state.global_translation
    assert np.allclose(gt[3].numpy(), [0.4, 0.4, 0.0], atol=1e-6) and np.allclose(gt[9].numpy(), [-0.4, -0.4, 0.0], atol=1e-6)
    dropped = skel_state.drop_nodes_by_names(["aux_1", "aux_2"])
    assert len(dropped.skeleton_tree) == 8 and "aux_1" not in dropped.skeleton_tree.node_names
    # joint positions of the kept nodes are preserved (the dropped offsets are folded into their children)
    kept = [skel_tree.index(n) for n in dropped.skeleton_tree]
    assert np.allclose(dropped.global_translation.numpy(), gt[kept].numpy(), atol=1e-6)
    # with the per-node rest rotations of the MJCF (wxyz -> xyzw) and a short motion written to / read from .npy
    tree_q = SkeletonTree.from_mjcf(path, load_quat=True)
    assert np.allclose(tree_q.quat[3].numpy(), [0, 0, 0.3826834, 0.9238795], atol=1e-6)
    g = torch.Generator().manual_seed(0)
    from poselib.poselib.core.rotation3d import exp_map_to_quat
    lq = exp_map_to_quat(0.2 * torch.randn(12, 10, 3, generator=g))
    st = SkeletonState.from_rotation_and_root_translation(tree_q, lq, torch.zeros(12, 3), is_local=True)
    mot = SkeletonMotion.from_skeleton_state(st, fps=30)
    f = os.path.join(tmp_path, "m.npy")
    mot.to_file(f)
    back = SkeletonMotion.from_file(f)
    assert torch.equal(back.tensor, mot.tensor) and back.fps == 30 and back.skeleton_tree.node_names == tree_q.node_names
    # round trip local -> global -> local through the kernels
    again = st.global_repr().local_repr()
    assert float((again.rotation - st.rotation).abs().max()) <= 1e-6
