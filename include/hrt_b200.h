/* hrt_b200.h -- C ABI of the B200-native mocap -> humanoid retarget hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI: its seam is the
 * Python module boundary of retarget/spatial_transform/transform3d.py, retarget/torch_ext.py,
 * robot_kinematics_model/ and retarget/retarget_solver/.  Every entry point below names the
 * reference function whose numeric body it replaces (paths relative to the reference root).
 * The Python host layer (humanoid_real_time_retarget_b200/) binds these with ctypes and keeps
 * the reference's names and argument meanings; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  `d_` = device pointer (the caller owns the
 *     memory, e.g. a torch CUDA tensor's data_ptr), `h_` = host pointer.
 *   - all tensors are contiguous fp32 in the reference's own layouts: quaternions xyzw (.., 4),
 *     vectors (.., 3), frames on the leading axis.  Device pointers must be 16-byte aligned.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - every call returns 0 on success, a positive cudaError_t value on a CUDA failure, or a
 *     negative HRT_E_* code on an argument error; hrt_last_error_string() describes the last
 *     failure of the calling thread.  NaN in -> NaN out, like the reference.
 *   - no global mutable state: skeleton tables live in the context and travel to the kernels as
 *     __grid_constant__ parameter blocks.  A context may be used from one thread at a time;
 *     different contexts are independent.  Nothing is allocated in the hot calls except by the
 *     explicit *_host / stream entry points, which grow context-owned buffers on first use.
 */
#ifndef HRT_B200_H
#define HRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRT_ABI_VERSION 1

#define HRT_E_INVALID_ARG (-1)
#define HRT_E_NOT_CONFIGURED (-2)
#define HRT_E_UNSUPPORTED_TREE (-3)
#define HRT_E_ALIGNMENT (-4)
#define HRT_E_NO_DEVICE (-5)

#define HRT_MAX_TREES 4
/* conventional tree slots used by the Python host layer */
#define HRT_TREE_ROBOT 0      /* Hu / Hu v5 (robot_config/Hu.py, Hu_v5.py + zero-pose offsets) */
#define HRT_TREE_SOURCE 1     /* vtrdyn, 21 joints (robot_config/VTRDYN.py:33-48) */
#define HRT_TREE_SOURCE_FULL 2 /* vtrdyn_full, 59 joints (robot_config/VTRDYN_FULL.py:9-69) */

/* flags of hrt_fk_* */
#define HRT_FK_EXACT 1u       /* reference rounding order (bit-exact fp32 chain); default is FMA-fast */
/* flags of hrt_retarget_body_quat */
#define HRT_BQ_CLAMP 1u       /* clamp hinge angles to the joint limits */
#define HRT_BQ_IK 2u          /* fused damped-least-squares refinement (implies clamp) */
#define HRT_BQ_PRE_TRANSFORMED 4u /* input already zero-pose re-referenced: skip the a24 step */
#define HRT_BQ_ACTIVE_SET 16u /* refinement: freeze hinges sitting on a limit while the gradient pushes outwards (always on in the position path) */
#define HRT_BQ_PERSISTENT 32u  /* hrt_stream_open only: resident server kernel polling the mailbox (see hrt_stream_pos_open) */
#define HRT_BQ_PACKED_IK 8u   /* experimental: refinement on packed fp32x2 (FFMA2), two arms per thread; dof / link-position outputs only */

typedef struct hrt_ctx hrt_ctx;

int hrt_abi_version(void);
const char* hrt_last_error_string(void);

/* Context: owns skeleton tables, streams and (for the *_host / stream calls) pinned + device staging. */
int hrt_ctx_create(int device, hrt_ctx** out);
int hrt_ctx_destroy(hrt_ctx* ctx);
int hrt_ctx_sm_count(const hrt_ctx* ctx);

/* Install a kinematic tree.  parents[J] (-1 for the root, parents before children), offsets[J*3] =
 * zero-pose local translations (SkeletonTree.local_translation, skeleton3d.py:22-263; RobotZeroPose,
 * robot_kinematics_model/base_robot.py:24-119).  dof_axis[J-1] / lower[J-1] / upper[J-1]: per-DOF hinge
 * axis index and limits (robot_config/Hu.py:4-25, Hu_v5.py:12-33) or NULL for a mocap skeleton.
 * t2z[J*4]: T-pose -> zero-pose global quats (parse_mocap.py:71-78,97-104) or NULL. */
int hrt_set_tree(hrt_ctx* ctx, int tree, int J, const int32_t* parents, const float* offsets,
                 const uint8_t* dof_axis, const float* lower, const float* upper, const float* t2z);

/* cal_forward_kinematics (robot_kinematics_model/kinematics.py:13-39) and, through it,
 * BaseForwardModel.forward_kinematics (base_forward_model.py:13-14):
 * d_local_q (B,J,4), d_root_t (B,3) or NULL -> d_gq (B,J,4), d_gt (B,J,3) (either may be NULL). */
int hrt_fk_local_quats(hrt_ctx* ctx, int tree, int64_t B, const float* d_local_q, const float* d_root_t,
                       float* d_gq, float* d_gt, unsigned flags, void* stream);

/* HuForwardModel.forward_kinematics (robot_kinematics_model/hu_forward_model.py:17-33):
 * d_angles (B,J-1), d_root_t (B,3)|NULL, d_root_q (B,4)|NULL, clip -> joint limits with the forward
 * value of the straight-through clamp (:27-33). */
int hrt_fk_angles(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t,
                  const float* d_root_q, int clip, float* d_gq, float* d_gt, unsigned flags, void* stream);

/* Geometric Jacobian of K links (no reference counterpart):
 * d_jac (B, K, 6, J-1): rows 0-2 linear a_i x (p_k - p_i), rows 3-5 angular a_i. */
int hrt_fk_jacobian(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t,
                    const float* d_root_q, int clip, const int32_t* links, int K, float* d_jac, void* stream);

/* Backward of hrt_fk_angles: the vector-Jacobian product that lets autograd flow through
 * HuForwardModel.forward_kinematics, which is what the reference's straight-through clamp
 * `(clamp(x) - x).detach() + x` (robot_kinematics_model/hu_forward_model.py:27-33) exists for.
 * d_g_gq (B,J,4)|NULL, d_g_gt (B,J,3)|NULL: upstream gradients of the two outputs ->
 * d_g_angles (B,J-1), d_g_root_t (B,3)|NULL, d_g_root_q (B,4)|NULL.  Evaluated at the clipped angles; the clamp
 * itself passes the gradient through unchanged.  The (B,K,6,D) Jacobian is never formed. */
int hrt_fk_vjp(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t, const float* d_root_q,
               int clip, const float* d_g_gq, const float* d_g_gt, float* d_g_angles, float* d_g_root_t,
               float* d_g_root_q, void* stream);

/* Stand-alone damped-least-squares refinement of both arms of the configured robot (no reference counterpart,
 * SURVEY.md F2 / section 8(b); spec: DESIGN.md section 5).  The same device code as the stage fused into
 * hrt_retarget_body_quat / hrt_retarget_full_body_pos_ex.  Needs hrt_configure_body_quat (arm tables).
 * d_theta0 (B,2,7) warm start (clamped on entry), targets in the robot root frame: d_pe_t / d_pw_t (B,2,3)
 * elbow-pitch / wrist-yaw link positions, d_qw_t (B,2,4) wrist-yaw link orientation -> d_theta (B,2,7);
 * d_residual (B,2,iters+1)|NULL receives ||e|| before every step and after the last.  flags: HRT_BQ_ACTIVE_SET. */
int hrt_ik_refine(hrt_ctx* ctx, int64_t B, const float* d_theta0, const float* d_pe_t, const float* d_pw_t,
                  const float* d_qw_t, int iters, float damping, float rot_weight, unsigned flags, float* d_theta,
                  float* d_residual, void* stream);

/* cal_local_rotation (robot_kinematics_model/kinematics.py:41-63); equals SkeletonState.local_rotation
 * (skeleton3d.py:460-484) when tree.quat is identity. */
int hrt_local_from_global(hrt_ctx* ctx, int tree, int64_t B, const float* d_gq, float* d_lq, void* stream);

/* vtrdyn_zero_pose_transform / vtrdyn_full_zero_pose_transform (variant 0, about z) and
 * vtrdyn_broadcast_zero_pose_transform (variant 1, about x): retarget/utils/parse_mocap.py:81-89,
 * 106-114,126-134; retarget/retarget_solver/zero_pose_transform.py:33-41. */
int hrt_zero_pose_transform(hrt_ctx* ctx, int tree, int64_t B, const float* d_gq, int variant, float* d_out,
                            void* stream);

/* Wire the fused quaternion-path pipeline: which source joints feed which arm.
 * src_joints[2][5] = {torso, shoulder, upper arm, lower arm, hand} for the left then the right arm
 * (vtrdyn: {10,17,18,19,20},{10,13,14,15,16}; body_retargeter.py:40-53), rob_first[2] = robot joint index
 * of each shoulder-pitch link (Hu v5: 12, 21; body_retargeter.py:57-73).
 * HRT_E_UNSUPPORTED_TREE when the 7 arm links below a shoulder-pitch link do not form a chain with two gripper links
 * under the last.  Calls with HRT_BQ_CLAMP / HRT_BQ_IK (and hrt_ik_refine, HRT_POS_CLAMP / HRT_POS_IK on the position
 * path) additionally need every arm hinge limit of the robot tree finite and within +-pi -- the refinement's half-angle
 * sine / cosine is written for angles inside such limits -- and return HRT_E_UNSUPPORTED_TREE otherwise; the
 * reference-parity calls (no limits, no refinement) work with any table. */
int hrt_configure_body_quat(hrt_ctx* ctx, int src_tree, int rob_tree, const int32_t* src_joints,
                            const int32_t* rob_first);

/* Fused: zero-pose transform (a24) -> global-to-local (a21) -> Mocap2HuBodyRetargeter.retarget_from_pose
 * (retarget/retarget_solver/body_retargeter.py:34-81: intrinsic 'YXZ'/'ZYX' Euler split in fp64 as SciPy
 * does, transform3d.py:52-59) -> quat_to_dof_pos (transform3d.py:177-183) [-> limits -> ik_iters damped
 * least-squares steps] -> FK of the result (kinematics.py:13-39).
 * d_src_gq (B,Js,4) -> d_robot_local_q (B,Jr,4), d_dof (B,Jr-1), d_link_pos (B,Jr,3); any output may be NULL.
 * With flags == 0 the first two outputs are the reference's (robot_local_rotation, dof_pos). */
int hrt_retarget_body_quat(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters,
                           float damping, float rot_weight, float* d_robot_local_q, float* d_dof,
                           float* d_link_pos, void* stream);

/* The same call on HOST buffers (the reference-facing form: the reference works on CPU tensors).
 * Chunks the clip, overlaps H2D / kernel / D2H on three streams with context-owned device staging.
 * Pass page-locked memory for full PCIe speed; pageable memory works but is slower. Synchronous. */
int hrt_retarget_body_quat_host(hrt_ctx* ctx, int64_t B, const float* h_src_gq, unsigned flags, int ik_iters,
                                float damping, float rot_weight, float* h_robot_local_q, float* h_dof,
                                float* h_link_pos);

/* Multi-GPU reassembly fused into the kernel's store (BASELINE.json configs[4]: a clip sharded over the GPUs of one box,
 * dof_pos reassembled on every rank; SURVEY.md section 8(e) -- the reference has no multi-device code).  One process per
 * GPU.  Every rank owns a clip-wide dof buffer and a flag array, both plain device allocations exported with CUDA IPC:
 *   hrt_peer_alloc   allocate + zero `bytes` of device memory, return its 64-byte IPC handle
 *   hrt_peer_open    map another rank's allocation (enables peer access over NVLink) / hrt_peer_close
 *   hrt_retarget_body_quat_gather   hrt_retarget_body_quat whose dof rows go, as TMA bulk stores, to frame
 *                    frame0 + f of EVERY rank's buffer d_peer_dof[0..n_peer) (own buffer included); d_link_pos stays local
 *   hrt_peer_barrier after it on the same stream: publishes "my stores are done" to every rank's flag array and waits for
 *                    theirs (d_peer_flags[r] = rank r's array of HRT_MAX_PEERS words; epoch must grow by one per step) */
#define HRT_IPC_HANDLE_BYTES 64
#define HRT_MAX_PEERS 8
int hrt_peer_alloc(hrt_ctx* ctx, size_t bytes, void** d_ptr, unsigned char* handle64);
int hrt_peer_free(hrt_ctx* ctx, void* d_ptr);
int hrt_peer_open(hrt_ctx* ctx, const unsigned char* handle64, void** d_ptr);
int hrt_peer_close(hrt_ctx* ctx, void* d_ptr);
int hrt_retarget_body_quat_gather(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters,
                                  float damping, float rot_weight, float* d_link_pos, int n_peer,
                                  float* const* d_peer_dof, int64_t frame0, void* stream);
int hrt_peer_barrier(hrt_ctx* ctx, int n_peer, int my_rank, unsigned* const* d_peer_flags, unsigned epoch, void* stream);
/* The same reassembly through NVSwitch multicast (NVLS): d_mc_dof is the multicast address of the ranks' clip-wide dof
 * buffers (a CUDA multicast object with every rank's buffer bound at offset 0 and mapped; e.g. the `multicast_ptr` of a
 * torch symmetric-memory rendezvous).  Every warp publishes its dof span ONCE with multimem.st and the switch writes all
 * N copies (this rank's included): 1/N of the egress bytes of hrt_retarget_body_quat_gather, no NCCL.  Close the step
 * with hrt_peer_barrier as above. */
/* Reassembly INSIDE the compute kernel, packed wire format (the form bench.py's configs[4] line runs).  Only the 14 arm
 * hinge angles of a frame travel (the other DOFs of this solver are structurally 0).  d_symm is this rank's copy,
 * d_symm_mc the multicast address, of ONE symmetric allocation of hrt_reassembly_layout()'s total_bytes per rank,
 * zero-filled once: one 912-byte group per 16 clip frames = 16 x 14 hinge angles + a 16-byte check block {xor and
 * position-weighted sum of the 224 words, salted with the step's epoch; the epoch}.  Every warp publishes its group with
 * multimem.st (no flag, no fence: a group validates itself), and the same warps fetch the peers' groups of two rounds
 * earlier, accept those whose check block matches the words that arrived and this step's epoch, and expand them into
 * d_full_dof (n_total x D, this rank's clip-wide result; own rows are written directly) between their own rounds, so
 * transfer and unpacking hide under the solve; a group that does not check out yet is fetched again later.
 * Shards tile the clip in rank order and start at multiples of 16 frames; all ranks run the same GPU model (a group's
 * round and warp follow from the launch geometry).  epoch starts at 1 and grows by one per step; separate consecutive
 * steps with hrt_peer_barrier (a slow rank may still be unpacking the staging groups the next step overwrites).  A peer
 * that never arrives traps after 20 s.  flag_offset == total_bytes == staging_bytes (kept for ABI stability); max_rounds
 * = rounds of 16-frame groups per warp of the longest shard.  d_symm_mc == d_symm means "no multicast mapping": the
 * rank publishes nothing and only unpacks the groups already present in d_symm (replay of one rank in one process). */
int hrt_reassembly_layout(hrt_ctx* ctx, int64_t n_total, int n_rank, const int64_t* shard_frames, size_t* staging_bytes,
                          size_t* flag_offset, size_t* total_bytes, int* max_rounds);
int hrt_retarget_body_quat_reassemble(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters,
                                      float damping, float rot_weight, float* d_link_pos, float* d_full_dof,
                                      int64_t n_total, int n_rank, int my_rank, const int64_t* shard_lo,
                                      const int64_t* shard_frames, void* d_symm, void* d_symm_mc, unsigned epoch,
                                      void* stream);
int hrt_retarget_body_quat_multicast(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters,
                                     float damping, float rot_weight, float* d_link_pos, float* d_mc_dof,
                                     int64_t frame0, void* stream);

/* Position-input solvers.  mode 0: VtrdynFullBodyPosRetargeter (retarget/retarget_solver/
 * full_body_pos_retargeter.py:25-217), mode 1: HuUpperBodyFromMocapRetarget (retarget_solver.py:40-99),
 * mode 2: VtrdynFullBodyRetargeter (full_body_retargeter.py:19-177).  The joint indices those classes
 * hard-code are hard-coded here too; the zero-pose offsets come from the installed source tree (59-joint
 * vtrdyn_full for modes 0 and 2, 21-joint vtrdyn for mode 1).  src_global_t[J*3]: zero-pose global
 * translations (mode 0 gripper reference, :184); precise_gripper as in the class constructor (:21,199). */
#define HRT_POS_FULL_BODY_POS 0
#define HRT_POS_UPPER_BODY 1
#define HRT_POS_FULL_BODY 2
int hrt_configure_pos(hrt_ctx* ctx, int mode, int src_tree, int rob_tree, const float* src_global_t,
                      int precise_gripper);

/* VtrdynFullBodyPosRetargeter.retarget: d_body_t (B,21,3), d_lhand_t / d_rhand_t (B,20,3) ->
 * d_robot_local_q (B,31,4), d_dof (B,30), d_body_gq (B,59,4) (any output may be NULL).  Built from
 * cal_joint_quat (Kabsch, transform3d.py:32-50), cal_shoulderPR / cal_elbowP_and_shoulderY
 * (full_body_pos_retargeter.py:221-278), quat_in_xyz_axis 'XYZ' (transform3d.py:52-59), quat_to_dof_pos. */
int hrt_retarget_full_body_pos(hrt_ctx* ctx, int64_t B, const float* d_body_t, const float* d_lhand_t,
                               const float* d_rhand_t, float* d_robot_local_q, float* d_dof, float* d_body_gq,
                               void* stream);

/* The same with joint limits and the fused limit-aware refinement (builder-specified, no reference counterpart:
 * DESIGN.md section 5).  HRT_POS_CLAMP: the 7 hinge angles of each arm are clamped to the robot limits.
 * HRT_POS_IK (implies clamp): ik_iters damped-least-squares steps pull the clamped arm towards the elbow / wrist
 * positions and wrist orientation of the UNCLAMPED closed-form pose (a frame inside the limits is a fixed point). */
#define HRT_POS_CLAMP 1u
#define HRT_POS_IK 2u
int hrt_retarget_full_body_pos_ex(hrt_ctx* ctx, int64_t B, const float* d_body_t, const float* d_lhand_t,
                                  const float* d_rhand_t, unsigned flags, int ik_iters, float damping,
                                  float rot_weight, float* d_robot_local_q, float* d_dof, float* d_body_gq,
                                  void* stream);

/* The same call on HOST buffers (chunked, H2D / kernel / D2H overlapped on three streams; pinned memory for full PCIe
 * speed).  h_robot_local_q and h_dof may be NULL.  Synchronous. */
int hrt_retarget_full_body_pos_host(hrt_ctx* ctx, int64_t B, const float* h_body_t, const float* h_lhand_t,
                                    const float* h_rhand_t, unsigned flags, int ik_iters, float damping,
                                    float rot_weight, float* h_robot_local_q, float* h_dof);

/* HuUpperBodyFromMocapRetarget.retarget_from_global_translation: d_body_t (B,21,3), flipped by
 * coord_transform(dir=[-1,-1,1]) inside (retarget_solver.py:41) -> d_robot_local_q (B,31,4), d_dof (B,30). */
int hrt_retarget_upper_body(hrt_ctx* ctx, int64_t B, const float* d_body_t, float* d_robot_local_q, float* d_dof,
                            void* stream);

/* VtrdynFullBodyRetargeter.retarget: measured quats d_body_q (B,21,4) give the arm parents and wrists. */
int hrt_retarget_full_body(hrt_ctx* ctx, int64_t B, const float* d_body_q, const float* d_body_t,
                           const float* d_lhand_t, const float* d_rhand_t, float* d_robot_local_q, float* d_dof,
                           void* stream);

/* Streaming teleop: one frame at a time, host in -> host out (sim_full_body_teleop.py:83-129 call
 * pattern).  open() allocates mapped pinned mailboxes and captures the launch; frame() is the hot call. */
int hrt_stream_open(hrt_ctx* ctx, unsigned flags, int ik_iters, float damping, float rot_weight);
int hrt_stream_frame(hrt_ctx* ctx, const float* h_src_gq, float* h_robot_local_q, float* h_dof,
                     float* h_link_pos);
int hrt_stream_close(hrt_ctx* ctx);

/* Streaming teleop on the POSITION path (what sim_full_body_teleop.py:83-129 runs every frame):
 * VtrdynFullBodyPosRetargeter.retarget on one frame, host in -> host out through mapped pinned
 * mailboxes.  h_body_t (21*3), h_lhand_t / h_rhand_t (20*3) -> h_robot_local_q (31*4)|NULL, h_dof (30)|NULL.
 * HRT_STREAM_WIRE_LAYOUT: the inputs are in the mocap wire layout instead -- h_body_t (23*3) and hands in
 * HandNodes order -- and the 23->21 / hand reorder of sim_full_body_teleop.py:109-112 is folded into the
 * kernel's index tables.
 * HRT_STREAM_PERSISTENT: one CTA stays resident and polls a sequence number in the mailbox; frame() then
 * is two memcpys and a spin on host memory (no driver call).  The resident kernel leaves by itself after
 * 20 ms without a frame (and is relaunched by the next frame), so device-wide synchronisation elsewhere in
 * the process is delayed by at most that. */
int hrt_retarget_full_body_pos_wire(hrt_ctx* ctx, int64_t B, const float* d_body23_t, const float* d_lhand_t,
                                    const float* d_rhand_t, float* d_robot_local_q, float* d_dof, float* d_body_gq,
                                    void* stream);       /* batched form of the same wire-layout call */
#define HRT_STREAM_WIRE_LAYOUT 1   /* inputs in the mocap wire layout (23 body rows, HandNodes finger order) */
#define HRT_STREAM_PERSISTENT 2    /* resident server kernel polling the mailbox: no launch / sync per frame */
#define HRT_STREAM_CLAMP 4         /* joint limits on the arm hinges (HRT_POS_CLAMP) */
#define HRT_STREAM_IK 8            /* + 10 limit-aware refinement steps (HRT_POS_IK, damping 0.1, rotation weight 0.2) */
#define HRT_STREAM_BODY_GQ 16      /* also publish the third return of VtrdynFullBodyPosRetargeter.retarget (59 x 4 body quats) */
#define HRT_STREAM_MODE_SHIFT 8    /* bits 8-9: position solver served (0 full_body_pos, 1 upper_body, 2 full_body) */
int hrt_stream_pos_open(hrt_ctx* ctx, int flags);
int hrt_stream_pos_frame(hrt_ctx* ctx, const float* h_body_t, const float* h_lhand_t, const float* h_rhand_t,
                         float* h_robot_local_q, float* h_dof);
/* General form.  The stream serves the position solver selected at open time: flags | (mode << HRT_STREAM_MODE_SHIFT) with
 * mode 0 = VtrdynFullBodyPosRetargeter (body, hands), 1 = HuUpperBodyFromMocapRetarget (body only; hands / body_q ignored),
 * 2 = VtrdynFullBodyRetargeter (body_q (21,4), body, hands).  h_body_gq (third return of mode 0) needs HRT_STREAM_BODY_GQ.
 * Wire layout, resident server and body quaternions exist for mode 0 only. */
int hrt_stream_pos_frame_ex(hrt_ctx* ctx, const float* h_body_t, const float* h_lhand_t, const float* h_rhand_t,
                            const float* h_body_q, float* h_robot_local_q, float* h_dof, float* h_body_gq);
int hrt_stream_pos_close(hrt_ctx* ctx);

/* RetargetHuV5fromMocap (retarget/main.py:51-279), mode 3 of the position solvers: arms from joint
 * positions with BOTH arm parents = the measured/rebuilt global quat of joint 10 (main.py:203-240).
 * d_body_q (B,21,4), d_body_t (B,21,3) -> d_robot_local_q (B,31,4), d_dof (B,30). */
#define HRT_POS_MAIN 3
int hrt_retarget_main_arms(hrt_ctx* ctx, int64_t B, const float* d_body_q, const float* d_body_t,
                           float* d_robot_local_q, float* d_dof, void* stream);

/* Retarget.rescale_motion_to_standard_size (retarget/main.py:36-47) on the installed tree's offsets, with
 * coord_transform(dir) (main.py:170, transform3d.py:24-29) folded in (dir = NULL: none).
 * d_gt (B,J,3) -> d_out (B,J,3). */
int hrt_rescale_motion(hrt_ctx* ctx, int tree, int64_t B, const float* d_gt, const float* dir3, float* d_out,
                       void* stream);

/* RetargetHuV5fromMocap._rebuild_with_vtrdyn_zero_pose, rotations only (main.py:116-152): Kabsch fits
 * (cal_joint_quat) for n_kabsch <= 2 joints from 3 points each, quat_between_two_vecs per remaining bone
 * written at the parent's index, identity elsewhere.  kabsch_joint[n_kabsch], kabsch_pts[n_kabsch*3].
 * d_gt (B,J,3) -> d_out_gq (B,J,4). */
int hrt_rebuild_global_rotation(hrt_ctx* ctx, int tree, int64_t B, const float* d_gt, int n_kabsch,
                                const int32_t* kabsch_joint, const int32_t* kabsch_pts, float* d_out_gq,
                                void* stream);

/* SkeletonMotion._compute_velocity / _compute_angular_velocity (poselib/poselib/skeleton/skeleton3d.py:
 * 1126-1146): np.gradient along frames / dt, resp. axis*angle of r[t+1]*inv(r[t]) / dt, then
 * scipy.ndimage.gaussian_filter1d(sigma=2, mode="nearest") when gaussian != 0.  T frames, J joints.
 * d_scratch: T*J*3 floats.  d_gt (T,J,3) | d_gq (T,J,4) -> d_out (T,J,3). */
int hrt_motion_velocity(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gt, float dt, int gaussian,
                        float* d_scratch, float* d_out, void* stream);
int hrt_motion_angular_velocity(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gq, float dt, int gaussian,
                                float* d_scratch, float* d_out, void* stream);

/* SkeletonState.compute_forward_vector (poselib/poselib/skeleton/skeleton3d.py:542-566): facing direction per frame
 * from the shoulder and hip positions, smoothed along frames by scipy.ndimage.gaussian_filter1d(sigma,
 * mode="nearest") and normalised.  d_gt (T, J, 3) fp32; d_scratch, d_out (T, 3) fp64 (numpy promotes the cross
 * product with the integer up vector to float64 and the reference returns that).  sigma in (0, 32). */
int hrt_forward_vector(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gt, int left_shoulder, int right_shoulder,
                       int left_hip, int right_hip, double sigma, double* d_scratch, double* d_out, void* stream);

/* Element-wise rotation algebra: the free functions of poselib/poselib/core/rotation3d.py:15-661 and
 * retarget/spatial_transform/transform3d.py:9-183, one op code each (HRT_OP_*; rows AoS fp32 like the
 * reference's tensors).  Operand k has n rows when period[k] == 0, else period[k] rows that repeat
 * (broadcast of a single row or of a (J,W) table against (B,J,W)).  hrt_rot_op_info reports the row
 * widths of an op's operands and results. */
enum {
    HRT_OP_QUAT_MUL = 0,             /* rotation3d.py:15-27      (4),(4) -> (4) */
    HRT_OP_QUAT_MUL_NORM,            /* :197-202 */
    HRT_OP_QUAT_MUL_THREE,           /* :571-577 */
    HRT_OP_QUAT_MUL_FOUR,            /* :560-567 */
    HRT_OP_QUAT_POS,                 /* :31-38 */
    HRT_OP_QUAT_ABS,                 /* :42-47                   (4) -> (1) */
    HRT_OP_QUAT_UNIT,                /* :51-56 */
    HRT_OP_QUAT_NORMALIZE,           /* :93-98 */
    HRT_OP_QUAT_CONJUGATE,           /* :60-64, quat_inverse :215-219 */
    HRT_OP_QUAT_ROTATE,              /* :206-211                 (4),(3) -> (3) */
    HRT_OP_QUAT_FROM_ANGLE_AXIS,     /* :123-143                 (1),(3) -> (4); iparam: degree */
    HRT_OP_QUAT_FROM_ROTATION_MATRIX,/* :147-193                 (9) -> (4) */
    HRT_OP_QUAT_ANGLE_AXIS,          /* :231-240                 (4) -> (1),(3) */
    HRT_OP_QUAT_YAW_ROTATION,        /* :244-261                 iparam: z_up */
    HRT_OP_TRANSFORM_INVERSE,        /* :301-306                 (7) -> (7) */
    HRT_OP_TRANSFORM_MUL,            /* :318-326 */
    HRT_OP_TRANSFORM_APPLY,          /* :330-335                 (7),(3) -> (3) */
    HRT_OP_ROT_MATRIX_DET,           /* :339-350                 (9) -> (1) */
    HRT_OP_ROT_MATRIX_FROM_QUATERNION,/* :399-427                (4) -> (9) */
    HRT_OP_PROJECT_QUAT_TO_AXIS,     /* :480-530                 iparam: 0 x, 1 y, 2 z, 3 xy, 4 xz */
    HRT_OP_EXTRACT_ROTATION_ALONG_AXIS,/* :535-556               (4) -> (1); iparam: axis */
    HRT_OP_NORMALIZE_ANGLE,          /* :583-584 */
    HRT_OP_QUAT_TO_ANGLE_AXIS,       /* :588-608 */
    HRT_OP_QUAT_TO_EXP_MAP,          /* :621-627 */
    HRT_OP_EXP_MAP_TO_ANGLE_AXIS,    /* :630-646 */
    HRT_OP_EXP_MAP_TO_QUAT,          /* :649-652, transform3d.py:146-150 */
    HRT_OP_ANGLE_AXIS_TO_EXP_MAP,    /* :612-617 */
    HRT_OP_QUAT_BETWEEN_TWO_VECS,    /* transform3d.py:9-21      iparam: whole-batch early-out taken */
    HRT_OP_PROJ_IN_PLANE,            /* transform3d.py:62-75 */
    HRT_OP_RADIANS_BETWEEN_VECS,     /* transform3d.py:78-100    (3),(3),(3) -> (1) */
    HRT_OP_QUAT_SLERP,               /* transform3d.py:153-174   (4),(4),(1) -> (4) */
    HRT_OP_QUAT_TO_DOF_POS,          /* transform3d.py:177-183   (4),(1: hinge axis as float) -> (1) */
    HRT_OP_EULER_SPLIT,              /* transform3d.py:52-59     (4) -> (4),(4),(4); iparam: sequence code */
    HRT_OP_EULER_ANGLES_F64,         /* rotation3d.py:659-661    (4) -> 3 doubles; iparam: code | 0x80 degrees */
    HRT_OP_COORD_TRANSFORM,          /* transform3d.py:24-29     (3),(3: dir) -> (3); iparam: axis order code */
    HRT_OP_CAL_SHOULDER_PR,          /* retarget_solver.py:127-158  v1 (3), v0 (3), parent quat (4) -> pitch quat, roll quat */
    HRT_OP_CAL_ELBOWP_SHOULDERY,     /* retarget_solver.py:103-125  v1 (3), v0 (3), parent quat (4) -> yaw quat, elbow-pitch quat */
    HRT_OP_COUNT
};
int hrt_rot_op_info(int op, int* n_in, int* in_width4, int* n_out, int* out_width3);
int hrt_rot_op(hrt_ctx* ctx, int op, int64_t n, const float* const* d_in4, const int64_t* period4, int iparam,
               float fparam, float* const* d_out3, void* stream);

/* torch.norm(v, dim=-1).max() of (n,3) rows (transform3d.py:11), returned on the host (synchronises). */
int hrt_max_norm3(hrt_ctx* ctx, int64_t n, const float* d_v, float* h_out, void* stream);

/* cal_joint_quat (transform3d.py:32-50): Kabsch fit of n_points point pairs per row.
 * d_zero (n or zero_period rows, n_points, 3), d_motion (n, n_points, 3) -> d_out_q (n, 4). */
int hrt_cal_joint_quat(hrt_ctx* ctx, int64_t n, int n_points, const float* d_zero, int64_t zero_period,
                       const float* d_motion, float* d_out_q, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRT_B200_H */
