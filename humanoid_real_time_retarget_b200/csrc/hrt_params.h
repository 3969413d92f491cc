// Kernel parameter blocks.  They are passed BY VALUE as __grid_constant__ kernel parameters, i.e.
// they live in the constant bank: warp-uniform indexed reads are broadcast LDCs, nothing is
// global mutable state, and two contexts with different skeletons never interfere.
#pragma once
#include <stdint.h>

#define HRT_MAX_JOINTS 64
#define HRT_MAX_SLOTS 4
#define HRT_MAX_CHAIN 16
#define HRT_MAX_LINKS 4
#ifndef HRT_MAX_PEERS
#define HRT_MAX_PEERS 8
#endif

#ifdef __CUDACC__
#define HRT_HD __host__ __device__
#else
#define HRT_HD
#endif

namespace hrt {

// A kinematic tree (robot_kinematics_model/kinematics.py: parent_indices + zero-pose offsets),
// plus, for the robot, the per-DOF hinge axis and limits (retarget/robot_config/Hu*.py).
// Per-joint data is packed so that one LDC.128 + one LDC.64 fetch everything joint j needs.
struct JointRec {
    float off[3];     // zero-pose offset of joint j in its parent's frame
    // bits 0-1: hinge axis of joint j (DOF j-1); bits 4-7: 1 + slot holding the parent's
    // transform (0 = registers: parent == j-1); bits 8-11: 1 + slot to park joint j in (0 = none)
    uint32_t meta;
};
struct TreeParams {
    int J;
    int n_slots;
    int8_t parent[HRT_MAX_JOINTS];
    JointRec jr[HRT_MAX_JOINTS];
    float lim[HRT_MAX_JOINTS][2];        // lower / upper limit of joint j (DOF j-1)
};

HRT_HD inline int jr_axis(uint32_t m) { return (int)(m & 3u); }
HRT_HD inline int jr_src(uint32_t m) { return (int)((m >> 4) & 15u) - 1; }
HRT_HD inline int jr_save(uint32_t m) { return (int)((m >> 8) & 15u) - 1; }

// Limb-parallel FK schedule (built on the host by hrt_set_tree): FK_LANES lanes cooperate on one
// configuration; at step t lane p computes joint sched[t][p] (or idles).  24 bytes per entry so a
// lane fetches it with one LDS.128 + one LDS.64.
#define HRT_FK_LANES 4
#define HRT_MAX_STEPS 64
struct StepRec {
    float off[3];      // zero-pose offset of the joint
    // bits 0-7 joint index (0xFF = idle), 8-15 parent joint, 16-17 hinge axis,
    // bit 18: the parent is the joint this lane computed in the previous step (carried in registers)
    uint32_t meta;
    float lo, hi;      // joint limits
};

// Links whose Jacobian is requested: chain[k][c] = joints from the first non-root ancestor down
// to links[k] itself.
struct JacParams {
    int K;
    int link[HRT_MAX_LINKS];
    int depth[HRT_MAX_LINKS];
    int8_t chain[HRT_MAX_LINKS][HRT_MAX_CHAIN];
};

// Everything the fused quaternion-path kernel needs about one arm (side 0 = left, 1 = right).
struct ArmParams {
    int src_torso, src_shoulder, src_upper, src_lower, src_hand;   // vtrdyn joint indices
    int rob_first;                 // robot joint index of the shoulder-pitch link (12 / 21)
    float t2z[5][4];               // inv(T2Z) is applied: T2Z quats of the 5 source joints above
    float off[9][3];               // robot offsets of rob_first .. rob_first+8 (7 hinges + 2 gripper links)
    float lower[7], upper[7];
    float seg_elbow[3];            // p[elbow_pitch] - p[shoulder_pitch] at the zero pose
    float seg_wrist[3];            // p[wrist_yaw]   - p[elbow_pitch]    at the zero pose
};

struct BodyQuatParams {
    int J_src;                     // 21
    int J_rob;                     // 31
    float rot_z90[4];              // quat_from_angle_axis(pi/2, z) as the reference builds it (fp32)
    float torso_p[3];              // robot torso link position (root frame, torso angle 0)
    float shoulder_p[2][3];        // robot shoulder-pitch link positions
    ArmParams arm[2];
    // robot joints that are not on an arm keep identity rotation: position = zero-pose position
    float rest_pos[HRT_MAX_JOINTS * 3];
    float lower_all[HRT_MAX_JOINTS], upper_all[HRT_MAX_JOINTS];
    uint8_t axis_all[HRT_MAX_JOINTS];
};

}  // namespace hrt
