// Quaternion-input retarget path (BASELINE config 3q) for sm_100a.
//
// Replaces the numeric bodies of
//   vtrdyn_zero_pose_transform & friends       retarget/utils/parse_mocap.py:81-89,106-114,126-134
//   cal_local_rotation                         robot_kinematics_model/kinematics.py:41-63
//   Mocap2HuBodyRetargeter.retarget_from_pose  retarget/retarget_solver/body_retargeter.py:34-81
//   quat_in_xyz_axis (SciPy Euler split)       retarget/spatial_transform/transform3d.py:52-59
//   quat_to_dof_pos                            retarget/spatial_transform/transform3d.py:177-183
//   + joint limits, FK of the result (kinematics.py:13-39) and the fused DLS IK refinement
//     (no reference implementation; spec in DESIGN.md section 5).
#pragma once
#include "hrt_fk_limb.cuh"
#include "hrt_math.cuh"
#include "hrt_pack2.cuh"
#include "hrt_params.h"

namespace hrt {

// ---------------------------------------------------------------------------------------------
// Element-wise kernels: one thread per (frame, joint), 16-byte coalesced loads and stores.
// ---------------------------------------------------------------------------------------------
// q' = norm(norm(q * rot) * inv(T2Z[j]))            parse_mocap.py:106-114
// Four elements per thread and iteration, their loads issued first (one quaternion per thread did not keep enough bytes
// in flight: 0.76 of the HBM peak), the joint index advanced by additions instead of a 64-bit modulo per element.
constexpr int ZPT_UNROLL = 4;
__global__ void __launch_bounds__(256)
zero_pose_transform_kernel(const float4* __restrict__ gq, const float4* __restrict__ t2z, float4 rot,
                           int J, long long n_items, float4* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int j = (int)(i % J);
    const int dj = (int)(stride % J);
    for (; i < n_items; i += stride * ZPT_UNROLL) {
        float4 q[ZPT_UNROLL];
        int jj[ZPT_UNROLL];
#pragma unroll
        for (int u = 0; u < ZPT_UNROLL; ++u) {
            jj[u] = j;
            j += dj;
            if (j >= J) j -= J;
            if (i + u * stride < n_items) q[u] = __ldcs(gq + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < ZPT_UNROLL; ++u) {
            if (i + u * stride < n_items) {
                float4 r = quat_mul_norm_x(q[u], rot);
                r = quat_mul_norm_x(r, quat_conj(__ldg(t2z + jj[u])));
                __stcs(out + i + u * stride, r);
            }
        }
    }
}

// l[0] = g[0];  l[j] = norm(inv(g[parent]) * g[j])     kinematics.py:41-63
__global__ void __launch_bounds__(256)
local_from_global_kernel(const float4* __restrict__ gq, const int* __restrict__ parents, int J,
                         long long n_items, float4* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % J);
        const int p = __ldg(parents + j);
        float4 q = __ldg(gq + i);
        if (p >= 0) q = quat_mul_norm_x(quat_conj(__ldg(gq + (i - j + p))), q);
        __stcs(out + i, q);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused pipeline.  One thread per (frame, arm): lanes 2f and 2f+1 of a warp own the left and the
// right arm of frame f, so a warp owns 16 consecutive frames whose 336-byte input rows, 120-byte
// dof rows, 496-byte local-rotation rows and 372-byte link-position rows are each ONE contiguous
// span of HBM.  The input span arrives by cp.async (LDGSTS); the outputs are assembled as
// shared-memory images of their spans and leave with TMA bulk stores (cp.async.bulk, SASS UBLKCP).
// Both lanes run the same instruction stream on different table entries (no divergence).
// ---------------------------------------------------------------------------------------------
constexpr unsigned BQ_CLAMP = 1u;     // clamp hinge angles to the robot limits
constexpr unsigned BQ_IK = 2u;        // run the damped-least-squares refinement (implies clamp)
constexpr unsigned BQ_PACKED_IK = 8u; // refinement on packed fp32x2 registers, two arms per thread (experimental)
constexpr unsigned BQ_ACTIVE_SET = 16u; // freeze hinges that sit on a limit with the gradient pushing outwards

constexpr int BQ_FRAMES_PER_WARP = 16;
// One CTA per SM.  16 warps when only dof / link positions are published (the headline path),
// 8 when the 496-byte local-rotation rows are staged as well (shared-memory budget).
#ifndef HRT_BQ_WARPS_WIDE
#define HRT_BQ_WARPS_WIDE 16
#endif
constexpr int BQ_WARPS_WIDE = HRT_BQ_WARPS_WIDE;
// calls without the refinement run an instantiation with the IK loop compiled out: fewer registers, more warps per SM for
// the latency-bound closed form (exact-order libm chains, fp64 Euler split)
// (94 registers instead of 124).  Measured on 2^20 frames, dof + link positions: 0.385 ms (IK instantiation, 16 warps) ->
// 0.347 (16) / 0.333 (20) / 0.323 (24) / 0.312 ms (28 warps, 72 registers, 4 B spill).
#ifndef HRT_BQ_WARPS_NOIK
#define HRT_BQ_WARPS_NOIK 28
#endif
constexpr int BQ_WARPS_NOIK = HRT_BQ_WARPS_NOIK;
constexpr int BQ_WARPS_NARROW = 8;                // single frames and short clips
// warp-private staging, in words: [input rows, later the link-position image] [dof image]
HRT_HD inline int bq_io_words(int JS, int JR) { return BQ_FRAMES_PER_WARP * (JS * 4 > JR * 3 ? JS * 4 : JR * 3); }
HRT_HD inline int bq_dof_words(int JR) { return (BQ_FRAMES_PER_WARP * (JR - 1) + 3) / 4 * 4; }
// (the local rotations are identity except for the arm hinges: they go straight to HBM -- coalesced identity fill, then each
// lane patches its own rows -- so the tile does not depend on the outputs requested)
HRT_HD inline int bq_tile_words(int JS, int JR, bool /*with_lq*/) {
    return bq_io_words(JS, JR) + bq_dof_words(JR);
}
// CTA-shared constants: both ArmParams, the robot's rest positions, and a 16-frame image of them (the link-position
// tile of every group starts as a 16-byte-wise copy of it)
constexpr int BQ_REST_IMG_WORD = (2 * (int)sizeof(ArmParams) / 4 + HRT_MAX_JOINTS * 3 + 3) / 4 * 4;
constexpr int BQ_CONST_WORDS = BQ_REST_IMG_WORD + (BQ_FRAMES_PER_WARP * 31 * 3 + 3) / 4 * 4;

struct BodyQuatArgs {
    long long B;
    const float* __restrict__ src_gq;     // (B, 21, 4) raw mocap global quats (sensor frame)
    int pre_transformed;                  // 1: src_gq is already zero-pose re-referenced (a24 skipped)
    unsigned flags;
    int ik_iters;
    float damping;                        // lambda
    float rot_weight;                     // w_o
    float* __restrict__ out_local_q;      // (B, 31, 4) or nullptr
    float* __restrict__ out_dof;          // (B, 30)    or nullptr
    float* __restrict__ out_link_pos;     // (B, 31, 3) or nullptr
    // reassembly fused into the store (multi-GPU, configs[4]): when n_peer > 0 every warp also sends its dof span to the
    // clip-wide dof buffer of EVERY rank (peer-mapped device memory, this rank's own buffer included), at frame
    // peer_frame0 + f of that buffer -- the all-gather is n_peer bulk stores per 16 frames, no second kernel
    int n_peer;
    long long peer_frame0;
    float* peer_dof[HRT_MAX_PEERS];
    // the same with ONE store per span: mc_dof is the NVSwitch multicast address of the ranks' clip-wide buffers
    // (multimem.st: the switch replicates the store into every rank's copy, this rank's included)
    float* mc_dof;
    // reassembly INSIDE the kernel with a packed wire format (GATHER instantiation): only the 14 arm
    // hinge angles of a frame travel (the other DOFs of this solver are structurally 0); every warp publishes its 16-frame
    // group (224 words + a 4-word check block) through the multicast address of the ranks' staging buffers, and the same
    // warps fetch the peers' groups of an earlier round, accept them when the check block matches this step's epoch and
    // the words that arrived, and expand them into this rank's clip-wide dof buffer: no flags, no fences, no counters --
    // the transfer and the unpacking hide under the issue-bound solve.
    struct Gather {
        float* mc_pk;                     // multicast address of the staging buffers: ceil(n_total / 16) groups of BQ_PK_GROUP words
        const float* pk;                  // this rank's copy of it
        float* full;                      // this rank's clip-wide dof buffer (n_total, D); own rows are written directly
        long long lo[HRT_MAX_PEERS];      // first clip frame of every rank's shard (multiples of 16)
        long long n[HRT_MAX_PEERS];       // frames of every rank's shard
        int n_rank, me;
        int shard_groups;                 // 16-frame groups between consecutive ranks' shard starts when lo[p] = p * stride, else 0
        unsigned epoch;                   // grows by one per step, never 0
        unsigned long long timeout_ns;    // a lost peer traps instead of hanging the box
        // diagnostics (env HRT_GATHER_DEBUG): 1 no data stores, 2 no unpack stores, 4 no unpack loads, 16 no drain,
        // 32 every fetched group counts as valid (single-GPU cost probes with fake peers)
        unsigned debug;
    } g;
};
constexpr int BQ_PK = 14;                                                  // packed floats per frame: 2 arms x 7 hinges
constexpr int BQ_PK_WORDS = BQ_FRAMES_PER_WARP * BQ_PK;                    // 224 data words per 16-frame group
constexpr int BQ_PK_GROUP = BQ_PK_WORDS + 4;                               // + the check block: on the wire and in the staging buffers
constexpr int BQ_PK_SLOT = BQ_PK_GROUP;
// The reassembly kernel is warp-specialised: 16 compute warps (the plain pipeline + one multicast publish per round) and
// 4 unpack warps, one per scheduler, that fetch / check / expand the peers' groups.  The solve is issue-bound and uses
// the whole register file at 128 registers per thread, so the registers are re-split after the prologue (setmaxnreg):
// the CTA is launched with 640 x 96 = 61,440 registers (the most 20 warps can get) and that pool becomes 16 x 32 x 112 for
// the compute warps + 4 x 32 x 32 for the unpack warps (an .inc beyond the CTA's own pool would wait forever).  The unpack warps' instructions
// (latency-bound shared-memory / global traffic) then fill issue slots the solve leaves idle instead of sitting in
// every compute warp's instruction stream, where they cost 13-22 % (measured, profiles/r02_notes.md).
constexpr int BQ_GATHER_UNPACK_WARPS = 4;
constexpr int BQ_GATHER_COMPUTE_REGS = 112;
constexpr int BQ_GATHER_UNPACK_REGS = 32;
// a peer's group of round j is fetched once this rank's own warp has finished round j + LAG: the ranks run at the same
// pace, so the group has had LAG rounds (~40 us each) to cross the switch; a group that fails its check is fetched again
constexpr int BQ_GATHER_LAG = 1;
#ifndef HRT_GATHER_POLL_NS
#define HRT_GATHER_POLL_NS 5000
#endif
// an unpack warp's ring: one slot per peer + the 16-row image the slots are expanded through
constexpr int BQ_GATHER_RING_WORDS = ((HRT_MAX_PEERS - 1) * BQ_PK_SLOT + BQ_FRAMES_PER_WARP * 32 + 3) / 4 * 4;
// per-warp reassembly state in shared memory (ints): [0] cursor = next peer round to unpack, [1] rounds this warp has
// groups of any peer for, [2] leading rounds in which it has a FULL group of EVERY peer at an address that follows from
// the group index alone (the fast path: equal shards, 30 DOFs), [3] -; then per peer slot q (the general path: ragged
// ends, any shard layout): {source address lo, hi, destination address lo, hi (both of the warp's group of round 0),
// rounds this warp has a group of that peer for, frames of its last group}
constexpr int BQ_GATHER_STATE_HDR = 4;
constexpr int BQ_GATHER_STATE_WORDS = BQ_GATHER_STATE_HDR + 6 * (HRT_MAX_PEERS - 1) + 2;
constexpr int BQ_GATHER_FAST_D = 30;                                       // DOF count the unrolled path is written for
// shared-memory block behind the tiles: [compute warps x publish staging] [unpack warps x ring] [compute warps x state]
// [compute warps x rounds finished]
HRT_HD inline int bq_gather_ring_word(int warps) { return warps * BQ_PK_SLOT; }
HRT_HD inline int bq_gather_state_word(int warps) { return bq_gather_ring_word(warps) + BQ_GATHER_UNPACK_WARPS * BQ_GATHER_RING_WORDS; }
HRT_HD inline int bq_gather_progress_word(int warps) { return bq_gather_state_word(warps) + warps * BQ_GATHER_STATE_WORDS; }
HRT_HD inline int bq_gather_words(int warps) { return (bq_gather_progress_word(warps) + warps + 3) / 4 * 4; }
// check block of a group: {xor of the 224 words ^ salt, their sum + salt, epoch, 0}: a group is accepted when all three
// match what was fetched.  Stale groups carry the previous epoch; a group caught half-way through its arrival
// fails the sums (or holds, bit for bit, the same words as the new one).
HRT_HD inline unsigned bq_gather_salt(unsigned epoch) { return epoch * 0x9E3779B9u + 0x7F4A7C15u; }

// the arm's hinge axes (Hu_DOF_AXIS[11..17] == Hu_DOF_AXIS[20..26]); checked on the host
#define HRT_ARM_AXIS(c) ((c) == 0 ? 1 : (c) == 1 ? 0 : (c) == 2 ? 2 : (c) == 3 ? 1 : (c) == 4 ? 0 : (c) == 5 ? 1 : 2)

// one warp publishes a staged span through an NVSwitch multicast address: 16 bytes per lane per instruction, the
// switch writes every bound device's copy (PTX multimem.st; plain st on a multimem address is undefined)
HRT_DEV void warp_multimem_store_span(float* mc_dst, const float* tile, int n_words, int lane) {
    const int n4 = n_words >> 2;
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    for (int i = lane; i < n4; i += 32) {
        const float4 v = t4[i];
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(mc_dst + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
    for (int i = 4 * n4 + lane; i < n_words; i += 32)
        asm volatile("multimem.st.weak.global.f32 [%0], %1;\n" ::"l"(mc_dst + i), "f"(tile[i]) : "memory");
}

HRT_DEV void warp_store_span(float* __restrict__ dst, const float* tile, int n_words, int lane) {
    // dst is 16-byte aligned (frame-group base); vector body + scalar tail
    const int n4 = n_words >> 2;
    for (int i = lane; i < n4; i += 32)
        __stcs(reinterpret_cast<float4*>(dst) + i, *reinterpret_cast<const float4*>(tile + i * 4));
    for (int i = (n4 << 2) + lane; i < n_words; i += 32) __stcs(dst + i, tile[i]);
}

template <int C>
HRT_DEV float4 arm_local_quat(float th) {
    float s, c;
    sincos_half_lim(0.5f * th, &s, &c);
    constexpr int k = HRT_ARM_AXIS(C);
    return make_float4(k == 0 ? s : 0.f, k == 1 ? s : 0.f, k == 2 ? s : 0.f, c);
}

// 7x7 SPD solve (Cholesky), fully unrolled, A packed lower-triangular row-major: A[i*(i+1)/2 + j].
// The diagonal is kept as its reciprocal (one rsqrt per pivot, no divisions).
HRT_DEV void chol_solve7(float* A, float* b) {
    float inv[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float s = A[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
            if (i == j) inv[i] = rsqrt_fast(s);         // pivots are >= lambda^2 > 0: the bare MUFU.RSQ, no subnormal rescue
            else A[i * (i + 1) / 2 + j] = s * inv[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= A[i * (i + 1) / 2 + k] * b[k];
        b[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 6; i >= 0; --i) {
        float s = b[i];
#pragma unroll
        for (int k = i + 1; k < 7; ++k) s -= A[k * (k + 1) / 2 + i] * b[k];
        b[i] = s * inv[i];
    }
}

// the 7-hinge arm chain at angles th: world axis and position of every hinge, final orientation G (not normalised)
HRT_DEV void arm_chain_f(const float th[7], const vec3 p_sh, const float (*off)[3], vec3 ax[7], vec3 pc[7], float4& G) {
    G = make_float4(0.f, 0.f, 0.f, 1.f);
    vec3 p = p_sh;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        ax[c] = (HRT_ARM_AXIS(c) == 0) ? quat_axis_f<0>(G) : (HRT_ARM_AXIS(c) == 1) ? quat_axis_f<1>(G) : quat_axis_f<2>(G);
        pc[c] = p;
        float s, cs;
        sincos_half_lim(0.5f * th[c], &s, &cs);
        // products of unit quaternions: renormalising once, where G is used, is enough
        G = quat_mul_axis_f(G, HRT_ARM_AXIS(c), s, cs);
        if (c < 6) p = add3(p, quat_rotate_f(G, make_vec3(off[c + 1][0], off[c + 1][1], off[c + 1][2])));
    }
}

// One damped-least-squares step: e = [pe* - p_elbow; pw* - p_wrist; w_o rotvec(R* R^T)],
// dtheta = (J^T J + lambda^2 I)^-1 J^T e (7x7 Cholesky in registers), theta <- clamp(theta + dtheta).
// active_set: a hinge that sits on a limit while the gradient J^T e pushes it further out is frozen for this
// step (its row / column leave the system), which makes the clamped iteration a descent method.
//
// The chain is carried as a rotation MATRIX (three columns).  As a quaternion a hinge costs ~55 instructions (hinge
// axis = a column of R(G): 9, G * q_hinge: 8, rotating the next offset: 21, half-angle pair: 14, position add: 3); with
// the columns at hand the axis is free, a hinge turns two columns into each other (12), the offset is nine FMAs
// straight into the position, and the full-angle pair comes from the half-angle one by the double-angle formulas (4):
// ~39.  The orientation residual is read off E = H R^T (H = target orientation as a matrix, loop invariant):
// sin(phi) n = 1/2 sum_k r_k x h_k, cos(phi) = (tr E - 1) / 2, rotvec = phi n.  J^T J + lambda^2 I and J^T e are single
// FMA chains over the rows column i has (orientation rows carry the weight: axw = wo * ax).
// 930 -> 842 instructions per step; 10-step call on 2^20 frames 0.934 -> 0.846 ms; against the float64 run of the same
// spec the result is as close as the quaternion chain's was (profiles/parity_r02.json, profiles/r02_notes.md).
struct mat3c { vec3 c[3]; };                         // columns

HRT_DEV mat3c quat_to_mat3c(const float4 q) {
    mat3c m;
    m.c[0] = quat_axis_f<0>(q); m.c[1] = quat_axis_f<1>(q); m.c[2] = quat_axis_f<2>(q);
    return m;
}

// phi / sin(phi) for n = sin(phi) >= 0, w = cos(phi) of either sign (atan as in rotvec_scale_f)
HRT_DEV float angle_over_sin_f(float n, float w) {
    const float aw = fabsf(w);
    const float mx = fmaxf(n, aw), mn = fminf(n, aw);
    const float t = __fdividef(mn, mx);
    const bool big = t > 0.41421356f;
    const float u = big ? __fdividef(t - 1.f, t + 1.f) : t;
    const float z = u * u;
    float a = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * u + u;
    a = big ? a + 0.78539816339f : a;
    float phi = (n > aw) ? 1.57079637f - a : a;      // atan2(n, |w|)
    phi = (w < 0.f) ? 3.14159274f - phi : phi;
    return (n > 1e-8f) ? __fdividef(phi, n) : 1.f;
}

// the 7-hinge arm chain at angles th with the orientation carried as a matrix: world axis and position of every hinge,
// final orientation R (columns)
HRT_DEV void arm_chain_m(const float th[7], const vec3 p_sh, const float (*off)[3], vec3 ax[7], vec3 pc[7], mat3c& R) {
    vec3 r0 = make_vec3(1.f, 0.f, 0.f), r1 = make_vec3(0.f, 1.f, 0.f), r2 = make_vec3(0.f, 0.f, 1.f);
    vec3 p = p_sh;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const int k = HRT_ARM_AXIS(c);
        ax[c] = k == 0 ? r0 : (k == 1 ? r1 : r2);
        pc[c] = p;
        float sh, ch;
        sincos_half_lim(0.5f * th[c], &sh, &ch);
        const float s = (sh + sh) * ch, cs = fmaf(-2.f * sh, sh, 1.f);
        // R <- R * Rot_k(theta): the two columns after k (cyclically) turn into each other
        vec3& ca = k == 0 ? r1 : (k == 1 ? r2 : r0);
        vec3& cb = k == 0 ? r2 : (k == 1 ? r0 : r1);
        const vec3 na = make_vec3(fmaf(s, cb.x, cs * ca.x), fmaf(s, cb.y, cs * ca.y), fmaf(s, cb.z, cs * ca.z));
        const vec3 nb = make_vec3(fmaf(-s, ca.x, cs * cb.x), fmaf(-s, ca.y, cs * cb.y), fmaf(-s, ca.z, cs * cb.z));
        ca = na; cb = nb;
        if (c < 6) {
            const float ox = off[c + 1][0], oy = off[c + 1][1], oz = off[c + 1][2];
            p.x = fmaf(r0.x, ox, fmaf(r1.x, oy, fmaf(r2.x, oz, p.x)));
            p.y = fmaf(r0.y, ox, fmaf(r1.y, oy, fmaf(r2.y, oz, p.y)));
            p.z = fmaf(r0.z, ox, fmaf(r1.z, oy, fmaf(r2.z, oz, p.z)));
        }
    }
    R.c[0] = r0; R.c[1] = r1; R.c[2] = r2;
}

HRT_DEV void ik_step_f(float th[7], const vec3 p_sh, const float (*off)[3], const float* lower, const float* upper,
                       const vec3 pe_t, const vec3 pw_t, const mat3c& H, const float lam2, const float wo, const bool active_set) {
    vec3 ax[7], pc[7];
    mat3c R;
    arm_chain_m(th, p_sh, off, ax, pc, R);
    const vec3 r0 = R.c[0], r1 = R.c[1], r2 = R.c[2];
    float e[9];
    e[0] = pe_t.x - pc[3].x; e[1] = pe_t.y - pc[3].y; e[2] = pe_t.z - pc[3].z;
    e[3] = pw_t.x - pc[6].x; e[4] = pw_t.y - pc[6].y; e[5] = pw_t.z - pc[6].z;
    {
        // twice the axial vector of E = H R^T and its trace
        vec3 v = cross3_f(r0, H.c[0]);
        v.x = fmaf(r1.y, H.c[1].z, fmaf(-r1.z, H.c[1].y, v.x)); v.y = fmaf(r1.z, H.c[1].x, fmaf(-r1.x, H.c[1].z, v.y)); v.z = fmaf(r1.x, H.c[1].y, fmaf(-r1.y, H.c[1].x, v.z));
        v.x = fmaf(r2.y, H.c[2].z, fmaf(-r2.z, H.c[2].y, v.x)); v.y = fmaf(r2.z, H.c[2].x, fmaf(-r2.x, H.c[2].z, v.y)); v.z = fmaf(r2.x, H.c[2].y, fmaf(-r2.y, H.c[2].x, v.z));
        float tr = r0.x * H.c[0].x;
        tr = fmaf(r0.y, H.c[0].y, tr); tr = fmaf(r0.z, H.c[0].z, tr);
        tr = fmaf(r1.x, H.c[1].x, tr); tr = fmaf(r1.y, H.c[1].y, tr); tr = fmaf(r1.z, H.c[1].z, tr);
        tr = fmaf(r2.x, H.c[2].x, tr); tr = fmaf(r2.y, H.c[2].y, tr); tr = fmaf(r2.z, H.c[2].z, tr);
        const float n2 = 0.25f * (v.x * v.x + v.y * v.y + v.z * v.z);
        const float n = n2 > 1e-30f ? n2 * rsqrt_fast(n2) : 0.f;       // sin(phi)
        const float sc = 0.5f * wo * angle_over_sin_f(n, 0.5f * (tr - 1.f));
        e[6] = v.x * sc; e[7] = v.y * sc; e[8] = v.z * sc;
    }
    vec3 je[3], jw[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) je[c] = cross3_f(ax[c], sub3(pc[3], pc[c]));
#pragma unroll
    for (int c = 0; c < 6; ++c) jw[c] = cross3_f(ax[c], sub3(pc[6], pc[c]));
    vec3 axw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) axw[i] = make_vec3(wo * ax[i].x, wo * ax[i].y, wo * ax[i].z);
    float A[28], g[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        float gi = axw[i].x * e[6];
        gi = fmaf(axw[i].y, e[7], gi); gi = fmaf(axw[i].z, e[8], gi);
        if (i < 6) { gi = fmaf(jw[i].x, e[3], gi); gi = fmaf(jw[i].y, e[4], gi); gi = fmaf(jw[i].z, e[5], gi); }
        if (i < 3) { gi = fmaf(je[i].x, e[0], gi); gi = fmaf(je[i].y, e[1], gi); gi = fmaf(je[i].z, e[2], gi); }
        g[i] = gi;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float s = (i == j) ? fmaf(axw[i].x, axw[j].x, lam2) : axw[i].x * axw[j].x;
            s = fmaf(axw[i].y, axw[j].y, s); s = fmaf(axw[i].z, axw[j].z, s);
            if (i < 6) { s = fmaf(jw[i].x, jw[j].x, s); s = fmaf(jw[i].y, jw[j].y, s); s = fmaf(jw[i].z, jw[j].z, s); }
            if (i < 3) { s = fmaf(je[i].x, je[j].x, s); s = fmaf(je[i].y, je[j].y, s); s = fmaf(je[i].z, je[j].z, s); }
            A[i * (i + 1) / 2 + j] = s;
        }
    }
    if (active_set) {
        float m[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const bool blocked = (th[i] >= upper[i] && g[i] > 0.f) || (th[i] <= lower[i] && g[i] < 0.f);
            m[i] = blocked ? 0.f : 1.f;
            g[i] *= m[i];
        }
#pragma unroll
        for (int i = 1; i < 7; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j) A[i * (i + 1) / 2 + j] *= m[i] * m[j];
    }
    chol_solve7(A, g);
#pragma unroll
    for (int c = 0; c < 7; ++c) th[c] = fminf(fmaxf(th[c] + g[c], lower[c]), upper[c]);
}

// Warps that share a scheduler (warp % 4) re-align at the top of every IK iteration with a named
// barrier: the unrolled iteration body is ~16 KB of SASS, far more than an SMSP's L0 instruction
// cache, and ncu showed "no instruction" as the top stall (2.4 warps per issue) when the warps
// of a scheduler drifted apart and each streamed the body on its own (profiles/r01_body_quat.md).
#ifndef HRT_BQ_ALIGN
#define HRT_BQ_ALIGN 1
#endif
#ifndef HRT_BQ_ALIGN_EULER
#define HRT_BQ_ALIGN_EULER 1
#endif

template <int WARPS>
HRT_DEV void smsp_align(int warp) {
#if HRT_BQ_ALIGN
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + (warp & 3)), "n"(WARPS * 8) : "memory");
#endif
}
// per-arm tables and rest positions: constant bank -> shared memory once per CTA (the two lanes of a frame
// index different arms, which would serialise every constant-bank read)
HRT_DEV void bq_setup(const BodyQuatParams& bp, float* smem) {
    const float* src = reinterpret_cast<const float*>(&bp.arm[0]);
    for (int i = threadIdx.x; i < 2 * (int)sizeof(ArmParams) / 4; i += blockDim.x) smem[i] = src[i];
    float* rp = smem + 2 * sizeof(ArmParams) / 4;
    for (int i = threadIdx.x; i < bp.J_rob * 3; i += blockDim.x) rp[i] = bp.rest_pos[i];
    const int W = bp.J_rob * 3;
    for (int i = threadIdx.x; i < BQ_FRAMES_PER_WARP * W; i += blockDim.x) smem[BQ_REST_IMG_WORD + i] = bp.rest_pos[i % W];
    __syncthreads();
}

// SYSMEM: resident one-warp server reading a frame from mapped host memory (no cross-warp alignment there)
template <int WARPS, bool SYSMEM>
HRT_DEV void bq_align(int warp) {
    if (!SYSMEM) smsp_align<WARPS>(warp);
}

// all frame groups of `a` that fall to CTA `cta` of `n_ctas`
HRT_DEV unsigned long long bq_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}
// ---------------------------------------------------------------------------------------------
// In-kernel reassembly (body_quat_gather_kernel).
//   * a compute warp, at the end of a round: writes its 14 hinge angles per frame and the group's check block to every
//     rank through the multicast address (plain multimem.st: nothing orders them, the check block is what makes a group
//     self-validating) and bumps its round counter in shared memory;
//   * an unpack warp (one per scheduler, serving that scheduler's four compute warps): once a compute warp has finished
//     round j + LAG it cp.async's that warp's group of round j of EVERY peer (staging groups, 912 B each) into its ring,
//     checks every slot against its check block and, when all of them are this step's, turns each into a full 16-row
//     dof image (lane = (frame, arm) drops its 7 hinge angles into an image whose other words stay zero) and copies the
//     image out with 128-bit stores; a round that does not check out yet is fetched again after a short sleep.
// ---------------------------------------------------------------------------------------------
HRT_DEV void gather_setup(const BodyQuatArgs::Gather& g, float* gbase, int warps, int n_ctas, int cta, int D) {
    int* state = reinterpret_cast<int*>(gbase + bq_gather_state_word(warps));
    if (threadIdx.x < warps) {
        reinterpret_cast<int*>(gbase + bq_gather_progress_word(warps))[threadIdx.x] = 0;
        int* st = state + threadIdx.x * BQ_GATHER_STATE_WORDS;
        const long long T = (long long)n_ctas * warps, w0 = (long long)cta * warps + threadIdx.x;
        int peer_rounds = 0;
        for (int q = 0; q < g.n_rank - 1; ++q) {
            const int p = q + (q >= g.me ? 1 : 0);
            const long long gp = (g.n[p] + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
            const int rq = gp > w0 ? (int)((gp - w0 + T - 1) / T) : 0;
            peer_rounds = max(peer_rounds, rq);
            const unsigned long long src = reinterpret_cast<unsigned long long>(g.pk + (g.lo[p] / BQ_FRAMES_PER_WARP + w0) * BQ_PK_GROUP);
            const unsigned long long dst = reinterpret_cast<unsigned long long>(g.full + (g.lo[p] + w0 * BQ_FRAMES_PER_WARP) * D);
            const long long last_f = ((long long)(rq - 1) * T + w0) * BQ_FRAMES_PER_WARP;
            st[BQ_GATHER_STATE_HDR + 6 * q + 0] = (int)(unsigned)src; st[BQ_GATHER_STATE_HDR + 6 * q + 1] = (int)(unsigned)(src >> 32);
            st[BQ_GATHER_STATE_HDR + 6 * q + 2] = (int)(unsigned)dst; st[BQ_GATHER_STATE_HDR + 6 * q + 3] = (int)(unsigned)(dst >> 32);
            st[BQ_GATHER_STATE_HDR + 6 * q + 4] = rq;
            st[BQ_GATHER_STATE_HDR + 6 * q + 5] = rq > 0 ? (int)min((long long)BQ_FRAMES_PER_WARP, g.n[p] - last_f) : 0;
        }
        // fast path: as long as every peer has a whole group for this warp, at a group index that is affine in the rank
        int rq_full = 0;
        if (g.shard_groups > 0 && D == BQ_GATHER_FAST_D && g.n_rank > 1) {
            rq_full = 0x7fffffff;
            for (int p = 0; p < g.n_rank; ++p) {
                if (p == g.me) continue;
                const long long gf = g.n[p] / BQ_FRAMES_PER_WARP;                     // whole groups of that shard
                rq_full = min(rq_full, gf > w0 ? (int)((gf - w0 + T - 1) / T) : 0);
                if (g.lo[p] != (long long)p * g.shard_groups * BQ_FRAMES_PER_WARP) rq_full = 0;
            }
        }
        st[0] = 0;
        st[1] = peer_rounds;
        st[2] = rq_full;
        st[3] = 0;
    }
    __syncthreads();
}

// ---- fast path (state[2]): unrolled over the peer slots, addresses from the group index, no per-peer state ----------
// The three passes over the peers' slots are rolled loops on purpose: the unpack warps run on 32 registers and share
// their scheduler's instruction fetch with four issue-bound compute warps, so their code is kept short (a few dozen
// instructions per peer) and their addresses advance by additions.
// g_rel = peer round * (compute warps of the grid) + the compute warp's number
HRT_DEV void gather_fetch_fast(const BodyQuatArgs::Gather& g, float* ring, unsigned g_rel, int lane) {
    const unsigned long long stride = (unsigned long long)(unsigned)g.shard_groups * (unsigned long long)(BQ_PK_GROUP * 4);
    const char* src = reinterpret_cast<const char*>(g.pk) + (unsigned long long)g_rel * (unsigned long long)(BQ_PK_GROUP * 4) + lane * 16;
    float* slot = ring + lane * 4;
    const int np = g.n_rank - 1;
#pragma unroll 1
    for (int q = 0; q < np; ++q) {
        if (q == g.me) src += stride;                       // this rank's own shard is skipped
        cp_async16(slot, src);
        if (lane < BQ_PK_GROUP / 4 - 32) cp_async16(slot + 128, src + 512);
        src += stride;
        slot += BQ_PK_SLOT;
    }
}

HRT_DEV bool gather_validate_fast(const BodyQuatArgs::Gather& g, const float* ring, int lane) {
    const int np = g.n_rank - 1;
    const unsigned salt = bq_gather_salt(g.epoch);
    unsigned bad = 0u;
    const unsigned* slot = reinterpret_cast<const unsigned*>(ring);
#pragma unroll 1
    for (int q = 0; q < np; ++q) {
        const unsigned* mine = slot + lane * 7;
        unsigned x = mine[0] ^ mine[1] ^ mine[2], y = mine[0] + mine[1] + mine[2];
        x ^= mine[3] ^ mine[4]; y += mine[3] + mine[4];
        x ^= mine[5] ^ mine[6]; y += mine[5] + mine[6];
        x = __reduce_xor_sync(0xffffffffu, x);
        y = __reduce_add_sync(0xffffffffu, y);
        const uint4 chk = *reinterpret_cast<const uint4*>(slot + BQ_PK_WORDS);
        bad |= (chk.x ^ x ^ salt) | (chk.y - y - salt) | (chk.z ^ g.epoch);
        slot += BQ_PK_SLOT;
    }
    // one verdict for the warp (every lane read the check blocks on its own), and every lane is done reading the slots
    // before any lane refetches them
    bad = __reduce_or_sync(0xffffffffu, bad);
    return bad == 0u || (g.debug & 32u) != 0;
}

// every slot of an accepted fast round -> 16 full dof rows each: lane = (frame, arm) drops its 7 hinge angles into the
// image (`col`: first hinge column of the lane's arm), the image leaves with 128-bit stores
HRT_DEV void gather_store_fast(const BodyQuatArgs::Gather& g, const float* ring, float* img, unsigned g_rel, int lane, int col) {
    if (g.debug & 2u) return;
    constexpr int N4 = BQ_FRAMES_PER_WARP * BQ_GATHER_FAST_D / 4;           // 120 sixteen-byte blocks per image
    const unsigned long long stride = (unsigned long long)(unsigned)g.shard_groups * (unsigned long long)(N4 * 16);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<char*>(g.full) + (unsigned long long)g_rel * (unsigned long long)(N4 * 16)) + lane;
    const float* mine = ring + lane * 7;
    float* row = img + (lane >> 1) * BQ_GATHER_FAST_D + col;
    const float4* img4 = reinterpret_cast<const float4*>(img) + lane;
    const int np = g.n_rank - 1;
#pragma unroll 1
    for (int q = 0; q < np; ++q) {
        if (q == g.me) dst = reinterpret_cast<float4*>(reinterpret_cast<char*>(dst) + stride);
        __syncwarp();                                       // the image's previous copy-out (or its zero fill) is done
#pragma unroll
        for (int c = 0; c < 7; ++c) row[c] = mine[c];
        __syncwarp();
#pragma unroll
        for (int t = 0; t < (N4 + 31) / 32; ++t)
            if (t < N4 / 32 || lane < N4 % 32) __stcs(dst + 32 * t, img4[32 * t]);
        dst = reinterpret_cast<float4*>(reinterpret_cast<char*>(dst) + stride);
        mine += BQ_PK_SLOT;
    }
}

// the image the slots are expanded through: 16 x D words whose non-hinge words stay zero
HRT_DEV void gather_image_clear(float* img, int D, int lane) {
    float4* img4 = reinterpret_cast<float4*>(img);
    for (int i = lane; i < BQ_FRAMES_PER_WARP * D / 4; i += 32) img4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- general path (ragged last groups, unequal shards, other DOF counts): per-peer state, one peer at a time -------
// Run by the COMPUTE warps after their last round (they have the registers, and it is a handful of groups per clip):
// peer slot q's group of peer round `j` that belongs to this warp is fetched into `slot`, checked, expanded through
// `img` (cleared by the caller) and stored; repeated until the group checks out.
HRT_DEV void gather_tail_one(const BodyQuatArgs::Gather& g, const int* st, float* slot, float* img, int warps, int n_ctas, int lane,
                             int D, int col, int j, int q) {
    const int rq = st[BQ_GATHER_STATE_HDR + 6 * q + 4];
    if (j >= rq) return;
    const unsigned long long soff = (unsigned long long)j * (unsigned long long)((unsigned)(n_ctas * warps) * (unsigned)(BQ_PK_GROUP * 4)) + (unsigned)lane * 16u;
    const char* src = reinterpret_cast<const char*>(((unsigned long long)(unsigned)st[BQ_GATHER_STATE_HDR + 6 * q + 1] << 32 | (unsigned)st[BQ_GATHER_STATE_HDR + 6 * q]) + soff);
    const unsigned salt = bq_gather_salt(g.epoch);
    const unsigned* us = reinterpret_cast<const unsigned*>(slot);
    const unsigned long long t0 = bq_timer_ns();
    for (;;) {
        if (!(g.debug & 4u)) {
            cp_async16(slot + lane * 4, src);
            if (lane < BQ_PK_GROUP / 4 - 32) cp_async16(slot + lane * 4 + 128, src + 512);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        const unsigned* mine = us + lane * 7;
        unsigned x = mine[0] ^ mine[1] ^ mine[2], y = mine[0] + mine[1] + mine[2];
        x ^= mine[3] ^ mine[4]; y += mine[3] + mine[4];
        x ^= mine[5] ^ mine[6]; y += mine[5] + mine[6];
        x = __reduce_xor_sync(0xffffffffu, x);
        y = __reduce_add_sync(0xffffffffu, y);
        const bool good = us[BQ_PK_WORDS] == (x ^ salt) && us[BQ_PK_WORDS + 1] == y + salt && us[BQ_PK_WORDS + 2] == g.epoch;
        if (__all_sync(0xffffffffu, good) || (g.debug & 32u)) break;                     // one verdict for the warp
        if (bq_timer_ns() - t0 > g.timeout_ns) __trap();        // a lost peer must not hang the box
        __nanosleep(2000);
        __syncwarp();
    }
    if (g.debug & 2u) return;
    float* row = img + (lane >> 1) * D + col;
#pragma unroll
    for (int c = 0; c < 7; ++c) row[c] = slot[lane * 7 + c];
    __syncwarp();
    const unsigned long long doff = (unsigned long long)j * (unsigned long long)((unsigned)(n_ctas * warps) * (unsigned)(BQ_FRAMES_PER_WARP * 4) * (unsigned)D);
    float* dst = reinterpret_cast<float*>(((unsigned long long)(unsigned)st[BQ_GATHER_STATE_HDR + 6 * q + 3] << 32 | (unsigned)st[BQ_GATHER_STATE_HDR + 6 * q + 2]) + doff);
    const int cnt = (j == rq - 1) ? st[BQ_GATHER_STATE_HDR + 6 * q + 5] : BQ_FRAMES_PER_WARP;
    for (int i = lane; i < cnt * D; i += 32) __stcs(dst + i, img[i]);          // a ragged group stores its own rows only
    __syncwarp();
}

// one (peer round, compute warp) item of an unpack warp: fetch, check (again until it checks out), expand
HRT_DEV void gather_item_fast(const BodyQuatArgs::Gather& g, float* ring, float* img, unsigned g_rel, int lane, int col) {
    unsigned long long t0 = 0ull;
    for (;;) {
        if (!(g.debug & 4u)) gather_fetch_fast(g, ring, g_rel, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (gather_validate_fast(g, ring, lane)) break;
        const unsigned long long now = bq_timer_ns();
        if (t0 == 0ull) t0 = now;
        if (now - t0 > g.timeout_ns) __trap();                  // a lost peer must not hang the box
        __nanosleep(2000);
    }
    gather_store_fast(g, ring, img, g_rel, lane, col);
    __syncwarp();                                               // the slots may be overwritten
}

// One unpack warp (uw = 0 .. 3): the peers' groups of the fast rounds of the compute warps of its own scheduler.
template <int WARPS>
HRT_DEV void gather_unpack_warp(const BodyQuatArgs::Gather& g, float* gbase, int uw, int lane, int n_ctas, int cta, int col) {
    float* ring = gbase + bq_gather_ring_word(WARPS) + uw * BQ_GATHER_RING_WORDS;
    float* img = ring + (HRT_MAX_PEERS - 1) * BQ_PK_SLOT;
    const int* state = reinterpret_cast<const int*>(gbase + bq_gather_state_word(WARPS)) + uw * BQ_GATHER_STATE_WORDS;
    const volatile int* progress = reinterpret_cast<const volatile int*>(gbase + bq_gather_progress_word(WARPS)) + uw;
    const unsigned T = (unsigned)(n_ctas * WARPS);
    gather_image_clear(img, BQ_GATHER_FAST_D, lane);           // the zero words stay: only hinge words are ever rewritten
    __syncwarp();
    int max_rounds = 0;
#pragma unroll
    for (int k = 0; k < WARPS / BQ_GATHER_UNPACK_WARPS; ++k) max_rounds = max(max_rounds, state[BQ_GATHER_UNPACK_WARPS * k * BQ_GATHER_STATE_WORDS + 2]);
    if (g.debug & 16u) max_rounds = 0;
#pragma unroll 1
    for (int j = 0; j < max_rounds; ++j) {
        // pacing: the peers run at this rank's pace, and the four compute warps of a scheduler move in lockstep: one poll
        // per round on the first of them (a finished compute warp reports INT_MAX); correctness rests on the check blocks
        while (*progress < j + 1 + BQ_GATHER_LAG) __nanosleep(HRT_GATHER_POLL_NS);
#pragma unroll 1
        for (int k = 0; k < WARPS / BQ_GATHER_UNPACK_WARPS; ++k) {
            const int rq_full = state[BQ_GATHER_UNPACK_WARPS * k * BQ_GATHER_STATE_WORDS + 2];
            if (j >= rq_full) continue;
            const unsigned g_rel = (unsigned)j * T + (unsigned)(cta * WARPS + uw + BQ_GATHER_UNPACK_WARPS * k);
            gather_item_fast(g, ring, img, g_rel, lane, col);
        }
    }
}

// End of a round: the warp's packed hinge angles (lane = (frame, arm): words 7 lane .. 7 lane + 6) and the group's check
// block go to every rank through the multicast address.
HRT_DEV void gather_publish(const BodyQuatArgs::Gather& g, float* gbase, int warp, int lane, const float th[7], int nfr, long long f0) {
    if (nfr <= 0) return;
    float* pk_t = gbase + warp * BQ_PK_SLOT;                                 // the compute warp's own publish staging
    unsigned x = 0u, y = 0u;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const float v = (lane >> 1) < nfr ? th[c] : 0.f;                    // a ragged group travels whole, zero padded
        pk_t[lane * 7 + c] = v;
        x ^= __float_as_uint(v);
        y += __float_as_uint(v);
    }
    x = __reduce_xor_sync(0xffffffffu, x);
    y = __reduce_add_sync(0xffffffffu, y);
    if (lane == 0) {
        const unsigned salt = bq_gather_salt(g.epoch);
        unsigned* chk = reinterpret_cast<unsigned*>(pk_t + BQ_PK_WORDS);
        chk[0] = x ^ salt; chk[1] = y + salt; chk[2] = g.epoch; chk[3] = 0u;
    }
    __syncwarp();
    if (!(g.debug & 1u))
        warp_multimem_store_span(g.mc_pk + ((g.lo[g.me] + f0) / BQ_FRAMES_PER_WARP) * BQ_PK_GROUP, pk_t, BQ_PK_GROUP, lane);
    __syncwarp();
}

template <int BQ_WARPS_PER_CTA, bool SYSMEM, bool WITH_IK = true, bool GATHER = false>
HRT_DEV void bq_process(const BodyQuatParams& bp, const BodyQuatArgs& a, float* smem, int n_ctas, int cta) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fl = lane >> 1;
    const int side = lane & 1;
    const ArmParams& ap = reinterpret_cast<const ArmParams*>(smem)[side];
    const bool with_lq = a.out_local_q != nullptr;
    float* tile = smem + BQ_CONST_WORDS + warp * bq_tile_words(bp.J_src, bp.J_rob, with_lq);
    float* lp_t = tile;                                   // input rows, later the link-position image
    float* dof_t = tile + bq_io_words(bp.J_src, bp.J_rob);
    bool pending_store = false;
    const long long n_groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    const int JS = bp.J_src, JR = bp.J_rob;
    const int D = JR - 1, W = JR * 3;
    const bool do_clamp = (a.flags & (BQ_CLAMP | BQ_IK)) != 0;
    const bool do_ik = (a.flags & BQ_IK) != 0;
    const vec3 p_sh = make_vec3(bp.shoulder_p[side][0], bp.shoulder_p[side][1], bp.shoulder_p[side][2]);

    // in-kernel reassembly (body_quat_gather_kernel): the compute warps only publish; gbase = the block behind the tiles
    float* gbase = smem + BQ_CONST_WORDS + BQ_WARPS_PER_CTA * bq_tile_words(bp.J_src, bp.J_rob, with_lq);

    // every warp of the CTA runs the same number of rounds (the named barriers below need that);
    // a warp without a group in the last round shadows the last group and publishes nothing
    const long long total_warps = (long long)n_ctas * BQ_WARPS_PER_CTA;
    const long long rounds = (n_groups + total_warps - 1) / total_warps;
    for (long long rnd = 0; rnd < rounds; ++rnd) {
        const long long grp_raw = rnd * total_warps + (long long)cta * BQ_WARPS_PER_CTA + warp;
        const bool live = grp_raw < n_groups;
        const long long grp = live ? grp_raw : n_groups - 1;
        const long long f0 = grp * BQ_FRAMES_PER_WARP;
        const int nfr = live ? (int)min((long long)BQ_FRAMES_PER_WARP, a.B - f0) : 0;
        const int nld = (int)min((long long)BQ_FRAMES_PER_WARP, a.B - f0);   // rows to load
        const int fr = min(fl, nld - 1);          // tail lanes recompute the last valid frame

        // the previous group's bulk stores must have finished reading the staging tiles
        if (pending_store) {
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            pending_store = false;
        }
        // ---- 1. stage the group's input rows (one contiguous span) ----------------------------
        if (SYSMEM) {
            // one frame in mapped host memory: all PCIe reads in flight together, caches bypassed
            const int n_words = nld * JS * 4;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (lane + 32 * k < n_words) ? __ldcv(a.src_gq + f0 * JS * 4 + lane + 32 * k) : 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (lane + 32 * k < n_words) tile[lane + 32 * k] = v[k];
        } else {
            warp_span_g2s(tile, a.src_gq + f0 * JS * 4, nld * JS * 4, lane);
        }
        cp_async_commit();
        // the next round's rows start their way from HBM to L2 now (no room for a second staging buffer)
        if (!SYSMEM && lane == 0 && grp_raw + total_warps < n_groups) {
            const long long fn = (grp_raw + total_warps) * BQ_FRAMES_PER_WARP;
            l2_prefetch_span(a.src_gq + fn * JS * 4, (int)min((long long)BQ_FRAMES_PER_WARP, a.B - fn) * JS * 4);
        }
        // while the copy is in flight: pre-fill the output images with their constant parts
        const bool want_dof = a.out_dof != nullptr || a.n_peer > 0 || a.mc_dof != nullptr;
        if (want_dof) for (int i = lane; i < nfr * D; i += 32) dof_t[i] = 0.f;
        cp_async_wait<0>();
        __syncwarp();
        const float* row = tile + fr * JS * 4;
        float4 zT = *reinterpret_cast<const float4*>(row + ap.src_torso * 4);
        float4 zS = *reinterpret_cast<const float4*>(row + ap.src_shoulder * 4);
        float4 zU = *reinterpret_cast<const float4*>(row + ap.src_upper * 4);
        float4 zL = *reinterpret_cast<const float4*>(row + ap.src_lower * 4);
        float4 zH = *reinterpret_cast<const float4*>(row + ap.src_hand * 4);
        __syncwarp();                              // the input rows may now be overwritten:
        if (a.out_link_pos) {                      // link-position image starts as the robot's rest pose
            const float* img = smem + BQ_REST_IMG_WORD;
            const int n4 = (nfr * W) >> 2;
            for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(lp_t)[i] = reinterpret_cast<const float4*>(img)[i];
            for (int i = (n4 << 2) + lane; i < nfr * W; i += 32) lp_t[i] = img[i];
        }

        // ---- 2. zero-pose re-referencing (a24), exact rounding order -------------------------
        bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
        if (!a.pre_transformed) {
            const float4 rot = make_float4(bp.rot_z90[0], bp.rot_z90[1], bp.rot_z90[2], bp.rot_z90[3]);
#define HRT_ZPT(q, n) q = quat_mul_norm_x(quat_mul_norm_x(q, rot), \
            make_float4(-ap.t2z[n][0], -ap.t2z[n][1], -ap.t2z[n][2], ap.t2z[n][3]))
            HRT_ZPT(zT, 0); HRT_ZPT(zS, 1); HRT_ZPT(zU, 2); HRT_ZPT(zL, 3); HRT_ZPT(zH, 4);
#undef HRT_ZPT
        }
        // ---- 3. global -> local (a21) for the upper and the lower arm ------------------------
        const float4 lU = quat_mul_norm_x(quat_conj(zS), zU);
        const float4 lL = quat_mul_norm_x(quat_conj(zU), zL);

        // ---- 4. intrinsic Euler splits in fp64 (a16) + joint mapping (a30) --------------------
        float4 rl[7];
        {
            double sS[3], cS[3], sE[3], cE[3];
            bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
            euler_intrinsic_half_sincos_f64<1, 0, 2>(lU, sS, cS);      // 'YXZ': pitch, roll, yaw
#if HRT_BQ_ALIGN_EULER
            bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
#endif
            euler_intrinsic_half_sincos_f64<2, 1, 0>(lL, sE, cE);      // 'ZYX': yaw, pitch, roll
            bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
            rl[0] = axis_quat_from_sc(sS[0], cS[0], 1);
            rl[1] = axis_quat_from_sc(sS[1], cS[1], 0);
            rl[2] = quat_mul_norm_x(axis_quat_from_sc(sE[0], cE[0], 2), axis_quat_from_sc(sS[2], cS[2], 2));
            rl[3] = axis_quat_from_sc(sE[1], cE[1], 1);
            rl[4] = axis_quat_from_sc(sE[2], cE[2], 0);
            rl[5] = make_float4(0.f, 0.f, 0.f, 1.f);
            rl[6] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        // ---- 5. hinge angles (a17) ------------------------------------------------------------
        bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
        float th[7];
        th[0] = quat_to_dof_x(rl[0], 1);
        th[1] = quat_to_dof_x(rl[1], 0);
        th[2] = quat_to_dof_x(rl[2], 2);
        th[3] = quat_to_dof_x(rl[3], 1);
        th[4] = quat_to_dof_x(rl[4], 0);
        th[5] = 0.f;
        th[6] = 0.f;

        // fminf / fmaxf swallow NaN: remember a poisoned frame so that "NaN in -> NaN out" also holds with limits / IK
        float nan_probe = ((th[0] + th[1]) + (th[2] + th[3])) + th[4];
        if (do_clamp) {
#pragma unroll
            for (int c = 0; c < 7; ++c) th[c] = fminf(fmaxf(th[c], ap.lower[c]), ap.upper[c]);
        }

        // ---- 6. fused damped-least-squares refinement (never leaves the SM) -------------------
        if (WITH_IK && do_ik) {
            const float4 Tc = quat_conj(zT);
            const float4 Ru = quat_mul_norm_f(Tc, zU);
            const float4 Rf = quat_mul_norm_f(Tc, zL);
            const float4 Rh = quat_mul_norm_f(Tc, zH);
            const vec3 pe_t = add3(p_sh, quat_rotate_f(Ru, make_vec3(ap.seg_elbow[0], ap.seg_elbow[1], ap.seg_elbow[2])));
            const vec3 pw_t = add3(pe_t, quat_rotate_f(Rf, make_vec3(ap.seg_wrist[0], ap.seg_wrist[1], ap.seg_wrist[2])));
            nan_probe += ((pe_t.x + pe_t.y) + (pe_t.z + pw_t.x)) + ((pw_t.y + pw_t.z) + Rh.w);      // NaN targets poison the frame
            const float lam2 = a.damping * a.damping;
            const float wo = a.rot_weight;
            const mat3c H = quat_to_mat3c(Rh);
            for (int it = 0; it < a.ik_iters; ++it) {
                bq_align<BQ_WARPS_PER_CTA, SYSMEM>(warp);
                ik_step_f(th, p_sh, ap.off, ap.lower, ap.upper, pe_t, pw_t, H, lam2, wo, (a.flags & BQ_ACTIVE_SET) != 0);
            }
        }
        if (do_clamp) {
            if (nan_probe != nan_probe) {
#pragma unroll
                for (int c = 0; c < 7; ++c) th[c] = nan_probe;
            }
            // the published local rotations are those of the final hinge angles
            rl[0] = arm_local_quat<0>(th[0]); rl[1] = arm_local_quat<1>(th[1]); rl[2] = arm_local_quat<2>(th[2]);
            rl[3] = arm_local_quat<3>(th[3]); rl[4] = arm_local_quat<4>(th[4]); rl[5] = arm_local_quat<5>(th[5]);
            rl[6] = arm_local_quat<6>(th[6]);
        }

        // ---- 7. outputs: each lane drops its arm into the warp's staged images ---------------------
        if (want_dof && fl < nfr) {
            float* r = dof_t + fl * D + (ap.rob_first - 1);
#pragma unroll
            for (int c = 0; c < 7; ++c) r[c] = th[c];
        }
        if (with_lq) {
            float4* lq_g = reinterpret_cast<float4*>(a.out_local_q + f0 * JR * 4);
            for (int i = lane; i < nfr * JR; i += 32) lq_g[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            __syncwarp();                                  // orders the fill before the patches of other lanes' rows
            if (fl < nfr) {
#pragma unroll
                for (int c = 0; c < 7; ++c) lq_g[fl * JR + ap.rob_first + c] = rl[c];
            }
        }
        if (a.out_link_pos && fl < nfr) {
            float* r = lp_t + fl * W + ap.rob_first * 3;
            float4 G = make_float4(0.f, 0.f, 0.f, 1.f);
            vec3 p = p_sh;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                r[c * 3] = p.x; r[c * 3 + 1] = p.y; r[c * 3 + 2] = p.z;
                G = quat_mul_norm_f(G, rl[c]);
                if (c < 6) p = add3(p, quat_rotate_f(G, make_vec3(ap.off[c + 1][0], ap.off[c + 1][1], ap.off[c + 1][2])));
            }
            // the two gripper links hang off the wrist-yaw link with identity local rotation
#pragma unroll
            for (int c = 7; c < 9; ++c) {
                vec3 pg = add3(p, quat_rotate_f(G, make_vec3(ap.off[c][0], ap.off[c][1], ap.off[c][2])));
                r[c * 3] = pg.x; r[c * 3 + 1] = pg.y; r[c * 3 + 2] = pg.z;
            }
        }
        // ---- 8. the images leave as whole contiguous spans (TMA bulk stores) ----------------------
        if (nfr == BQ_FRAMES_PER_WARP) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (a.out_dof) bulk_store_s2g(a.out_dof + f0 * D, dof_t, (unsigned)(BQ_FRAMES_PER_WARP * D * 4));
                if (a.out_link_pos) bulk_store_s2g(a.out_link_pos + f0 * W, lp_t, (unsigned)(BQ_FRAMES_PER_WARP * W * 4));
                // the same staged image goes to every rank's reassembly buffer over NVLink (TMA bulk stores to peer memory)
                for (int r = 0; r < a.n_peer; ++r)
                    bulk_store_s2g(a.peer_dof[r] + (a.peer_frame0 + f0) * D, dof_t, (unsigned)(BQ_FRAMES_PER_WARP * D * 4));
                bulk_commit();
            }
            if (a.mc_dof) warp_multimem_store_span(a.mc_dof + (a.peer_frame0 + f0) * D, dof_t, BQ_FRAMES_PER_WARP * D, lane);
            pending_store = true;
        } else if (nfr > 0) {
            __syncwarp();
            if (a.out_dof) warp_store_span(a.out_dof + f0 * D, dof_t, nfr * D, lane);
            if (a.out_link_pos) warp_store_span(a.out_link_pos + f0 * W, lp_t, nfr * W, lane);
            for (int r = 0; r < a.n_peer; ++r) warp_store_span(a.peer_dof[r] + (a.peer_frame0 + f0) * D, dof_t, nfr * D, lane);
            if (a.mc_dof) warp_multimem_store_span(a.mc_dof + (a.peer_frame0 + f0) * D, dof_t, nfr * D, lane);
            __syncwarp();
        }
        if (GATHER) {
            // ---- 9. publish the group's packed hinge angles to every rank and count the round ----------------------------
            gather_publish(a.g, gbase, warp, lane, th, nfr, f0);
            // the CTA's unpack warps pace themselves on the rounds this warp has finished
            if (lane == 0) reinterpret_cast<volatile int*>(gbase + bq_gather_progress_word(BQ_WARPS_PER_CTA))[warp] = (int)rnd + 1;
        }
    }
    if (pending_store && lane == 0) bulk_wait_read_all();
    if (GATHER) {
        if (lane == 0) reinterpret_cast<volatile int*>(gbase + bq_gather_progress_word(BQ_WARPS_PER_CTA))[warp] = 0x7fffffff;
        __syncwarp();                                                  // the last bulk stores have read the staging tiles
        // the peer rounds the unpack warps leave to this warp (general path): its publish staging is the slot, its dof
        // tile the image
        const int* st = reinterpret_cast<const int*>(gbase + bq_gather_state_word(BQ_WARPS_PER_CTA)) + warp * BQ_GATHER_STATE_WORDS;
        if (st[2] < st[1] && !(a.g.debug & 16u)) {
            gather_image_clear(dof_t, D, lane);
            __syncwarp();
            for (int j = st[2]; j < st[1]; ++j)
                for (int q = 0; q < a.g.n_rank - 1; ++q)
                    gather_tail_one(a.g, st, gbase + warp * BQ_PK_SLOT, dof_t, BQ_WARPS_PER_CTA, n_ctas, lane, D, ap.rob_first - 1, j, q);
        }
    }
}


template <int BQ_WARPS_PER_CTA, bool WITH_IK = true, bool GATHER = false>
__global__ void __launch_bounds__(BQ_WARPS_PER_CTA * 32, 1)
body_quat_kernel(const __grid_constant__ BodyQuatParams bp, const BodyQuatArgs a) {
    extern __shared__ __align__(16) float smem[];
    bq_setup(bp, smem);
    bq_process<BQ_WARPS_PER_CTA, false, WITH_IK, GATHER>(bp, a, smem, (int)gridDim.x, (int)blockIdx.x);
}

// The reassembly form of the same pipeline: WARPS compute warps + 4 unpack warps, registers re-split
// after the prologue.
template <int WARPS>
__global__ void __launch_bounds__((WARPS + BQ_GATHER_UNPACK_WARPS) * 32, 1)
body_quat_gather_kernel(const __grid_constant__ BodyQuatParams bp, const BodyQuatArgs a) {
    extern __shared__ __align__(16) float smem[];
    bq_setup(bp, smem);
    float* gbase = smem + BQ_CONST_WORDS + WARPS * bq_tile_words(bp.J_src, bp.J_rob, false);
    gather_setup(a.g, gbase, WARPS, (int)gridDim.x, (int)blockIdx.x, bp.J_rob - 1);
    const int warp = threadIdx.x >> 5;
    if (warp >= WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(BQ_GATHER_UNPACK_REGS));
        const int lane = threadIdx.x & 31;
        gather_unpack_warp<WARPS>(a.g, gbase, warp - WARPS, lane, (int)gridDim.x, (int)blockIdx.x, bp.arm[lane & 1].rob_first - 1);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(BQ_GATHER_COMPUTE_REGS));
        bq_process<WARPS, false, true, true>(bp, a, smem, (int)gridDim.x, (int)blockIdx.x);
    }
}

// Resident single-frame server of the quaternion path (same protocol as pos_stream_server_kernel in hrt_pos.cuh:
// ctrl words [0] seq_in, [1] stop (host) / [16] seq_out, [17] exited (device); one warp, one frame per request).
HRT_DEV unsigned bq_ld_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
HRT_DEV void bq_st_sys(unsigned* p, unsigned v) { asm volatile("st.volatile.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(32, 1)
bq_stream_server_kernel(const __grid_constant__ BodyQuatParams bp, const BodyQuatArgs a, unsigned* ctrl, unsigned served,
                        unsigned long long idle_ns) {
    extern __shared__ __align__(16) float smem[];
    bq_setup(bp, smem);
    for (;;) {
        unsigned cmd = 0;
        if (threadIdx.x == 0) {
            const unsigned long long t0 = bq_timer_ns();
            for (;;) {
                const unsigned s = bq_ld_sys(ctrl);
                if (s != served) { cmd = s; break; }
                if (bq_ld_sys(ctrl + 1) != 0u || bq_timer_ns() - t0 > idle_ns) break;
            }
        }
        cmd = __shfl_sync(0xffffffffu, cmd, 0);
        if (cmd == 0u) break;
        bq_process<1, true>(bp, a, smem, 1, 0);
        __threadfence_system();
        __syncwarp();
        if (threadIdx.x == 0) bq_st_sys(ctrl + 16, cmd);
        served = cmd;
    }
    if (threadIdx.x == 0) {
        __threadfence_system();
        bq_st_sys(ctrl + 17, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// Step barrier of the fused reassembly: lane r tells rank r "my stores into your buffer are done" (they were issued by
// the compute kernel that precedes this one on the stream) and waits for rank r's word in this rank's own flag array.
// flags[r] = rank r's array of HRT_MAX_PEERS words (peer-mapped); word [me] of it is written by rank `me` only.
// Epochs only grow; a wait that exceeds timeout_ns traps (a lost peer must not hang the box).
// ---------------------------------------------------------------------------------------------
struct PeerFlags { unsigned* flags[HRT_MAX_PEERS]; };

__global__ void __launch_bounds__(32)
peer_barrier_kernel(const PeerFlags pf, int n_peer, int me, unsigned epoch, unsigned long long timeout_ns) {
    const int r = threadIdx.x;
    if (r >= n_peer) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(pf.flags[r] + me), "r"(epoch) : "memory");
    const unsigned long long t0 = bq_timer_ns();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(pf.flags[me] + r) : "memory");
        if ((int)(v - epoch) >= 0) break;
        if (bq_timer_ns() - t0 > timeout_ns) __trap();
    }
}

// ---------------------------------------------------------------------------------------------
// The same pipeline with the IK refinement on packed fp32x2 registers (FFMA2): the headline path
// (dof + link positions published, flag BQ_IK).  A warp owns 32 consecutive frames as two halves of
// 16; the closed form (exact-order fp32 + fp64 Euler) runs per half with one lane per (frame, arm) as
// above and parks its result in shared memory; the refinement and the final FK then run ONCE with
// lane l holding the arm (frame l/2, side l&1) of BOTH halves in the .x / .y of every register pair.
// The per-arm tables are the same for both (same side) and enter FFMA2 as broadcast scalars.
// ---------------------------------------------------------------------------------------------
#ifndef HRT_BQ2_WARPS
#define HRT_BQ2_WARPS 12
#endif
constexpr int BQ2_WARPS = HRT_BQ2_WARPS;
#ifndef HRT_BQ2_UNROLL_HALVES
#define HRT_BQ2_UNROLL_HALVES 1
#endif
constexpr int kBq2UnrollHalves = HRT_BQ2_UNROLL_HALVES;
constexpr int BQ2_PARK = 17;      // th[7], pe_t[3], pw_t[3], Rh[4] per arm
// the parked values (17 x 64 words) alias the first half's input rows, which are dead by then
HRT_HD inline int bq2_warp_words(int JS, int JR) { return 2 * bq_tile_words(JS, JR, false); }

template <int WARPS>
HRT_DEV void bq_closed_form(const BodyQuatParams& bp, const ArmParams& ap, const float* row, bool pre_transformed,
                            bool do_clamp, int warp, float th[7], float4& zT, float4& zU, float4& zL, float4& zH) {
    zT = *reinterpret_cast<const float4*>(row + ap.src_torso * 4);
    float4 zS = *reinterpret_cast<const float4*>(row + ap.src_shoulder * 4);
    zU = *reinterpret_cast<const float4*>(row + ap.src_upper * 4);
    zL = *reinterpret_cast<const float4*>(row + ap.src_lower * 4);
    zH = *reinterpret_cast<const float4*>(row + ap.src_hand * 4);
    smsp_align<WARPS>(warp);
    if (!pre_transformed) {
        const float4 rot = make_float4(bp.rot_z90[0], bp.rot_z90[1], bp.rot_z90[2], bp.rot_z90[3]);
#define HRT_ZPT(q, n) q = quat_mul_norm_x(quat_mul_norm_x(q, rot), \
        make_float4(-ap.t2z[n][0], -ap.t2z[n][1], -ap.t2z[n][2], ap.t2z[n][3]))
        HRT_ZPT(zT, 0); HRT_ZPT(zS, 1); HRT_ZPT(zU, 2); HRT_ZPT(zL, 3); HRT_ZPT(zH, 4);
#undef HRT_ZPT
    }
    const float4 lU = quat_mul_norm_x(quat_conj(zS), zU);
    const float4 lL = quat_mul_norm_x(quat_conj(zU), zL);
    double sS[3], cS[3], sE[3], cE[3];
    smsp_align<WARPS>(warp);
    euler_intrinsic_half_sincos_f64<1, 0, 2>(lU, sS, cS);
    smsp_align<WARPS>(warp);
    euler_intrinsic_half_sincos_f64<2, 1, 0>(lL, sE, cE);
    smsp_align<WARPS>(warp);
    th[0] = quat_to_dof_x(axis_quat_from_sc(sS[0], cS[0], 1), 1);
    th[1] = quat_to_dof_x(axis_quat_from_sc(sS[1], cS[1], 0), 0);
    th[2] = quat_to_dof_x(quat_mul_norm_x(axis_quat_from_sc(sE[0], cE[0], 2), axis_quat_from_sc(sS[2], cS[2], 2)), 2);
    th[3] = quat_to_dof_x(axis_quat_from_sc(sE[1], cE[1], 1), 1);
    th[4] = quat_to_dof_x(axis_quat_from_sc(sE[2], cE[2], 0), 0);
    th[5] = 0.f;
    th[6] = 0.f;
    if (do_clamp) {
#pragma unroll
        for (int c = 0; c < 7; ++c) th[c] = fminf(fmaxf(th[c], ap.lower[c]), ap.upper[c]);
    }
}

template <int C>
HRT_DEV void ik2_joint(q4p& G, v3p& p, v3p* ax, v3p* pc, const f2 th, const ArmParams& ap) {
    constexpr int K = HRT_ARM_AXIS(C);
    ax[C] = quat_axis_p<K>(G);
    pc[C] = p;
    f2 s, cs;
    sincos_half_p(mul2(dup2(0.5f), th), &s, &cs);
    G = quat_mul_axis_p<K>(G, s, cs);
    if (C < 6) p = quat_rotate_add_p(G, ap.off[C + 1], p);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
body_quat_ik2_kernel(const __grid_constant__ BodyQuatParams bp, const BodyQuatArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fl = lane >> 1;
    const int side = lane & 1;
    {
        const float* src = reinterpret_cast<const float*>(&bp.arm[0]);
        for (int i = threadIdx.x; i < 2 * (int)sizeof(ArmParams) / 4; i += blockDim.x) smem[i] = src[i];
        float* rp = smem + 2 * sizeof(ArmParams) / 4;
        for (int i = threadIdx.x; i < bp.J_rob * 3; i += blockDim.x) rp[i] = bp.rest_pos[i];
    }
    __syncthreads();
    const ArmParams& ap = reinterpret_cast<const ArmParams*>(smem)[side];
    const float* rest_s = smem + 2 * sizeof(ArmParams) / 4;
    const int JS = bp.J_src, JR = bp.J_rob;
    const int D = JR - 1, W = JR * 3;
    const int tile_words = bq_tile_words(JS, JR, false);
    float* wbase = smem + BQ_CONST_WORDS + warp * bq2_warp_words(JS, JR);
    float* park = wbase;                                   // [value][lane] as float2 {half 0, half 1}; aliases half 0's input rows
    bool pending_store = false;
    const long long n_groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    const long long n_pairs = (n_groups + 1) / 2;
    const long long total_warps = (long long)gridDim.x * WARPS;
    const long long rounds = (n_pairs + total_warps - 1) / total_warps;
    const float p_sh[3] = {bp.shoulder_p[side][0], bp.shoulder_p[side][1], bp.shoulder_p[side][2]};
    const f2 lam2 = dup2(a.damping * a.damping);
    const f2 wo = dup2(a.rot_weight), wo2 = dup2(a.rot_weight * a.rot_weight);

    for (long long rnd = 0; rnd < rounds; ++rnd) {
        const long long pair_raw = rnd * total_warps + (long long)blockIdx.x * WARPS + warp;
        const bool pair_live = pair_raw < n_pairs;
        const long long pair = pair_live ? pair_raw : n_pairs - 1;
        if (pending_store) {
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            pending_store = false;
        }
        // ---- inputs of both halves (two contiguous spans; the second may be ragged or absent) ------
        int nfr_h[2];
        long long f0_h[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long grp = min(pair * 2 + h, n_groups - 1);
            f0_h[h] = grp * BQ_FRAMES_PER_WARP;
            const int nld = (int)min((long long)BQ_FRAMES_PER_WARP, a.B - f0_h[h]);
            nfr_h[h] = (pair_live && pair * 2 + h < n_groups) ? nld : 0;
            warp_span_g2s(wbase + h * tile_words, a.src_gq + f0_h[h] * JS * 4, nld * JS * 4, lane);
        }
        cp_async_commit();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float* dof_t = wbase + h * tile_words + bq_io_words(JS, JR);
            if (a.out_dof) for (int i = lane; i < nfr_h[h] * D; i += 32) dof_t[i] = 0.f;
        }
        cp_async_wait<0>();
        __syncwarp();

        // ---- closed form per half, parked as {half 0, half 1} pairs ---------------------------------
#pragma unroll kBq2UnrollHalves
        for (int h = 0; h < 2; ++h) {
            const long long grp = min(pair * 2 + h, n_groups - 1);
            const int nld = (int)min((long long)BQ_FRAMES_PER_WARP, a.B - grp * BQ_FRAMES_PER_WARP);
            const float* row = wbase + h * tile_words + min(fl, nld - 1) * JS * 4;
            float th[7];
            float4 zT, zU, zL, zH;
            bq_closed_form<WARPS>(bp, ap, row, a.pre_transformed != 0, true, warp, th, zT, zU, zL, zH);
            const float4 Tc = quat_conj(zT);
            const float4 Ru = quat_mul_norm_f(Tc, zU);
            const float4 Rf = quat_mul_norm_f(Tc, zL);
            const float4 Rh = quat_mul_norm_f(Tc, zH);
            const vec3 pe_t = add3(make_vec3(p_sh[0], p_sh[1], p_sh[2]),
                                   quat_rotate_f(Ru, make_vec3(ap.seg_elbow[0], ap.seg_elbow[1], ap.seg_elbow[2])));
            const vec3 pw_t = add3(pe_t, quat_rotate_f(Rf, make_vec3(ap.seg_wrist[0], ap.seg_wrist[1], ap.seg_wrist[2])));
            __syncwarp();                                  // every lane has read its input row of this half
            float* pk = park + lane * 2 + h;
#pragma unroll
            for (int c = 0; c < 7; ++c) pk[c * 64] = th[c];
            pk[7 * 64] = pe_t.x; pk[8 * 64] = pe_t.y; pk[9 * 64] = pe_t.z;
            pk[10 * 64] = pw_t.x; pk[11 * 64] = pw_t.y; pk[12 * 64] = pw_t.z;
            pk[13 * 64] = Rh.x; pk[14 * 64] = Rh.y; pk[15 * 64] = Rh.z; pk[16 * 64] = Rh.w;
        }
        __syncwarp();

        // ---- packed refinement: .x = half 0, .y = half 1 ---------------------------------------------
        const f2* pk2 = reinterpret_cast<const f2*>(park) + lane;
        f2 th[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) th[c] = pk2[c * 32];
        const v3p pe_t = make_v3p(pk2[7 * 32], pk2[8 * 32], pk2[9 * 32]);
        const v3p pw_t = make_v3p(pk2[10 * 32], pk2[11 * 32], pk2[12 * 32]);
        q4p Rh;
        Rh.x = pk2[13 * 32]; Rh.y = pk2[14 * 32]; Rh.z = pk2[15 * 32]; Rh.w = pk2[16 * 32];
        const v3p p0 = make_v3p(dup2(p_sh[0]), dup2(p_sh[1]), dup2(p_sh[2]));
        // a NaN anywhere in the warm start or the targets poisons the frame (clamping would hide it)
        f2 nan_probe = add2(add2(add2(th[0], th[1]), add2(th[2], th[3])), add2(add2(th[4], Rh.w), add2(pe_t.x, pw_t.x)));
        nan_probe = add2(nan_probe, add2(add2(pe_t.y, pe_t.z), add2(pw_t.y, pw_t.z)));
        __syncwarp();
        // link-position images start as the robot's rest pose (input rows and parked values are dead now)
        if (a.out_link_pos)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                for (int r = 0; r < nfr_h[h]; ++r)
                    for (int i = lane; i < W; i += 32) wbase[h * tile_words + r * W + i] = rest_s[i];
        __syncwarp();
        for (int it = 0; it < a.ik_iters; ++it) {
            smsp_align<WARPS>(warp);
            v3p ax[7], pc[7];
            q4p G;
            G.x = dup2(0.f); G.y = dup2(0.f); G.z = dup2(0.f); G.w = dup2(1.f);
            v3p p = p0;
            ik2_joint<0>(G, p, ax, pc, th[0], ap); ik2_joint<1>(G, p, ax, pc, th[1], ap); ik2_joint<2>(G, p, ax, pc, th[2], ap);
            ik2_joint<3>(G, p, ax, pc, th[3], ap); ik2_joint<4>(G, p, ax, pc, th[4], ap); ik2_joint<5>(G, p, ax, pc, th[5], ap);
            ik2_joint<6>(G, p, ax, pc, th[6], ap);
            // residual
            const v3p ee = sub3p(pe_t, pc[3]), ew = sub3p(pw_t, pc[6]);
            v3p eo;
            {
                const q4p qe = quat_normalize_p(quat_mul_p(Rh, quat_conj_p(G)));
                const f2 n2 = fma2(qe.z, qe.z, fma2(qe.y, qe.y, mul2(qe.x, qe.x)));
                const f2 n = make_float2(sqrtf(n2.x), sqrtf(n2.y));
                const f2 sc = mul2(wo, make_float2(rotvec_scale_f(n.x, qe.w.x), rotvec_scale_f(n.y, qe.w.y)));
                eo = make_v3p(mul2(qe.x, sc), mul2(qe.y, sc), mul2(qe.z, sc));
            }
            v3p je[3], jw[6];
#pragma unroll
            for (int c = 0; c < 3; ++c) je[c] = cross3p(ax[c], sub3p(pc[3], pc[c]));
#pragma unroll
            for (int c = 0; c < 6; ++c) jw[c] = cross3p(ax[c], sub3p(pc[6], pc[c]));
            f2 A[28], g[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                f2 gi = mul2(wo, dot3p(ax[i], eo));
                if (i < 6) gi = fma2(jw[i].z, ew.z, fma2(jw[i].y, ew.y, fma2(jw[i].x, ew.x, gi)));
                if (i < 3) gi = fma2(je[i].z, ee.z, fma2(je[i].y, ee.y, fma2(je[i].x, ee.x, gi)));
                g[i] = gi;
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    f2 s = (i == j) ? add2(wo2, lam2) : mul2(wo2, dot3p(ax[i], ax[j]));
                    if (i < 6) s = fma2(jw[i].z, jw[j].z, fma2(jw[i].y, jw[j].y, fma2(jw[i].x, jw[j].x, s)));
                    if (i < 3) s = fma2(je[i].z, je[j].z, fma2(je[i].y, je[j].y, fma2(je[i].x, je[j].x, s)));
                    A[i * (i + 1) / 2 + j] = s;
                }
            }
            chol_solve7_p(A, g);
#pragma unroll
            for (int c = 0; c < 7; ++c) th[c] = min2(max2(add2(th[c], g[c]), dup2(ap.lower[c])), dup2(ap.upper[c]));
        }

        // ---- final FK of the refined angles (packed) and the output images ---------------------------
        smsp_align<WARPS>(warp);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (nan_probe.x != nan_probe.x) th[c].x = nan_probe.x;
            if (nan_probe.y != nan_probe.y) th[c].y = nan_probe.y;
        }
        {
            v3p ax[7], pc[9];
            q4p G;
            G.x = dup2(0.f); G.y = dup2(0.f); G.z = dup2(0.f); G.w = dup2(1.f);
            v3p p = p0;
            if (a.out_link_pos) {
                ik2_joint<0>(G, p, ax, pc, th[0], ap); ik2_joint<1>(G, p, ax, pc, th[1], ap); ik2_joint<2>(G, p, ax, pc, th[2], ap);
                ik2_joint<3>(G, p, ax, pc, th[3], ap); ik2_joint<4>(G, p, ax, pc, th[4], ap); ik2_joint<5>(G, p, ax, pc, th[5], ap);
                ik2_joint<6>(G, p, ax, pc, th[6], ap);
                G = quat_normalize_p(G);
                pc[7] = quat_rotate_add_p(G, ap.off[7], pc[6]);      // the two gripper links: identity local rotation
                pc[8] = quat_rotate_add_p(G, ap.off[8], pc[6]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (fl < nfr_h[h]) {
                    float* lp_t = wbase + h * tile_words;
                    float* dof_t = lp_t + bq_io_words(JS, JR);
                    if (a.out_dof) {
                        float* r = dof_t + fl * D + (ap.rob_first - 1);
#pragma unroll
                        for (int c = 0; c < 7; ++c) r[c] = h ? th[c].y : th[c].x;
                    }
                    if (a.out_link_pos) {
                        float* r = lp_t + fl * W + ap.rob_first * 3;
#pragma unroll
                        for (int c = 0; c < 9; ++c) {
                            r[c * 3] = h ? pc[c].x.y : pc[c].x.x;
                            r[c * 3 + 1] = h ? pc[c].y.y : pc[c].y.x;
                            r[c * 3 + 2] = h ? pc[c].z.y : pc[c].z.x;
                        }
                    }
                }
            }
        }
        // ---- the images leave as whole contiguous spans ----------------------------------------------
        fence_proxy_async_smem();
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float* lp_t = wbase + h * tile_words;
            float* dof_t = lp_t + bq_io_words(JS, JR);
            if (nfr_h[h] == BQ_FRAMES_PER_WARP) {
                if (lane == 0) {
                    if (a.out_dof) bulk_store_s2g(a.out_dof + f0_h[h] * D, dof_t, (unsigned)(BQ_FRAMES_PER_WARP * D * 4));
                    if (a.out_link_pos) bulk_store_s2g(a.out_link_pos + f0_h[h] * W, lp_t, (unsigned)(BQ_FRAMES_PER_WARP * W * 4));
                    bulk_commit();
                }
                pending_store = true;
            } else if (nfr_h[h] > 0) {
                if (a.out_dof) warp_store_span(a.out_dof + f0_h[h] * D, dof_t, nfr_h[h] * D, lane);
                if (a.out_link_pos) warp_store_span(a.out_link_pos + f0_h[h] * W, lp_t, nfr_h[h] * W, lane);
            }
        }
        __syncwarp();
    }
    if (pending_store && lane == 0) bulk_wait_read_all();
}

}  // namespace hrt
