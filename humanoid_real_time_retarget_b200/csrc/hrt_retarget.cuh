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
#include "hrt_params.h"

namespace hrt {

// ---------------------------------------------------------------------------------------------
// Element-wise kernels: one thread per (frame, joint), 16-byte coalesced loads and stores.
// ---------------------------------------------------------------------------------------------
// q' = norm(norm(q * rot) * inv(T2Z[j]))            parse_mocap.py:106-114
__global__ void __launch_bounds__(256)
zero_pose_transform_kernel(const float4* __restrict__ gq, const float4* __restrict__ t2z, float4 rot,
                           int J, long long n_items, float4* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % J);
        float4 q = __ldcs(gq + i);
        q = quat_mul_norm_x(q, rot);
        q = quat_mul_norm_x(q, quat_conj(__ldg(t2z + j)));
        __stcs(out + i, q);
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

constexpr int BQ_FRAMES_PER_WARP = 16;
// One CTA per SM.  16 warps when only dof / link positions are published (the headline path),
// 8 when the 496-byte local-rotation rows are staged as well (shared-memory budget).
constexpr int BQ_WARPS_WIDE = 16;
constexpr int BQ_WARPS_NARROW = 8;
// warp-private staging, in words: [input rows, later the link-position image] [dof image]
// [local-rotation image, only when that output is requested]
HRT_HD inline int bq_io_words(int JS, int JR) { return BQ_FRAMES_PER_WARP * (JS * 4 > JR * 3 ? JS * 4 : JR * 3); }
HRT_HD inline int bq_dof_words(int JR) { return (BQ_FRAMES_PER_WARP * (JR - 1) + 3) / 4 * 4; }
HRT_HD inline int bq_tile_words(int JS, int JR, bool with_lq) {
    return bq_io_words(JS, JR) + bq_dof_words(JR) + (with_lq ? BQ_FRAMES_PER_WARP * JR * 4 : 0);
}
// CTA-shared constants: both ArmParams + the robot's rest positions
constexpr int BQ_CONST_WORDS = (2 * (int)sizeof(ArmParams) / 4 + HRT_MAX_JOINTS * 3 + 3) / 4 * 4;

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
};

// the arm's hinge axes (Hu_DOF_AXIS[11..17] == Hu_DOF_AXIS[20..26]); checked on the host
#define HRT_ARM_AXIS(c) ((c) == 0 ? 1 : (c) == 1 ? 0 : (c) == 2 ? 2 : (c) == 3 ? 1 : (c) == 4 ? 0 : (c) == 5 ? 1 : 2)

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
    sincos_half_nf(0.5f * th, &s, &c);
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
            if (i == j) inv[i] = rsqrtf(s);
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

// Warps that share a scheduler (warp % 4) re-align at the top of every IK iteration with a named
// barrier: the unrolled iteration body is ~16 KB of SASS, far more than an SMSP's L0 instruction
// cache, and ncu showed "no instruction" as the top stall (2.4 warps per issue) when the warps
// of a scheduler drifted apart and each streamed the body on its own (profiles/r01_body_quat.md).
#ifndef HRT_BQ_ALIGN
#define HRT_BQ_ALIGN 1
#endif
template <int WARPS>
HRT_DEV void smsp_align(int warp) {
#if HRT_BQ_ALIGN
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + (warp & 3)), "n"(WARPS * 8) : "memory");
#endif
}
template <int BQ_WARPS_PER_CTA>
__global__ void __launch_bounds__(BQ_WARPS_PER_CTA * 32, 1)
body_quat_kernel(const __grid_constant__ BodyQuatParams bp, const BodyQuatArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fl = lane >> 1;
    const int side = lane & 1;
    // per-arm tables and rest positions: constant bank -> shared memory once per CTA (the two lanes of
    // a frame index different arms, which would serialise every constant-bank read)
    {
        const float* src = reinterpret_cast<const float*>(&bp.arm[0]);
        for (int i = threadIdx.x; i < 2 * (int)sizeof(ArmParams) / 4; i += blockDim.x) smem[i] = src[i];
        float* rp = smem + 2 * sizeof(ArmParams) / 4;
        for (int i = threadIdx.x; i < bp.J_rob * 3; i += blockDim.x) rp[i] = bp.rest_pos[i];
    }
    __syncthreads();
    const ArmParams& ap = reinterpret_cast<const ArmParams*>(smem)[side];
    const float* rest_s = smem + 2 * sizeof(ArmParams) / 4;
    const bool with_lq = a.out_local_q != nullptr;
    float* tile = smem + BQ_CONST_WORDS + warp * bq_tile_words(bp.J_src, bp.J_rob, with_lq);
    float* lp_t = tile;                                   // input rows, later the link-position image
    float* dof_t = tile + bq_io_words(bp.J_src, bp.J_rob);
    float* lq_t = dof_t + bq_dof_words(bp.J_rob);
    bool pending_store = false;
    const long long n_groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    const int JS = bp.J_src, JR = bp.J_rob;
    const int D = JR - 1, W = JR * 3;
    const bool do_clamp = (a.flags & (BQ_CLAMP | BQ_IK)) != 0;
    const bool do_ik = (a.flags & BQ_IK) != 0;
    const vec3 p_sh = make_vec3(bp.shoulder_p[side][0], bp.shoulder_p[side][1], bp.shoulder_p[side][2]);

    // every warp of the CTA runs the same number of rounds (the named barriers below need that);
    // a warp without a group in the last round shadows the last group and publishes nothing
    const long long total_warps = (long long)gridDim.x * BQ_WARPS_PER_CTA;
    const long long rounds = (n_groups + total_warps - 1) / total_warps;
    for (long long rnd = 0; rnd < rounds; ++rnd) {
        const long long grp_raw = rnd * total_warps + (long long)blockIdx.x * BQ_WARPS_PER_CTA + warp;
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
        warp_span_g2s(tile, a.src_gq + f0 * JS * 4, nld * JS * 4, lane);
        cp_async_commit();
        // while the copy is in flight: pre-fill the output images with their constant parts
        if (a.out_dof) for (int i = lane; i < nfr * D; i += 32) dof_t[i] = 0.f;
        if (with_lq)
            for (int i = lane; i < nfr * JR; i += 32) *reinterpret_cast<float4*>(lq_t + i * 4) = make_float4(0.f, 0.f, 0.f, 1.f);
        cp_async_wait<0>();
        __syncwarp();
        const float* row = tile + fr * JS * 4;
        float4 zT = *reinterpret_cast<const float4*>(row + ap.src_torso * 4);
        float4 zS = *reinterpret_cast<const float4*>(row + ap.src_shoulder * 4);
        float4 zU = *reinterpret_cast<const float4*>(row + ap.src_upper * 4);
        float4 zL = *reinterpret_cast<const float4*>(row + ap.src_lower * 4);
        float4 zH = *reinterpret_cast<const float4*>(row + ap.src_hand * 4);
        __syncwarp();                              // the input rows may now be overwritten:
        if (a.out_link_pos)                        // link-position image starts as the robot's rest pose
            for (int r = 0; r < nfr; ++r)
                for (int i = lane; i < W; i += 32) lp_t[r * W + i] = rest_s[i];

        // ---- 2. zero-pose re-referencing (a24), exact rounding order -------------------------
        smsp_align<BQ_WARPS_PER_CTA>(warp);
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
            double eS[3], eE[3];
            smsp_align<BQ_WARPS_PER_CTA>(warp);
            euler_intrinsic_f64<1, 0, 2>(lU, eS);      // 'YXZ': pitch, roll, yaw
            smsp_align<BQ_WARPS_PER_CTA>(warp);
            euler_intrinsic_f64<2, 1, 0>(lL, eE);      // 'ZYX': yaw, pitch, roll
            smsp_align<BQ_WARPS_PER_CTA>(warp);
            rl[0] = axis_quat_from_f64(eS[0], 1);
            rl[1] = axis_quat_from_f64(eS[1], 0);
            rl[2] = quat_mul_norm_x(axis_quat_from_f64(eE[0], 2), axis_quat_from_f64(eS[2], 2));
            rl[3] = axis_quat_from_f64(eE[1], 1);
            rl[4] = axis_quat_from_f64(eE[2], 0);
            rl[5] = make_float4(0.f, 0.f, 0.f, 1.f);
            rl[6] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        // ---- 5. hinge angles (a17) ------------------------------------------------------------
        smsp_align<BQ_WARPS_PER_CTA>(warp);
        float th[7];
        th[0] = quat_to_dof_x(rl[0], 1);
        th[1] = quat_to_dof_x(rl[1], 0);
        th[2] = quat_to_dof_x(rl[2], 2);
        th[3] = quat_to_dof_x(rl[3], 1);
        th[4] = quat_to_dof_x(rl[4], 0);
        th[5] = 0.f;
        th[6] = 0.f;

        if (do_clamp) {
#pragma unroll
            for (int c = 0; c < 7; ++c) th[c] = fminf(fmaxf(th[c], ap.lower[c]), ap.upper[c]);
        }

        // ---- 6. fused damped-least-squares refinement (never leaves the SM) -------------------
        if (do_ik) {
            const float4 Tc = quat_conj(zT);
            const float4 Ru = quat_mul_norm_f(Tc, zU);
            const float4 Rf = quat_mul_norm_f(Tc, zL);
            const float4 Rh = quat_mul_norm_f(Tc, zH);
            const vec3 pe_t = add3(p_sh, quat_rotate_f(Ru, make_vec3(ap.seg_elbow[0], ap.seg_elbow[1], ap.seg_elbow[2])));
            const vec3 pw_t = add3(pe_t, quat_rotate_f(Rf, make_vec3(ap.seg_wrist[0], ap.seg_wrist[1], ap.seg_wrist[2])));
            const float lam2 = a.damping * a.damping;
            const float wo = a.rot_weight;
            for (int it = 0; it < a.ik_iters; ++it) {
                smsp_align<BQ_WARPS_PER_CTA>(warp);
                vec3 ax[7], pc[7];
                float4 G = make_float4(0.f, 0.f, 0.f, 1.f);
                vec3 p = p_sh;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    constexpr int dummy = 0; (void)dummy;
                    ax[c] = (HRT_ARM_AXIS(c) == 0) ? quat_axis_f<0>(G) : (HRT_ARM_AXIS(c) == 1) ? quat_axis_f<1>(G) : quat_axis_f<2>(G);
                    pc[c] = p;
                    float s, cs;
                    sincos_half_nf(0.5f * th[c], &s, &cs);
                    // products of unit quaternions: renormalising once, below, is enough inside the loop
                    G = quat_mul_axis_f(G, HRT_ARM_AXIS(c), s, cs);
                    if (c < 6) p = add3(p, quat_rotate_f(G, make_vec3(ap.off[c + 1][0], ap.off[c + 1][1], ap.off[c + 1][2])));
                }
                // residual
                float e[9];
                e[0] = pe_t.x - pc[3].x; e[1] = pe_t.y - pc[3].y; e[2] = pe_t.z - pc[3].z;
                e[3] = pw_t.x - pc[6].x; e[4] = pw_t.y - pc[6].y; e[5] = pw_t.z - pc[6].z;
                {
                    const float4 qe = quat_normalize_f(quat_mul_f(Rh, quat_conj(G)));
                    const float n = sqrtf(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
                    const float sc = wo * rotvec_scale_f(n, qe.w);
                    e[6] = qe.x * sc; e[7] = qe.y * sc; e[8] = qe.z * sc;
                }
                // Jacobian blocks
                vec3 je[3], jw[6];
#pragma unroll
                for (int c = 0; c < 3; ++c) je[c] = cross3_f(ax[c], sub3(pc[3], pc[c]));
#pragma unroll
                for (int c = 0; c < 6; ++c) jw[c] = cross3_f(ax[c], sub3(pc[6], pc[c]));
                float A[28], g[7];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    float gi = wo * (ax[i].x * e[6] + ax[i].y * e[7] + ax[i].z * e[8]);
                    if (i < 6) gi += jw[i].x * e[3] + jw[i].y * e[4] + jw[i].z * e[5];
                    if (i < 3) gi += je[i].x * e[0] + je[i].y * e[1] + je[i].z * e[2];
                    g[i] = gi;
#pragma unroll
                    for (int j = 0; j <= i; ++j) {
                        float s = wo * wo * dot3_f(ax[i], ax[j]);
                        if (i < 6) s += dot3_f(jw[i], jw[j]);
                        if (i < 3) s += dot3_f(je[i], je[j]);
                        if (i == j) s += lam2;
                        A[i * (i + 1) / 2 + j] = s;
                    }
                }
                chol_solve7(A, g);
#pragma unroll
                for (int c = 0; c < 7; ++c) th[c] = fminf(fmaxf(th[c] + g[c], ap.lower[c]), ap.upper[c]);
            }
        }
        if (do_clamp) {
            // the published local rotations are those of the final hinge angles
            rl[0] = arm_local_quat<0>(th[0]); rl[1] = arm_local_quat<1>(th[1]); rl[2] = arm_local_quat<2>(th[2]);
            rl[3] = arm_local_quat<3>(th[3]); rl[4] = arm_local_quat<4>(th[4]); rl[5] = arm_local_quat<5>(th[5]);
            rl[6] = arm_local_quat<6>(th[6]);
        }

        // ---- 7. outputs: each lane drops its arm into the warp's staged images ---------------------
        if (a.out_dof && fl < nfr) {
            float* r = dof_t + fl * D + (ap.rob_first - 1);
#pragma unroll
            for (int c = 0; c < 7; ++c) r[c] = th[c];
        }
        if (with_lq && fl < nfr) {
            float* r = lq_t + (fl * JR + ap.rob_first) * 4;
#pragma unroll
            for (int c = 0; c < 7; ++c) *reinterpret_cast<float4*>(r + c * 4) = rl[c];
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
                if (with_lq) bulk_store_s2g(a.out_local_q + f0 * JR * 4, lq_t, (unsigned)(BQ_FRAMES_PER_WARP * JR * 16));
                if (a.out_link_pos) bulk_store_s2g(a.out_link_pos + f0 * W, lp_t, (unsigned)(BQ_FRAMES_PER_WARP * W * 4));
                bulk_commit();
            }
            pending_store = true;
        } else if (nfr > 0) {
            __syncwarp();
            if (a.out_dof) warp_store_span(a.out_dof + f0 * D, dof_t, nfr * D, lane);
            if (with_lq) warp_store_span(a.out_local_q + f0 * JR * 4, lq_t, nfr * JR * 4, lane);
            if (a.out_link_pos) warp_store_span(a.out_link_pos + f0 * W, lp_t, nfr * W, lane);
            __syncwarp();
        }
    }
    if (pending_store && lane == 0) bulk_wait_read_all();
}

}  // namespace hrt
