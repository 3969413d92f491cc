// Backward of the batched tree FK for sm_100a: the vector-Jacobian product that makes
//   HuForwardModel.forward_kinematics   robot_kinematics_model/hu_forward_model.py:17-33
// differentiable, as the reference's straight-through clamp `(clamp(x) - x).detach() + x` (:27-33) intends
// (invariant 4).  The (B, K, 6, D) geometric Jacobian is never materialised: with
//   a_i = R_parent(i) e_i   world axis of hinge i          p_i   its pivot
//   f_k = dL/dp_k                                          tau_k = 1/2 (w g_u + u x g_u - g_w u),  q_k = (u, w), dL/dq_k = (g_u, g_w)
// the gradient of hinge i is   a_i . sum_{k in subtree(i)} [ tau_k + (p_k - p_i) x f_k ],
// i.e. one reverse sweep over the tree accumulating a wrench (S, F) = (sum tau_k + p_k x f_k, sum f_k) per subtree.
// The spec, term by term (root quaternion enters un-normalised: direct, tangential and radial parts), is
// `fk_vjp_analytic` of the test-side CPU restatement, which the CPU test suite checks against torch.autograd.
//
// One thread per configuration.  Forward walk in registers: the parent of joint j is joint j-1 or one of <= 4 parked
// branch points (TreeParams slots, allocated by liveness on the host).  Per joint the walk leaves pivot, axis and the
// joint's own wrench in thread-local memory (9 words; interleaved across the warp, so every access is one coalesced
// line); the reverse sweep runs the same slot scheme backwards: a chain passes its running wrench down in registers, a
// branch deposits it in its parent's slot accumulator.  Gradient rows are staged per warp in shared memory and leave
// as contiguous spans.
#pragma once
#include "hrt_fk_limb.cuh"
#include "hrt_math.cuh"
#include "hrt_params.h"
#include "hrt_retarget.cuh"

namespace hrt {

struct FkVjpArgs {
    long long B;
    const float* __restrict__ angles;    // (B, D)
    const float* __restrict__ root_t;    // (B, 3) or nullptr (= 0)
    const float* __restrict__ root_q;    // (B, 4) or nullptr (= identity)
    int clip;
    const float* __restrict__ g_gq;      // (B, J, 4) upstream gradient of the global quaternions, or nullptr (= 0)
    const float* __restrict__ g_gt;      // (B, J, 3) upstream gradient of the link positions, or nullptr (= 0)
    float* __restrict__ g_angles;        // (B, D)
    float* __restrict__ g_root_t;        // (B, 3) or nullptr
    float* __restrict__ g_root_q;        // (B, 4) or nullptr
};

constexpr int VJP_WARPS = 4;
HRT_HD inline int vjp_row_words(int D) { return D | 1; }                       // odd stride: conflict-free row writes
HRT_HD inline size_t vjp_smem_bytes(int D) { return (size_t)VJP_WARPS * 32 * vjp_row_words(D) * sizeof(float); }

struct Wrench { vec3 S, F; };
HRT_DEV Wrench wrench_zero() { Wrench w; w.S = make_vec3(0.f, 0.f, 0.f); w.F = make_vec3(0.f, 0.f, 0.f); return w; }
HRT_DEV void wrench_add(Wrench& a, const Wrench& b) { a.S = add3(a.S, b.S); a.F = add3(a.F, b.F); }

__global__ void __launch_bounds__(VJP_WARPS * 32)
fk_vjp_kernel(const __grid_constant__ TreeParams tp, const FkVjpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int J = tp.J, D = J - 1, RW = vjp_row_words(D);
    float* rows = smem + warp * 32 * RW;
    const long long n_groups = (a.B + 31) / 32;
    for (long long grp = (long long)blockIdx.x * VJP_WARPS + warp; grp < n_groups; grp += (long long)gridDim.x * VJP_WARPS) {
        const long long f0 = grp * 32;
        const int rows_live = (int)min(32LL, a.B - f0);
        const long long b = f0 + min(lane, rows_live - 1);               // tail lanes shadow the last configuration
        // ---- root -------------------------------------------------------------------------------------------
        const float4 q0 = a.root_q ? __ldg(reinterpret_cast<const float4*>(a.root_q) + b) : make_float4(0.f, 0.f, 0.f, 1.f);
        vec3 t0 = make_vec3(0.f, 0.f, 0.f);
        if (a.root_t) t0 = make_vec3(__ldg(a.root_t + b * 3), __ldg(a.root_t + b * 3 + 1), __ldg(a.root_t + b * 3 + 2));
        const float s2 = q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w;
        const float inv_s = rsqrtf(fmaxf(s2, 1e-30f));
        const float4 n0 = make_float4(q0.x * inv_s, q0.y * inv_s, q0.z * inv_s, q0.w * inv_s);
        // ---- forward walk ----------------------------------------------------------------------------------------
        vec3 P[HRT_MAX_JOINTS], A[HRT_MAX_JOINTS], S[HRT_MAX_JOINTS];     // thread-local memory
        float4 slot_q[HRT_MAX_SLOTS];
        vec3 slot_p[HRT_MAX_SLOTS];
        float4 cq = n0;
        vec3 cp = t0;
        {
            const int sv = jr_save(tp.jr[0].meta);
#pragma unroll
            for (int s = 0; s < HRT_MAX_SLOTS; ++s) if (sv == s) { slot_q[s] = cq; slot_p[s] = cp; }
        }
        for (int j = 1; j < J; ++j) {
            const float4 rec = *reinterpret_cast<const float4*>(&tp.jr[j]);
            const uint32_t meta = __float_as_uint(rec.w);
            const int src = jr_src(meta), sv = jr_save(meta), kx = jr_axis(meta);
            float4 pq = cq;
            vec3 pp = cp;
#pragma unroll
            for (int s = 0; s < HRT_MAX_SLOTS; ++s) if (src == s) { pq = slot_q[s]; pp = slot_p[s]; }
            const vec3 e = make_vec3(kx == 0 ? 1.f : 0.f, kx == 1 ? 1.f : 0.f, kx == 2 ? 1.f : 0.f);
            const vec3 ax = quat_rotate_f(pq, e);
            vec3 step = quat_rotate_f(pq, make_vec3(rec.x, rec.y, rec.z));
            if (tp.parent[j] == 0) step = make_vec3(step.x * s2, step.y * s2, step.z * s2);    // quat_rotate by the RAW root quat
            const vec3 pj = add3(pp, step);
            float ang = __ldg(a.angles + b * D + (j - 1));
            if (a.clip) {
                const float cl = fminf(fmaxf(ang, tp.lim[j][0]), tp.lim[j][1]);
                ang = add_rn(sub_rn(cl, ang), ang);                      // forward value of the straight-through clamp
            }
            float sn, cs;
            sincos_half_f(0.5f * ang, &sn, &cs);
            const float4 qj = quat_normalize_f(quat_mul_axis_f(pq, kx, sn, cs));    // w >= 0: the published representative
            // this joint's own wrench: tau_j + p_j x f_j
            vec3 own = make_vec3(0.f, 0.f, 0.f);
            if (a.g_gq) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(a.g_gq) + b * J + j);
                const vec3 u = make_vec3(qj.x, qj.y, qj.z), gu = make_vec3(g.x, g.y, g.z);
                const vec3 c = cross3_f(u, gu);
                own = make_vec3(0.5f * (qj.w * gu.x + c.x - g.w * u.x), 0.5f * (qj.w * gu.y + c.y - g.w * u.y),
                                0.5f * (qj.w * gu.z + c.z - g.w * u.z));
            }
            if (a.g_gt) {
                const float* gp = a.g_gt + (b * J + j) * 3;
                own = add3(own, cross3_f(pj, make_vec3(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2))));
            }
            P[j] = pj; A[j] = ax; S[j] = own;
            cq = qj; cp = pj;
#pragma unroll
            for (int s = 0; s < HRT_MAX_SLOTS; ++s) if (sv == s) { slot_q[s] = cq; slot_p[s] = cp; }
        }
        // ---- reverse sweep -----------------------------------------------------------------------------------------
        Wrench run = wrench_zero(), acc[HRT_MAX_SLOTS];
#pragma unroll
        for (int s = 0; s < HRT_MAX_SLOTS; ++s) acc[s] = wrench_zero();
        float g_s = 0.f;
        float* myrow = rows + lane * RW;
        for (int j = J - 1; j >= 1; --j) {
            const uint32_t meta = tp.jr[j].meta;
            const int src = jr_src(meta), sv = jr_save(meta);
            Wrench tot;
            tot.S = S[j];
            tot.F = make_vec3(0.f, 0.f, 0.f);
            if (a.g_gt) {
                const float* gp = a.g_gt + (b * J + j) * 3;
                tot.F = make_vec3(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2));
            }
            if (j + 1 < J && tp.parent[j + 1] == j) wrench_add(tot, run);     // the chain below hands its wrench up
#pragma unroll
            for (int s = 0; s < HRT_MAX_SLOTS; ++s)
                if (sv == s) { wrench_add(tot, acc[s]); acc[s] = wrench_zero(); }   // branches parked under this joint
            const vec3 pj = P[j];
            myrow[j - 1] = dot3_f(A[j], sub3(tot.S, cross3_f(pj, tot.F)));
            if (tp.parent[j] == 0) g_s += 2.f * dot3_f(sub3(pj, t0), tot.F);
            if (src < 0) {
                run = tot;
            } else {
#pragma unroll
                for (int s = 0; s < HRT_MAX_SLOTS; ++s) if (src == s) wrench_add(acc[s], tot);
            }
        }
        // ---- root gradients ----------------------------------------------------------------------------------------
        if (a.g_root_t || a.g_root_q) {
            Wrench C = wrench_zero();                                   // everything below the root
            if (J > 1) C = run;                                         // joint 1 always hangs off joint 0
            const int sv0 = jr_save(tp.jr[0].meta);
#pragma unroll
            for (int s = 0; s < HRT_MAX_SLOTS; ++s) if (sv0 == s) wrench_add(C, acc[s]);
            vec3 f0v = make_vec3(0.f, 0.f, 0.f);
            if (a.g_gt) f0v = make_vec3(__ldg(a.g_gt + b * J * 3), __ldg(a.g_gt + b * J * 3 + 1), __ldg(a.g_gt + b * J * 3 + 2));
            if (a.g_root_t && lane < rows_live) {
                a.g_root_t[b * 3] = C.F.x + f0v.x; a.g_root_t[b * 3 + 1] = C.F.y + f0v.y; a.g_root_t[b * 3 + 2] = C.F.z + f0v.z;
            }
            if (a.g_root_q && lane < rows_live) {
                const vec3 gphi = sub3(C.S, cross3_f(t0, C.F));
                const vec3 u = make_vec3(q0.x, q0.y, q0.z);
                const vec3 c = cross3_f(gphi, u);
                const float k2 = 2.f / s2, kr = g_s / s2;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.g_gq) g = __ldg(reinterpret_cast<const float4*>(a.g_gq) + b * J);      // link 0's rotation IS root_q
                g.x += k2 * (q0.w * gphi.x + c.x) + kr * q0.x;
                g.y += k2 * (q0.w * gphi.y + c.y) + kr * q0.y;
                g.z += k2 * (q0.w * gphi.z + c.z) + kr * q0.z;
                g.w += -k2 * dot3_f(gphi, u) + kr * q0.w;
                reinterpret_cast<float4*>(a.g_root_q)[b] = g;
            }
        }
        // ---- the warp's gradient rows leave as one contiguous span ----------------------------------------------------
        __syncwarp();
        float* dst = a.g_angles + f0 * D;
        for (int i = lane; i < rows_live * D; i += 32) dst[i] = rows[(i / D) * RW + (i % D)];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone damped-least-squares refinement of both arms (the stage that is fused into body_quat_kernel /
// pos_retarget_kernel, callable on its own: SURVEY.md section 8(b) `hrt_ik_refine`; spec in DESIGN.md section 5,
// `ik_refine_arm` of the test-side CPU restatement).  One thread per (frame, arm).
// ---------------------------------------------------------------------------------------------
struct IkRefineArgs {
    long long B;
    const float* __restrict__ theta0;    // (B, 2, 7) warm start (clamped to the limits on entry)
    const float* __restrict__ pe_t;      // (B, 2, 3) elbow-pitch link target position, robot root frame
    const float* __restrict__ pw_t;      // (B, 2, 3) wrist-yaw link target position
    const float* __restrict__ qw_t;      // (B, 2, 4) wrist-yaw link target orientation
    float* __restrict__ theta;           // (B, 2, 7)
    float* __restrict__ residual;        // (B, 2, iters + 1) ||e|| before every step and after the last, or nullptr
    int iters;
    float damping, rot_weight;
    int active_set;
};

struct IkArmTables {                     // both arms of the configured robot (filled from BodyQuatParams)
    float off[2][9][3];
    float lower[2][7], upper[2][7];
    float p_sh[2][3];
};

HRT_DEV float ik_residual_norm_f(const float th[7], const vec3 p_sh, const float (*off)[3], const vec3 pe_t, const vec3 pw_t,
                                 const float4 Rh, const float wo) {
    vec3 ax[7], pc[7];
    float4 G;
    arm_chain_f(th, p_sh, off, ax, pc, G);
    const vec3 ee = sub3(pe_t, pc[3]), ew = sub3(pw_t, pc[6]);
    const float4 qe = quat_normalize_f(quat_mul_f(Rh, quat_conj(G)));
    const float n = sqrtf(qe.x * qe.x + qe.y * qe.y + qe.z * qe.z);
    const float ang = wo * rotvec_scale_f(n, qe.w) * n;
    return sqrtf(dot3_f(ee, ee) + dot3_f(ew, ew) + ang * ang);
}

__global__ void __launch_bounds__(128)
ik_refine_kernel(const __grid_constant__ IkArmTables tb, const IkRefineArgs a) {
    const long long n = a.B * 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int side = (int)(i & 1);
        float th[7];
        float nan_probe = 0.f;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const float v = __ldg(a.theta0 + i * 7 + c);
            nan_probe += v;
            th[c] = fminf(fmaxf(v, tb.lower[side][c]), tb.upper[side][c]);
        }
        const vec3 p_sh = make_vec3(tb.p_sh[side][0], tb.p_sh[side][1], tb.p_sh[side][2]);
        const vec3 pe_t = make_vec3(__ldg(a.pe_t + i * 3), __ldg(a.pe_t + i * 3 + 1), __ldg(a.pe_t + i * 3 + 2));
        const vec3 pw_t = make_vec3(__ldg(a.pw_t + i * 3), __ldg(a.pw_t + i * 3 + 1), __ldg(a.pw_t + i * 3 + 2));
        const float4 Rh = __ldg(reinterpret_cast<const float4*>(a.qw_t) + i);
        nan_probe += ((pe_t.x + pe_t.y) + (pe_t.z + pw_t.x)) + ((pw_t.y + pw_t.z) + ((Rh.x + Rh.y) + (Rh.z + Rh.w)));
        const float lam2 = a.damping * a.damping;
        const mat3c H = quat_to_mat3c(Rh);
        for (int it = 0; it < a.iters; ++it) {
            if (a.residual) a.residual[i * (a.iters + 1) + it] = ik_residual_norm_f(th, p_sh, tb.off[side], pe_t, pw_t, Rh, a.rot_weight);
            ik_step_f(th, p_sh, tb.off[side], tb.lower[side], tb.upper[side], pe_t, pw_t, H, lam2, a.rot_weight, a.active_set != 0);
        }
        if (a.residual) a.residual[i * (a.iters + 1) + a.iters] = ik_residual_norm_f(th, p_sh, tb.off[side], pe_t, pw_t, Rh, a.rot_weight);
#pragma unroll
        for (int c = 0; c < 7; ++c) a.theta[i * 7 + c] = (nan_probe != nan_probe) ? nan_probe : th[c];     // NaN in -> NaN out
    }
}

}  // namespace hrt
