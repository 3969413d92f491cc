// Position-input retarget paths for sm_100a (what the live teleop loop runs).
//
// Replaces the numeric bodies of
//   cal_joint_quat (Kabsch)                      retarget/spatial_transform/transform3d.py:32-50
//   proj_in_plane / radians_between_vecs         transform3d.py:62-75, 78-100
//   cal_shoulderPR / cal_elbowP_and_shoulderY    retarget/retarget_solver/full_body_pos_retargeter.py:221-278
//                                                (= retarget_solver.py:103-158 = full_body_retargeter.py:184-241)
//   VtrdynFullBodyPosRetargeter.retarget         full_body_pos_retargeter.py:25-217            (mode POS)
//   HuUpperBodyFromMocapRetarget.retarget_from_global_translation   retarget_solver.py:40-99   (mode UPPER)
//   VtrdynFullBodyRetargeter.retarget            full_body_retargeter.py:19-177                (mode FULL)
//   quat_in_xyz_axis 'XYZ' (SciPy, fp64), quat_to_dof_pos
//
// One thread per (frame, arm), as in the quaternion path.  The fp32 steps restate the reference's
// formulas in its own evaluation order (acos(clamp(dot)) * sign(..), not atan2), because the
// reference is ill-conditioned there and only the same rounding sequence tracks it.
// The 3x3 SVD of the Kabsch step runs in fp64 as a Jacobi eigen-decomposition of A^T A: only
// the two leading singular pairs are used, the third axis is their cross product, which is exactly
// what the reference's det < 0 fix-up selects and stays well defined for the rank-2 torso frame.
#pragma once
#include "hrt_fk_limb.cuh"
#include "hrt_math.cuh"
#include "hrt_params.h"
#include "hrt_retarget.cuh"

namespace hrt {

// POS_MAIN: RetargetHuV5fromMocap (retarget/main.py:201-240): positions + ONE measured parent quat for both arms
enum PosMode { POS_FULL_BODY_POS = 0, POS_UPPER_BODY = 1, POS_FULL_BODY = 2, POS_MAIN = 3 };

struct PosArm {
    int b_sh, b_el, b_wr;          // body joints: shoulder-end, elbow, wrist (vtrdyn: 18,19,20 / 14,15,16)
    int q_parent, q_wrist;         // mode FULL: measured quats used as arm parent / wrist (17,20 / 13,16)
    int rob_first;                 // robot shoulder-pitch joint (12 / 21)
    int bq_wrist;                  // mode POS: slot of the wrist quat in body_global_rotation (14 / 39)
    float v0_upper[3], v0_lower[3];
    float zwrist[5][3];            // zero-pose offsets paired with the 5 finger points
};

// robot-side tables of one arm for the limit-aware refinement (flags POS_CLAMP / POS_IK)
struct ArmIkParams {
    float off[9][3];               // robot offsets of the 7 arm hinges + 2 gripper links
    float lower[7], upper[7];
    float p_sh[3];                 // shoulder-pitch link position at the zero pose (root frame)
};

constexpr unsigned POS_CLAMP = 1u; // clamp the 7 arm hinge angles of each arm to the robot limits
constexpr unsigned POS_IK = 2u;    // + damped-least-squares steps towards the pose of the UNCLAMPED closed form

struct PosParams {
    int mode;
    int J_rob;                     // 31
    int J_bq;                      // 59 (mode POS output), else 0
    int n_body, n_hand;            // rows per frame of body_t / hand_t: 21 (23 in the wire layout), 20
    int n_bodyq;                   // rows per frame of body_q (21)
    int torso_pts[3], torso_org;   // 17,13,11 ; 10
    int bq_torso;                  // 10
    int hand_kabsch[5], hand_org;  // 2,6,10,14,17 ; 0
    int hand_tips[5];              // POS: 4,8,12,16,19   FULL: 3,7,11,15,19
    float ztorso[3][3];            // zero-pose offsets paired with torso_pts
    float flip[3];                 // coord_transform dir (UPPER: -1,-1,1), else 1,1,1
    float orig_x;                  // mean zero-pose finger-tip x extent
    int precise_gripper;
    PosArm arm[2];
    ArmIkParams ik[2];
    // bone angles of the zero pose, [side][theta0_sh, phi0_sh, theta0_el, phi0_el]: evaluated ONCE per configuration by
    // pos_zero_angles_kernel with the very device code the frames run through (same libm, same rounding order)
    float zero_ang[2][4];
};
// CTA-shared constants in shared memory: [2 x PosArm][8 zero-pose angles][2 x ArmIkParams]
HRT_HD inline int pos_zero_ang_word() { return 2 * (int)sizeof(PosArm) / 4; }
HRT_HD inline int pos_ik_word() { return (pos_zero_ang_word() + 8 + 3) / 4 * 4; }
// zero-pose offsets of the Kabsch fits as doubles ([3][3] torso, then [side][5][3] wrist): converted once per CTA, not per frame
HRT_HD inline int pos_zd_word() { return (pos_ik_word() + 2 * (int)sizeof(ArmIkParams) / 4 + 3) / 4 * 4; }
HRT_HD inline int pos_const_words() { return pos_zd_word() + 2 * (9 + 2 * 15) + 2; }

struct PosArgs {
    long long B;
    const float* __restrict__ body_t;    // (B, 21, 3)
    const float* __restrict__ lhand_t;   // (B, 20, 3) or nullptr (UPPER)
    const float* __restrict__ rhand_t;   // (B, 20, 3) or nullptr
    const float* __restrict__ body_q;    // (B, 21, 4), mode FULL only
    float* __restrict__ out_local_q;     // (B, 31, 4) or nullptr
    float* __restrict__ out_dof;         // (B, 30) or nullptr
    float* __restrict__ out_body_gq;     // (B, 59, 4) or nullptr (mode POS)
    unsigned flags;                      // POS_CLAMP | POS_IK
    int ik_iters;
    float damping, rot_weight;
};

// ---------------------------------------------------------------------------------------------
// reference-order fp32 helpers
// ---------------------------------------------------------------------------------------------
HRT_DEV vec3 scale3_x(const vec3 v, float s) { return make_vec3(mul_rn(v.x, s), mul_rn(v.y, s), mul_rn(v.z, s)); }
HRT_DEV vec3 sub3_x(const vec3 a, const vec3 b) { return make_vec3(sub_rn(a.x, b.x), sub_rn(a.y, b.y), sub_rn(a.z, b.z)); }
// three numerators, one denominator: one refined reciprocal, then the exactly rounded quotient of each (hrt_math.cuh)
HRT_DEV vec3 div3_x(const vec3 v, float s) {
    const float y = rcp_refined(s);
    return make_vec3(div_by_rn(v.x, s, y), div_by_rn(v.y, s, y), div_by_rn(v.z, s, y));
}
HRT_DEV float sign_f(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }     // torch.sign: 0 for 0 (and for NaN here)

// transform3d.py:62-75   v - (dot(v,n) / ||n||^2) * n
HRT_DEV vec3 proj_in_plane_x(const vec3 v, const vec3 n) {
    const float nn = norm3_x(n);
    const float d = divn_rn(dot3_x(v, n), mul_rn(nn, nn));
    return sub3_x(v, scale3_x(n, d));
}

// transform3d.py:78-100   acos(clamp(v1.v2)) * sign(n.(v1 x v2)) after normalising all three
HRT_DEV float radians_between_x(vec3 v1, vec3 v2, vec3 n) {
    v1 = div3_x(v1, norm3_x(v1));
    v2 = div3_x(v2, norm3_x(v2));
    n = div3_x(n, norm3_x(n));
    const float c = fminf(fmaxf(dot3_x(v1, v2), -1.f), 1.f);
    const float ang = acosf(c);
    const float dir = dot3_x(n, cross3_x(v1, v2));
    return mul_rn(ang, sign_f(dir));
}

// theta / phi of a bone for the shoulder pitch-roll solve (plane normal y) or the
// shoulder-yaw / elbow-pitch solve (plane normal z); v already in the parent frame
#ifndef HRT_POS_ONEHOT_SHORTCUT
#define HRT_POS_ONEHOT_SHORTCUT 1
#endif
template <int PLANE>   // 1: xz-plane (normal y), 2: xy-plane (normal z)
HRT_DEV void bone_angles_x(const vec3 v, float* theta, float* phi) {
    const vec3 nrm = (PLANE == 1) ? make_vec3(0.f, 1.f, 0.f) : make_vec3(0.f, 0.f, 1.f);
#if HRT_POS_ONEHOT_SHORTCUT
    // proj_in_plane_x(v, e_k) and radians_between_x(e_x, vp, e_k) with the one-hot operands folded by hand (the compiler
    // cannot: x * 0 is not 0 for NaN / inf, and the reciprocal is inline PTX).  Every dropped operation is exact --
    // a / 1, a * 1, a * 0 = +-0, a + (+-0) -- so the values are the reference's own up to the sign of a zero, which
    // acos, the sign test and the products downstream do not see; a NaN in v reaches every component of v (it comes out
    // of a quaternion rotation), so NaN in -> NaN out still holds.
    const vec3 vp = (PLANE == 1) ? make_vec3(v.x, 0.f, v.z) : make_vec3(v.x, v.y, 0.f);
    {
        const vec3 v2 = div3_x(vp, norm3_x(vp));
        const float ang = acosf(fminf(fmaxf(v2.x, -1.f), 1.f));
        *theta = mul_rn(ang, sign_f((PLANE == 1) ? -v2.z : v2.y));
    }
#else
    const vec3 ex = make_vec3(1.f, 0.f, 0.f);
    const vec3 vp = proj_in_plane_x(v, nrm);
    *theta = radians_between_x(ex, vp, nrm);
#endif
    // shoulderPR: n = cross(v_proj, y);  elbow: n = cross(z, v_proj)
    const vec3 n2 = (PLANE == 1) ? cross3_x(vp, nrm) : cross3_x(nrm, vp);
    *phi = radians_between_x(vp, v, n2);
}

// rotation3d.py:147-193 in the reference's order, including the four sequential masked fix-ups
HRT_DEV float4 quat_from_rotation_matrix_x(const float m[3][3]) {
    const float d0 = m[0][0], d1 = m[1][1], d2 = m[2][2];
    // x / 4 is x * 0.25 bit for bit (a power of two)
    float w = sqrtn_rn(fmaxf(mul_rn(add_rn(add_rn(add_rn(d0, d1), d2), 1.f), 0.25f), 0.f));
    float x = sqrtn_rn(fmaxf(mul_rn(add_rn(sub_rn(sub_rn(d0, d1), d2), 1.f), 0.25f), 0.f));
    float y = sqrtn_rn(fmaxf(mul_rn(add_rn(sub_rn(add_rn(-d0, d1), d2), 1.f), 0.25f), 0.f));
    float z = sqrtn_rn(fmaxf(mul_rn(add_rn(add_rn(sub_rn(-d0, d1), d2), 1.f), 0.25f), 0.f));
    if ((w >= x) && (w >= y) && (w >= z)) {
        x = mul_rn(x, sign_f(sub_rn(m[2][1], m[1][2])));
        y = mul_rn(y, sign_f(sub_rn(m[0][2], m[2][0])));
        z = mul_rn(z, sign_f(sub_rn(m[1][0], m[0][1])));
    }
    if ((x >= w) && (x >= y) && (x >= z)) {
        w = mul_rn(w, sign_f(sub_rn(m[2][1], m[1][2])));
        y = mul_rn(y, sign_f(add_rn(m[1][0], m[0][1])));
        z = mul_rn(z, sign_f(add_rn(m[0][2], m[2][0])));
    }
    if ((y >= w) && (y >= x) && (y >= z)) {
        w = mul_rn(w, sign_f(sub_rn(m[0][2], m[2][0])));
        x = mul_rn(x, sign_f(add_rn(m[1][0], m[0][1])));
        z = mul_rn(z, sign_f(add_rn(m[2][1], m[1][2])));
    }
    if ((z >= w) && (z >= x) && (z >= y)) {
        w = mul_rn(w, sign_f(sub_rn(m[1][0], m[0][1])));
        x = mul_rn(x, sign_f(add_rn(m[2][0], m[0][2])));
        y = mul_rn(y, sign_f(add_rn(m[2][1], m[1][2])));
    }
    return quat_normalize_x(make_float4(x, y, z, w));
}

// ---------------------------------------------------------------------------------------------
// Kabsch (transform3d.py:32-50): A = M^T Z (3x3), R = U diag(1,1,det) V^T.  fp64.
// ---------------------------------------------------------------------------------------------
// One Jacobi rotation annihilating a_pq (the inner rotation, |theta| <= pi/4), branch-free and without a division or a
// square root, so that independent problems interleave and nothing leaves the fp64 FMA pipe but two MUFU seeds:
//   tau = (a_qq - a_pp)/2,  r = hypot(tau, a_pq),  cos 2theta = |tau| / r,  sin 2theta = sign(tau) a_pq / r,
//   c = sqrt(h), h = (1 + cos 2theta)/2 in [1/2, 1],  s = sin 2theta / (2c),  t a_pq = sign(tau) (r - |tau|)
// (the last identity replaces t = s/c: the diagonal moves by exactly the amount that keeps the trace).
HRT_DEV void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq,
                        double& v0p, double& v0q, double& v1p, double& v1q, double& v2p, double& v2q) {
    const double tau = 0.5 * (aqq - app);
    const double at = fabs(tau);
    const double x = fma(tau, tau, apq * apq);
    const bool live = x > 1e-290;                       // a_pq = 0 and a_pp = a_qq: nothing to rotate
    const double ri = drsqrt_n(live ? x : 1.0);
    const double r = x * ri;
    const double h = live ? fma(at * ri, 0.5, 0.5) : 1.0;
    const double ci = drsqrt_n(h);
    const double c = h * ci;
    const double s = (apq * ri) * (ci * copysign(0.5, tau));
    const double d = copysign(r - at, tau);
    app -= d; aqq += d; apq = 0.0;
    const double rp = c * arp - s * arq, rq = s * arp + c * arq;
    arp = rp; arq = rq;
    double a, b;
    a = c * v0p - s * v0q; b = s * v0p + c * v0q; v0p = a; v0q = b;
    a = c * v1p - s * v1q; b = s * v1p + c * v1q; v1p = a; v1q = b;
    a = c * v2p - s * v2q; b = s * v2p + c * v2q; v2p = a; v2q = b;
}

// Sweeps stop when the squared off-diagonal mass is below this fraction of the squared trace: off / trace <= 1e-11 leaves
// the rotation 1e-11 from the converged one, four orders below the fp32 rounding of its elements (it used to be 1e-36,
// i.e. 1e-18: one more sweep for 78 % of the wrist fits and, a warp running as long as its slowest lane, for every warp)
#ifndef HRT_KABSCH_OFF_TOL
#define HRT_KABSCH_OFF_TOL 1e-22
#endif
// NK independent Kabsch problems solved together (their fp64 dependency chains interleave: the torso
// and the wrist fit of one arm cost little more than one).  A = M^T Z in fp64 (exact products of fp32
// inputs) -> rotation quaternions.  Cyclic Jacobi on A^T A converges quadratically: sweeps stop when the
// off-diagonal mass is below 1e-11 of the trace for every problem (2 sweeps on the hand fits; 8 at most).
template <int NK>
HRT_DEV void kabsch_multi(const double (*A)[3][3], float4* out) {
    double s00[NK], s01[NK], s02[NK], s11[NK], s12[NK], s22[NK];
    double v00[NK], v01[NK], v02[NK], v10[NK], v11[NK], v12[NK], v20[NK], v21[NK], v22[NK];
#pragma unroll
    for (int n = 0; n < NK; ++n) {
        s00[n] = s01[n] = s02[n] = s11[n] = s12[n] = s22[n] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s00[n] = fma(A[n][i][0], A[n][i][0], s00[n]); s01[n] = fma(A[n][i][0], A[n][i][1], s01[n]);
            s02[n] = fma(A[n][i][0], A[n][i][2], s02[n]); s11[n] = fma(A[n][i][1], A[n][i][1], s11[n]);
            s12[n] = fma(A[n][i][1], A[n][i][2], s12[n]); s22[n] = fma(A[n][i][2], A[n][i][2], s22[n]);
        }
        v00[n] = 1; v01[n] = 0; v02[n] = 0; v10[n] = 0; v11[n] = 1; v12[n] = 0; v20[n] = 0; v21[n] = 0; v22[n] = 1;
    }
    for (int sweep = 0; sweep < 8; ++sweep) {
        bool done = true;
#pragma unroll
        for (int n = 0; n < NK; ++n) {
            const double off = fma(s01[n], s01[n], fma(s02[n], s02[n], s12[n] * s12[n]));
            const double tr = s00[n] + s11[n] + s22[n];
            done = done && !(off > HRT_KABSCH_OFF_TOL * tr * tr);
        }
        if (done) break;
#pragma unroll
        for (int n = 0; n < NK; ++n) jacobi_rot(s00[n], s11[n], s01[n], s02[n], s12[n], v00[n], v01[n], v10[n], v11[n], v20[n], v21[n]);
#pragma unroll
        for (int n = 0; n < NK; ++n) jacobi_rot(s00[n], s22[n], s02[n], s01[n], s12[n], v00[n], v02[n], v10[n], v12[n], v20[n], v22[n]);
#pragma unroll
        for (int n = 0; n < NK; ++n) jacobi_rot(s11[n], s22[n], s12[n], s01[n], s02[n], v01[n], v02[n], v11[n], v12[n], v21[n], v22[n]);
    }
#pragma unroll
    for (int n = 0; n < NK; ++n) {
        // order the eigenvalues: (a) largest, (b) second; only these two singular pairs are used
        const double e[3] = {s00[n], s11[n], s22[n]};
        const double V[3][3] = {{v00[n], v10[n], v20[n]}, {v01[n], v11[n], v21[n]}, {v02[n], v12[n], v22[n]}};
        int ia = 0, ib = 1, ic = 2;
        if (e[ib] > e[ia]) { int t = ia; ia = ib; ib = t; }
        if (e[ic] > e[ia]) { int t = ia; ia = ic; ic = t; }
        if (e[ic] > e[ib]) { int t = ib; ib = ic; ic = t; }
        double va[3], vb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            va[k] = (ia == 0) ? V[0][k] : (ia == 1) ? V[1][k] : V[2][k];
            vb[k] = (ib == 0) ? V[0][k] : (ib == 1) ? V[1][k] : V[2][k];
        }
        double ua[3], ub[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            ua[i] = A[n][i][0] * va[0] + A[n][i][1] * va[1] + A[n][i][2] * va[2];
            ub[i] = A[n][i][0] * vb[0] + A[n][i][1] * vb[1] + A[n][i][2] * vb[2];
        }
        const double na = drsqrt_n(ua[0] * ua[0] + ua[1] * ua[1] + ua[2] * ua[2]);
        ua[0] *= na; ua[1] *= na; ua[2] *= na;
        const double d = ua[0] * ub[0] + ua[1] * ub[1] + ua[2] * ub[2];
        ub[0] -= d * ua[0]; ub[1] -= d * ua[1]; ub[2] -= d * ua[2];
        const double nb = drsqrt_n(ub[0] * ub[0] + ub[1] * ub[1] + ub[2] * ub[2]);
        ub[0] *= nb; ub[1] *= nb; ub[2] *= nb;
        const double uc[3] = {ua[1] * ub[2] - ua[2] * ub[1], ua[2] * ub[0] - ua[0] * ub[2], ua[0] * ub[1] - ua[1] * ub[0]};
        const double vc[3] = {va[1] * vb[2] - va[2] * vb[1], va[2] * vb[0] - va[0] * vb[2], va[0] * vb[1] - va[1] * vb[0]};
        float R[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) R[i][k] = (float)(ua[i] * va[k] + ub[i] * vb[k] + uc[i] * vc[k]);
        out[n] = quat_from_rotation_matrix_x(R);
    }
}

HRT_DEV float4 kabsch_quat_from_A(const double A[3][3]) {
    float4 q;
    kabsch_multi<1>(reinterpret_cast<const double (*)[3][3]>(A), &q);
    return q;
}

template <int N>
HRT_DEV void kabsch_accumulate(const vec3* M, const vec3* Z, double A[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) A[i][k] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double m[3] = {(double)M[n].x, (double)M[n].y, (double)M[n].z};
        const double z[3] = {(double)Z[n].x, (double)Z[n].y, (double)Z[n].z};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) A[i][k] = fma(m[i], z[k], A[i][k]);
    }
}

// the same with the zero-pose offsets already widened (the widening is exact: same sums)
template <int N>
HRT_DEV void kabsch_accumulate_zd(const vec3* M, const double* Zd, double A[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) A[i][k] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
        const double m[3] = {(double)M[n].x, (double)M[n].y, (double)M[n].z};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) A[i][k] = fma(m[i], Zd[n * 3 + k], A[i][k]);
    }
}

// M: measured offsets (n x 3), Z: zero-pose offsets (n x 3), both fp32.  Returns the quaternion.
template <int N>
HRT_DEV float4 kabsch_quat(const vec3* M, const vec3* Z) {
    double A[3][3];
    kabsch_accumulate<N>(M, Z, A);
    return kabsch_quat_from_A(A);
}

HRT_DEV vec3 ld3(const float* p) { return make_vec3(p[0], p[1], p[2]); }

// staging sizes (words per warp of 16 frames)
HRT_HD inline int pos_in_words(int n_body, int n_hand, bool hands, bool quats) {
    return BQ_FRAMES_PER_WARP * (n_body * 3 + (hands ? 2 * n_hand * 3 : 0) + (quats ? 21 * 4 : 0));
}
// The local rotations and the body quaternions are identity except for a few rows per frame: they are written straight to
// HBM (coalesced identity fill, then each lane patches its own rows), not staged, so the tile is inputs + dof for every
// combination of outputs and the large CTAs fit whatever the caller asks for.
HRT_HD inline int pos_tile_words(const PosParams& pp, bool /*with_lq*/, bool /*with_bq*/) {
    const bool hands = pp.mode == POS_FULL_BODY_POS || pp.mode == POS_FULL_BODY;
    int in = pos_in_words(pp.n_body, pp.n_hand, hands, pp.mode == POS_FULL_BODY || pp.mode == POS_MAIN);
    return (in + 3) / 4 * 4 + bq_dof_words(pp.J_rob);
}
// Warps per CTA (one CTA per SM).  The closed form is a chain of fp64 Jacobi sweeps and fp32 libm calls: latency-bound
// with 8 warps (ncu: issue slots 34 % busy, 2.2 warps per issue in fixed-latency waits); 12 / 16 warps measured
// 1.30x / 1.47x on the dof-only batch (184 / 164 / 128 registers per thread).  The host picks the largest count whose
// staging tiles fit in shared memory.
constexpr int POS_WARPS_MIN = 8;
constexpr int POS_WARPS_MID = 12;
constexpr int POS_WARPS_MAX = 16;

// host-visible (mapped, pinned) inputs must bypass the caches when a resident kernel re-reads them per frame
template <bool SYSMEM>
HRT_DEV void pos_stage_span(float* dst, const float* src, int n_words, int lane) {
    if (SYSMEM) {
        // every PCIe read is ~2 us: issue all of a lane's loads before the first use (one frame is <= 96 words)
        if (n_words <= 96) {
            float v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = (lane + 32 * k < n_words) ? __ldcv(src + lane + 32 * k) : 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) if (lane + 32 * k < n_words) dst[lane + 32 * k] = v[k];
        } else {
            for (int i = lane; i < n_words; i += 32) dst[i] = __ldcv(src + i);
        }
    } else if (n_words == BQ_FRAMES_PER_WARP * 63) {      // a full group of 21-point body rows
        warp_span_g2s_n4<BQ_FRAMES_PER_WARP * 63 / 4>(dst, src, lane);
    } else if (n_words == BQ_FRAMES_PER_WARP * 60) {      // ... of 20-point hand rows
        warp_span_g2s_n4<BQ_FRAMES_PER_WARP * 60 / 4>(dst, src, lane);
    } else {
        warp_span_g2s(dst, src, n_words, lane);
    }
}
// the resident server runs ONE warp: no cross-warp instruction-fetch alignment there
// HRT_POS_ALIGN_GROUP: warps of a scheduler that re-align together at phase boundaries (the unrolled bodies exceed the
// instruction caches, so warps that share a scheduler should fetch the same lines).  4 = all of a 16-warp CTA's warps of
// that scheduler, one instruction stream; 2 (default) = two pairs per scheduler that start HRT_POS_PAIR_SKEW_NS apart, so
// that one pair's fp64 Jacobi phase (fp64-pipe-bound: all four warps in it at once saturate the pipe while the fp32
// pipes idle) runs under the other pair's fp32 phases.  Measured on config 3p / 2^20 frames (profiles/r02_notes.md):
// 4: 0.1201 / 0.3888 ms, pairs without skew: no change, pairs 5 us apart: 0.1146 / 0.3830 ms (3-8 us apart are within 3 % of each other; 6.5 us is the default), no alignment: +2 %.
#ifndef HRT_POS_ALIGN_GROUP
#define HRT_POS_ALIGN_GROUP 2
#endif
#ifndef HRT_POS_PAIR_SKEW_NS
#define HRT_POS_PAIR_SKEW_NS 6500
#endif
template <bool SYSMEM, int WARPS>
HRT_DEV void pos_align(int warp) {
    if (SYSMEM) return;
    if (WARPS == 16 && HRT_POS_ALIGN_GROUP == 2) {
        // warps w and w + 8 share a scheduler (w & 3) and a pair ((w >> 2) & 1)
        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + (warp & 7)) : "memory");
    } else if (HRT_POS_ALIGN_GROUP != 0) {
        smsp_align<WARPS>(warp);
    }
}

// CTA-shared constants: both PosArm tables + the zero-pose bone angles (once per CTA)
HRT_DEV void pos_setup(const PosParams& pp, float* smem) {
    float* zero_ang = smem + pos_zero_ang_word();         // [side][4] = theta0_sh, phi0_sh, theta0_el, phi0_el
    const float* src = reinterpret_cast<const float*>(&pp.arm[0]);
    for (int i = threadIdx.x; i < 2 * (int)sizeof(PosArm) / 4; i += blockDim.x) smem[i] = src[i];
    const float* iks = reinterpret_cast<const float*>(&pp.ik[0]);
    for (int i = threadIdx.x; i < 2 * (int)sizeof(ArmIkParams) / 4; i += blockDim.x) smem[pos_ik_word() + i] = iks[i];
    if (threadIdx.x < 8) zero_ang[threadIdx.x] = pp.zero_ang[threadIdx.x >> 2][threadIdx.x & 3];
    double* zd = reinterpret_cast<double*>(smem + pos_zd_word());
    for (int i = threadIdx.x; i < 9 + 30; i += blockDim.x) {             // (the resident server runs this with one warp)
        const int w = i - 9, sd = w / 15, r = w % 15;
        zd[i] = i < 9 ? (double)pp.ztorso[i / 3][i % 3] : (double)pp.arm[sd].zwrist[r / 3][r % 3];
    }
    __syncthreads();
}

// one thread per arm; launched by hrt_configure_pos, the result goes into PosParams::zero_ang
__global__ void pos_zero_angles_kernel(const PosArm a0, const PosArm a1, float* out) {
    if (threadIdx.x < 2) {
        const PosArm& ar = threadIdx.x == 0 ? a0 : a1;
        float t, p;
        bone_angles_x<1>(make_vec3(ar.v0_upper[0], ar.v0_upper[1], ar.v0_upper[2]), &t, &p);
        out[threadIdx.x * 4 + 0] = t; out[threadIdx.x * 4 + 1] = p;
        bone_angles_x<2>(make_vec3(ar.v0_lower[0], ar.v0_lower[1], ar.v0_lower[2]), &t, &p);
        out[threadIdx.x * 4 + 2] = t; out[threadIdx.x * 4 + 3] = p;
    }
}

// all frame groups of `a` that fall to CTA `cta` of `n_ctas`
template <int MODE, bool SYSMEM, int WARPS>
HRT_DEV void pos_process(const PosParams& pp, const PosArgs& a, float* smem, int n_ctas, int cta) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fl = lane >> 1;
    const int side = lane & 1;
    constexpr bool HANDS = MODE == POS_FULL_BODY_POS || MODE == POS_FULL_BODY;
    constexpr bool QUATS = MODE == POS_FULL_BODY || MODE == POS_MAIN;
    const PosArm* arms_s = reinterpret_cast<const PosArm*>(smem);
    const float* zero_ang = smem + pos_zero_ang_word();
    const ArmIkParams& ik = reinterpret_cast<const ArmIkParams*>(smem + pos_ik_word())[side];
    const PosArm& ap = arms_s[side];
    const int const_words = pos_const_words();
    const double* zd_s = reinterpret_cast<const double*>(smem + pos_zd_word());
    const bool with_lq = a.out_local_q != nullptr;
    const bool with_bq = (MODE == POS_FULL_BODY_POS) && a.out_body_gq != nullptr;
    float* tile = smem + const_words + warp * pos_tile_words(pp, with_lq, with_bq);
    const int NB = pp.n_body, NH = pp.n_hand, JR = pp.J_rob, D = JR - 1;
    const int in_words = pos_in_words(NB, NH, HANDS, QUATS);
    float* dof_t = tile + (in_words + 3) / 4 * 4;
    // input sub-regions (each a contiguous image of 16 rows)
    float* body_s = tile;
    float* lh_s = body_s + BQ_FRAMES_PER_WARP * NB * 3;
    float* rh_s = lh_s + BQ_FRAMES_PER_WARP * NH * 3;
    float* bodyq_s = HANDS ? rh_s + BQ_FRAMES_PER_WARP * NH * 3 : lh_s;
    bool pending_store = false;

    const long long n_groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    const long long total_warps = (long long)n_ctas * WARPS;
    const long long rounds = (n_groups + total_warps - 1) / total_warps;
    // the second pair of every scheduler starts a third of a round late (clips of a few rounds would only pay for it)
    if (!SYSMEM && WARPS == 16 && HRT_POS_ALIGN_GROUP == 2 && HRT_POS_PAIR_SKEW_NS > 0 && rounds >= 4 && ((warp >> 2) & 1))
        __nanosleep(HRT_POS_PAIR_SKEW_NS);
    for (long long rnd = 0; rnd < rounds; ++rnd) {
        const long long grp_raw = rnd * total_warps + (long long)cta * WARPS + warp;
        const bool live = grp_raw < n_groups;
        const long long grp = live ? grp_raw : n_groups - 1;
        const long long f0 = grp * BQ_FRAMES_PER_WARP;
        const int nld = (int)min((long long)BQ_FRAMES_PER_WARP, a.B - f0);
        const int nfr = live ? nld : 0;
        const int fr = min(fl, nld - 1);

        if (pending_store) {
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            pending_store = false;
        }
        // ---- 1. stage inputs (contiguous spans) -------------------------------------------------
        if (SYSMEM && HANDS && nld == 1 && NB * 3 <= 96 && NH * 3 <= 96) {
            // resident server, one frame in host memory: all PCIe reads of the frame in flight together
            const float* src[3] = {a.body_t, a.lhand_t, a.rhand_t};
            float* dst[3] = {body_s, lh_s, rh_s};
            const int cnt[3] = {NB * 3, NH * 3, NH * 3};
            float v[3][3];
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int k = 0; k < 3; ++k) v[sp][k] = (lane + 32 * k < cnt[sp]) ? __ldcv(src[sp] + lane + 32 * k) : 0.f;
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int k = 0; k < 3; ++k) if (lane + 32 * k < cnt[sp]) dst[sp][lane + 32 * k] = v[sp][k];
        } else {
            pos_stage_span<SYSMEM>(body_s, a.body_t + f0 * NB * 3, nld * NB * 3, lane);
            if (HANDS) {
                pos_stage_span<SYSMEM>(lh_s, a.lhand_t + f0 * NH * 3, nld * NH * 3, lane);
                pos_stage_span<SYSMEM>(rh_s, a.rhand_t + f0 * NH * 3, nld * NH * 3, lane);
            }
        }
        if (QUATS) pos_stage_span<SYSMEM>(bodyq_s, a.body_q + f0 * pp.n_bodyq * 4, nld * pp.n_bodyq * 4, lane);
        cp_async_commit();
        // the next round's rows start their way from HBM to L2 now: the staging tile has no room for a second buffer, but
        // the copy above then meets L2 latency instead of DRAM latency one round later
        if (!SYSMEM && lane == 0 && grp_raw + total_warps < n_groups) {
            const long long fn = (grp_raw + total_warps) * BQ_FRAMES_PER_WARP;
            const int nn = (int)min((long long)BQ_FRAMES_PER_WARP, a.B - fn);
            l2_prefetch_span(a.body_t + fn * NB * 3, nn * NB * 3);
            if (HANDS) {
                l2_prefetch_span(a.lhand_t + fn * NH * 3, nn * NH * 3);
                l2_prefetch_span(a.rhand_t + fn * NH * 3, nn * NH * 3);
            }
            if (QUATS) l2_prefetch_span(a.body_q + fn * pp.n_bodyq * 4, nn * pp.n_bodyq * 4);
        }
        if (a.out_dof) {
            if (nfr == BQ_FRAMES_PER_WARP) {                     // 16 rows are whole 16-byte blocks for any D
                for (int i = lane; i < BQ_FRAMES_PER_WARP * D / 4; i += 32) reinterpret_cast<float4*>(dof_t)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                for (int i = lane; i < nfr * D; i += 32) dof_t[i] = 0.f;
            }
        }
        cp_async_wait<0>();
        __syncwarp();

        const float* body = body_s + fr * NB * 3;
        const vec3 flip = make_vec3(pp.flip[0], pp.flip[1], pp.flip[2]);
        auto bpt = [&](int j) {      // coord_transform (transform3d.py:24-29): p * dir
            const vec3 p = ld3(body + j * 3);
            return (MODE == POS_UPPER_BODY) ? make_vec3(mul_rn(p.x, flip.x), mul_rn(p.y, flip.y), mul_rn(p.z, flip.z)) : p;
        };
        // ---- 2. arm parent frame ------------------------------------------------------------------
        float4 torso = make_float4(0.f, 0.f, 0.f, 1.f);
        float4 parent;
        float4 wrist_g = make_float4(0.f, 0.f, 0.f, 1.f);
        const float* hand = HANDS ? (side == 0 ? lh_s : rh_s) + fr * NH * 3 : nullptr;
        if (QUATS) {
            parent = *reinterpret_cast<const float4*>(bodyq_s + (fr * pp.n_bodyq + ap.q_parent) * 4);
        } else {
            pos_align<SYSMEM, WARPS>(warp);
            // the torso fit and (full_body_pos) this arm's wrist fit are independent: solved together
            double A[2][3][3];
            {
                vec3 M[3];
                const vec3 org = bpt(pp.torso_org);
#pragma unroll
                for (int n = 0; n < 3; ++n) M[n] = sub3_x(bpt(pp.torso_pts[n]), org);
                kabsch_accumulate_zd<3>(M, zd_s, A[0]);
            }
            if (MODE == POS_FULL_BODY_POS) {
                vec3 M[5];
                const vec3 org = ld3(hand + pp.hand_org * 3);
#pragma unroll
                for (int n = 0; n < 5; ++n) M[n] = sub3_x(ld3(hand + pp.hand_kabsch[n] * 3), org);
                kabsch_accumulate_zd<5>(M, zd_s + 9 + side * 15, A[1]);
                float4 q[2];
                kabsch_multi<2>(A, q);
                torso = q[0];
                wrist_g = q[1];
            } else {
                kabsch_multi<1>(A, &torso);
            }
            parent = torso;
        }
        // ---- 3. shoulder pitch / roll, shoulder yaw / elbow pitch ----------------------------------
        pos_align<SYSMEM, WARPS>(warp);
        float4 rl[7];
        {
            const vec3 v_up = sub3_x(bpt(ap.b_el), bpt(ap.b_sh));
            float th1, ph1;
            bone_angles_x<1>(quat_rotate_x(quat_conj(parent), v_up), &th1, &ph1);
            rl[0] = quat_from_angle_axis_k_x(sub_rn(th1, zero_ang[side * 4 + 0]), 1);
            rl[1] = quat_from_angle_axis_k_x(sub_rn(ph1, zero_ang[side * 4 + 1]), 0);
            const float4 elbow_parent = quat_mul_x(quat_mul_x(parent, rl[0]), rl[1]);
            const vec3 v_lo = sub3_x(bpt(ap.b_wr), bpt(ap.b_el));
            bone_angles_x<2>(quat_rotate_x(quat_conj(elbow_parent), v_lo), &th1, &ph1);
            rl[2] = quat_from_angle_axis_k_x(sub_rn(th1, zero_ang[side * 4 + 2]), 2);
            rl[3] = quat_from_angle_axis_k_x(sub_rn(ph1, zero_ang[side * 4 + 3]), 1);
        }
        rl[4] = rl[5] = rl[6] = make_float4(0.f, 0.f, 0.f, 1.f);
        // ---- 4. wrist -------------------------------------------------------------------------------
        if (HANDS) {
            pos_align<SYSMEM, WARPS>(warp);
            const float4 chain = quat_mul_x(quat_mul_x(quat_mul_x(rl[0], rl[1]), rl[2]), rl[3]);
            const float4 base = (MODE == POS_FULL_BODY) ? parent : torso;
            const float4 wparent = quat_mul_norm_x(base, chain);
            if (MODE == POS_FULL_BODY) wrist_g = *reinterpret_cast<const float4*>(bodyq_s + (fr * pp.n_bodyq + ap.q_wrist) * 4);
            const float4 wlocal = quat_mul_norm_x(quat_conj(wparent), wrist_g);
            pos_align<SYSMEM, WARPS>(warp);
            double es[3], ec[3];
            euler_intrinsic_half_sincos_f64<0, 1, 2>(wlocal, es, ec);          // 'XYZ'
            rl[4] = axis_quat_from_sc(es[0], ec[0], 0);
            rl[5] = axis_quat_from_sc(es[1], ec[1], 1);
            rl[6] = axis_quat_from_sc(es[2], ec[2], 2);
        }
        // ---- 5. hinge angles + gripper -----------------------------------------------------------------
        pos_align<SYSMEM, WARPS>(warp);
        float th[9];
        th[0] = quat_to_dof_x(rl[0], 1); th[1] = quat_to_dof_x(rl[1], 0); th[2] = quat_to_dof_x(rl[2], 2);
        th[3] = quat_to_dof_x(rl[3], 1);
        th[4] = HANDS ? quat_to_dof_x(rl[4], 0) : 0.f;
        th[5] = HANDS ? quat_to_dof_x(rl[5], 1) : 0.f;
        th[6] = HANDS ? quat_to_dof_x(rl[6], 2) : 0.f;
        th[7] = 0.f; th[8] = 0.f;
        if (HANDS) {
            // finger tips in the wrist frame (POS: rotate by the inverse; FULL: by the quat itself)
            const float4 q = (MODE == POS_FULL_BODY_POS) ? quat_conj(wrist_g) : wrist_g;
            const float x0 = quat_rotate_x(q, ld3(hand)).x;
            float sum = 0.f;
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                const float xn = quat_rotate_x(q, ld3(hand + pp.hand_tips[n] * 3)).x;
                const float dxn = sub_rn(xn, x0);
                sum = (n == 0) ? dxn : add_rn(sum, dxn);
            }
            const float ratio = divn_rn(divn_rn(sum, 5.f), pp.orig_x);
            if (MODE == POS_FULL_BODY_POS && pp.precise_gripper) {
                const float s = mul_rn(fminf(fmaxf(sub_rn(ratio, 0.5f), 0.f), 0.5f), 2.f);      // x / 0.5 == x * 2 exactly
                th[7] = mul_rn(s, 0.044f);
                th[8] = mul_rn(s, -0.044f);
            } else {
                const bool closed = ratio < 0.7f;
                th[7] = closed ? 0.f : 0.044f;
                th[8] = closed ? 0.f : -0.044f;
            }
        }
        // ---- 5b. joint limits / limit-aware refinement (builder-specified) ----------
        if (a.flags & (POS_CLAMP | POS_IK)) {
            const vec3 p_sh = make_vec3(ik.p_sh[0], ik.p_sh[1], ik.p_sh[2]);
            // fminf / fmaxf swallow NaN: a poisoned frame stays poisoned ("NaN in -> NaN out" also with limits / IK)
            const float nan_probe = ((th[0] + th[1]) + (th[2] + th[3])) + ((th[4] + th[5]) + th[6]);
            float thc[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) thc[c] = fminf(fmaxf(th[c], ik.lower[c]), ik.upper[c]);
            if (a.flags & POS_IK) {
                // targets: elbow / wrist positions and wrist orientation of the UNCLAMPED closed-form pose
                vec3 ax[7], pc[7];
                mat3c H;
                arm_chain_m(th, p_sh, ik.off, ax, pc, H);
                const vec3 pe_t = pc[3], pw_t = pc[6];
                const float lam2 = a.damping * a.damping;
                for (int it = 0; it < a.ik_iters; ++it) {
                    pos_align<SYSMEM, WARPS>(warp);
                    ik_step_f(thc, p_sh, ik.off, ik.lower, ik.upper, pe_t, pw_t, H, lam2, a.rot_weight, true);
                }
            }
#pragma unroll
            for (int c = 0; c < 7; ++c) th[c] = (nan_probe != nan_probe) ? nan_probe : thc[c];
            rl[0] = arm_local_quat<0>(th[0]); rl[1] = arm_local_quat<1>(th[1]); rl[2] = arm_local_quat<2>(th[2]);
            rl[3] = arm_local_quat<3>(th[3]); rl[4] = arm_local_quat<4>(th[4]); rl[5] = arm_local_quat<5>(th[5]);
            rl[6] = arm_local_quat<6>(th[6]);
        }
        __syncwarp();                                  // every lane is done with the input rows

        // ---- 6. outputs --------------------------------------------------------------------------------
        if (a.out_dof && fl < nfr) {
            float* r = dof_t + fl * D + (ap.rob_first - 1);
#pragma unroll
            for (int c = 0; c < 9; ++c) r[c] = th[c];
        }
        // local rotations / body quaternions: identity fill of the group's contiguous span (one 512-byte store per warp
        // instruction), then every lane overwrites its own hinges (a warp barrier orders the two stores to the same rows)
        const float4 ident = make_float4(0.f, 0.f, 0.f, 1.f);
        float4* lq_g = with_lq ? reinterpret_cast<float4*>(a.out_local_q + f0 * JR * 4) : nullptr;
        float4* bq_g = with_bq ? reinterpret_cast<float4*>(a.out_body_gq + f0 * pp.J_bq * 4) : nullptr;
        if (with_lq) for (int i = lane; i < nfr * JR; i += 32) lq_g[i] = ident;
        if (with_bq) for (int i = lane; i < nfr * pp.J_bq; i += 32) bq_g[i] = ident;
        if (with_lq || with_bq) __syncwarp();
        if (fl < nfr) {
            if (with_lq) {
#pragma unroll
                for (int c = 0; c < 7; ++c) lq_g[fl * JR + ap.rob_first + c] = rl[c];
            }
            if (with_bq) {
                if (side == 0) bq_g[fl * pp.J_bq + pp.bq_torso] = torso;
                bq_g[fl * pp.J_bq + ap.bq_wrist] = wrist_g;
            }
        }
        if (nfr == BQ_FRAMES_PER_WARP) {
            if (a.out_dof) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    bulk_store_s2g(a.out_dof + f0 * D, dof_t, (unsigned)(BQ_FRAMES_PER_WARP * D * 4));
                    bulk_commit();
                }
                pending_store = true;
            }
        } else if (nfr > 0) {
            __syncwarp();
            if (a.out_dof) warp_store_span(a.out_dof + f0 * D, dof_t, nfr * D, lane);
            __syncwarp();
        }
    }
    if (pending_store && lane == 0) bulk_wait_read_all();
}

template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
pos_retarget_kernel(const __grid_constant__ PosParams pp, const PosArgs a) {
    extern __shared__ __align__(16) float smem[];
    pos_setup(pp, smem);
    pos_process<MODE, false, WARPS>(pp, a, smem, (int)gridDim.x, (int)blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Resident single-frame server for the 120 Hz teleop loop (sim_full_body_teleop.py:83-129).
// One CTA stays on one SM and polls a sequence number in mapped pinned host memory; the host posts a
// frame by writing the inputs and then the next sequence number, and spins on the answer's sequence
// number: no kernel launch, no stream synchronisation, no driver call on the per-frame path.
// ctrl words (host memory): [0] seq_in (host), [1] stop (host), [16] seq_out (device), [17] exited (device).
// The kernel leaves when told to stop or after `idle_ns` without a frame, so an abandoned context
// never holds the GPU (the host relaunches it on the next frame).
// ---------------------------------------------------------------------------------------------
HRT_DEV unsigned ld_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
HRT_DEV void st_sys(unsigned* p, unsigned v) { asm volatile("st.volatile.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
HRT_DEV unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

template <int MODE>
__global__ void __launch_bounds__(32, 1)
pos_stream_server_kernel(const __grid_constant__ PosParams pp, const PosArgs a, unsigned* ctrl, unsigned served,
                         unsigned long long idle_ns) {
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned s_cmd;
    pos_setup(pp, smem);
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_timer_ns();
            unsigned cmd = 0;
            for (;;) {
                const unsigned s = ld_sys(ctrl);
                if (s != served) { cmd = s; break; }
                if (ld_sys(ctrl + 1) != 0u || global_timer_ns() - t0 > idle_ns) break;
            }
            s_cmd = cmd;
        }
        __syncthreads();
        const unsigned cmd = s_cmd;
        __syncthreads();
        if (cmd == 0u) break;
        pos_process<MODE, true, 1>(pp, a, smem, 1, 0);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_sys(ctrl + 16, cmd);
        served = cmd;
    }
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_sys(ctrl + 17, 1u);
    }
}

}  // namespace hrt
