// Element-wise rotation algebra for sm_100a: the free functions of
//   poselib/poselib/core/rotation3d.py:15-661          (quat_*, transform_*, rot_matrix_*, exp-map family)
//   retarget/spatial_transform/transform3d.py:9-183    (quat_between_two_vecs, proj_in_plane,
//                                                       radians_between_vecs, quat_slerp, quat_to_dof_pos,
//                                                       quat_in_xyz_axis, cal_joint_quat)
// as ONE kernel family.  Every op maps n items to n items; an item is a short fp32 row (a quaternion,
// a vector, a 3x3 matrix, a 7-word transform ...) and rows are stored AoS like the reference's tensors.
//
// Memory shape.  A warp owns 32 consecutive items, so for every operand and every result the warp's
// rows are ONE contiguous span of 32*W words.  Quaternion rows (W = 4) are one 16-byte access per lane,
// coalesced as they are.  Rows of other widths (3, 7, 9 words) move global <-> shared as the warp's
// contiguous span in 16-byte pieces and each lane then picks its own row out of shared memory (odd W:
// conflict-free).  Operands that broadcast (a single quaternion, or a (J,4) table against (B,J,4)) are
// read through the read-only path with index i % period.
//
// Numerics.  The fp32 bodies use the exact-order primitives of hrt_math.cuh (one rounded operation per
// torch op, reference evaluation order), so pure multiply/add chains are bit-identical to torch CPU and
// transcendental ones differ by the libm's last ulp only.
#pragma once
#include <type_traits>

#include "hrt_math.cuh"
#include "hrt_pos.cuh"

namespace hrt {

enum RotOp {
    OP_QUAT_MUL = 0, OP_QUAT_MUL_NORM, OP_QUAT_MUL_THREE, OP_QUAT_MUL_FOUR,
    OP_QUAT_POS, OP_QUAT_ABS, OP_QUAT_UNIT, OP_QUAT_NORMALIZE, OP_QUAT_CONJUGATE,
    OP_QUAT_ROTATE, OP_QUAT_FROM_ANGLE_AXIS, OP_QUAT_FROM_ROTATION_MATRIX, OP_QUAT_ANGLE_AXIS,
    OP_QUAT_YAW_ROTATION, OP_TRANSFORM_INVERSE, OP_TRANSFORM_MUL, OP_TRANSFORM_APPLY,
    OP_ROT_MATRIX_DET, OP_ROT_MATRIX_FROM_QUATERNION, OP_PROJECT_QUAT_TO_AXIS, OP_EXTRACT_ROTATION_ALONG_AXIS,
    OP_NORMALIZE_ANGLE, OP_QUAT_TO_ANGLE_AXIS, OP_QUAT_TO_EXP_MAP, OP_EXP_MAP_TO_ANGLE_AXIS, OP_EXP_MAP_TO_QUAT,
    OP_ANGLE_AXIS_TO_EXP_MAP, OP_QUAT_BETWEEN_TWO_VECS, OP_PROJ_IN_PLANE, OP_RADIANS_BETWEEN_VECS,
    OP_QUAT_SLERP, OP_QUAT_TO_DOF_POS, OP_EULER_SPLIT, OP_EULER_ANGLES_F64, OP_COORD_TRANSFORM,
    OP_CAL_SHOULDER_PR, OP_CAL_ELBOWP_SHOULDERY,
    OP_COUNT
};

struct RotOpArgs {
    long long n;
    const float* in[4];
    long long period[4];     // 0: the operand has n rows; p > 0: it has p rows, row index = i % p
    float* out[3];
    int iparam;              // axis / sequence code / flag, per op
    float fparam;
};

constexpr int ROT_WARPS = 8;

// ---------------------------------------------------------------------------------------------
// exact-order helpers not already in hrt_math.cuh
// ---------------------------------------------------------------------------------------------
HRT_DEV float norm4_x(const float4 q) {
    return sqrt_rn(add_rn(add_rn(add_rn(mul_rn(q.x, q.x), mul_rn(q.y, q.y)), mul_rn(q.z, q.z)), mul_rn(q.w, q.w)));
}
HRT_DEV float4 quat_pos_x(float4 q) {
    if (q.w < 0.f) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    return q;
}
HRT_DEV float4 quat_unit_x(const float4 q) {
    const float n = fmaxf(norm4_x(q), 1e-9f);
    return make_float4(div_rn(q.x, n), div_rn(q.y, n), div_rn(q.z, n), div_rn(q.w, n));
}
HRT_DEV float normalize_angle_x(float a) {
    float s, c;
    sincosf(a, &s, &c);
    return atan2f(s, c);
}
// rotation3d.py:123-143 with a general axis
HRT_DEV float4 quat_from_angle_axis_x(float angle, vec3 axis) {
    const float th = mul_rn(angle, 0.5f);
    const float n = fmaxf(norm3_x(axis), 1e-9f);
    axis = div3_x(axis, n);
    float s, c;
    sincosf(th, &s, &c);
    return quat_normalize_x(make_float4(mul_rn(axis.x, s), mul_rn(axis.y, s), mul_rn(axis.z, s), c));
}
// rotation3d.py:588-608
HRT_DEV void quat_to_angle_axis_x(const float4 q, float* angle, vec3* axis) {
    const float sin_theta = sqrt_rn(sub_rn(1.f, mul_rn(q.w, q.w)));
    const float a = normalize_angle_x(mul_rn(2.f, acosf(q.w)));
    const bool mask = fabsf(sin_theta) > 1e-5f;
    *angle = mask ? a : 0.f;
    *axis = mask ? make_vec3(div_rn(q.x, sin_theta), div_rn(q.y, sin_theta), div_rn(q.z, sin_theta)) : make_vec3(0.f, 0.f, 1.f);
}
// rotation3d.py:630-646
HRT_DEV void exp_map_to_angle_axis_x(const vec3 e, float* angle, vec3* axis) {
    const float n = norm3_x(e);
    const vec3 ax = div3_x(e, n);
    const float a = normalize_angle_x(n);
    const bool mask = fabsf(a) > 1e-5f;
    *angle = mask ? a : 0.f;
    *axis = mask ? ax : make_vec3(0.f, 0.f, 1.f);
}
// atan2(2 (w q_k + q_a q_b), 1 - 2 (q_k^2 + q_c^2)) as written in rotation3d.py:480-556
HRT_DEV float axis_angle_of_x(const float4 q, int k) {
    const float c[4] = {q.x, q.y, q.z, q.w};
    // k = 0: (w x + y z, x^2 + z^2)   k = 1: (w y + x z, y^2 + z^2)   k = 2: (w z + x y, z^2 + y^2)
    const int a = (k == 0) ? 1 : 0, b = (k == 2) ? 1 : 2, s = (k == 2) ? 1 : 2;
    const float num = mul_rn(2.f, add_rn(mul_rn(c[3], c[k]), mul_rn(c[a], c[b])));
    const float den = sub_rn(1.f, mul_rn(2.f, add_rn(mul_rn(c[k], c[k]), mul_rn(c[s], c[s]))));
    return atan2f(num, den);
}
HRT_DEV float4 axis_quat_x(float angle, int k) {
    float s, c;
    sincosf(mul_rn(angle, 0.5f), &s, &c);
    return make_float4(k == 0 ? s : 0.f, k == 1 ? s : 0.f, k == 2 ? s : 0.f, c);
}

// SciPy's Euler-angle extraction for any of the 24 sequences, fp64 (Bernardes & Viollet 2022;
// scipy/spatial/transform/_rotation_xp.py:365-401,1052-1118).  code: bits 0-1 / 2-3 / 4-5 = the three
// axes as written, bit 6 = extrinsic (lower-case sequence).
HRT_DEV void euler_general_f64(const float4 qf, int code, double ang[3]) {
    double q[4] = {(double)qf.x, (double)qf.y, (double)qf.z, (double)qf.w};
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
    const bool extrinsic = (code >> 6) & 1;
    int i = code & 3, j = (code >> 2) & 3, k = (code >> 4) & 3;
    if (!extrinsic) { const int t = i; i = k; k = t; }
    const bool symmetric = i == k;
    if (symmetric) k = 3 - i - j;
    const double sign = (double)((i - j) * (j - k) * (k - i) / 2);
    double a, b, c, d;
    if (symmetric) { a = q[3]; b = q[i]; c = q[j]; d = q[k] * sign; }
    else { a = q[3] - q[j]; b = q[i] + q[k] * sign; c = q[j] + q[3]; d = q[k] * sign - q[i]; }
    const double PI = 3.141592653589793;
    const double half_sum = atan2(b, a), half_diff = atan2(d, c);
    double out[3] = {0.0, 2.0 * atan2(hypot(c, d), hypot(a, b)), 0.0};
    const int first = extrinsic ? 0 : 2, third = extrinsic ? 2 : 0;
    const bool case1 = fabs(out[1]) <= 1e-7, case2 = fabs(out[1] - PI) <= 1e-7;
    if (!(case1 || case2)) {
        out[first] = half_sum - half_diff;
        out[third] = half_sum + half_diff;
    } else {
        out[2] = 0.0;
        out[0] = case1 ? 2.0 * half_sum : 2.0 * half_diff * (extrinsic ? -1.0 : 1.0);
    }
    if (!symmetric) { out[third] *= sign; out[1] -= PI / 2; }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        if (out[m] < -PI) out[m] += 2 * PI;
        else if (out[m] > PI) out[m] -= 2 * PI;
        ang[m] = out[m];
    }
}

// ---------------------------------------------------------------------------------------------
// op table: operand / result row widths and the per-row body
// ---------------------------------------------------------------------------------------------
template <int OP> struct RotOpTraits;
#define HRT_ROT_TRAITS(OP, NI_, I0, I1, I2, I3, NO_, O0, O1, O2)                       \
    template <> struct RotOpTraits<OP> {                                               \
        static constexpr int NI = NI_, NO = NO_;                                       \
        static constexpr int WI[4] = {I0, I1, I2, I3};                                 \
        static constexpr int WO[3] = {O0, O1, O2};                                     \
    };
HRT_ROT_TRAITS(OP_QUAT_MUL, 2, 4, 4, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_MUL_NORM, 2, 4, 4, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_MUL_THREE, 3, 4, 4, 4, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_MUL_FOUR, 4, 4, 4, 4, 4, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_POS, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_ABS, 1, 4, 0, 0, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_UNIT, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_NORMALIZE, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_CONJUGATE, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_ROTATE, 2, 4, 3, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_FROM_ANGLE_AXIS, 2, 1, 3, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_FROM_ROTATION_MATRIX, 1, 9, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_ANGLE_AXIS, 1, 4, 0, 0, 0, 2, 1, 3, 0)
HRT_ROT_TRAITS(OP_QUAT_YAW_ROTATION, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_TRANSFORM_INVERSE, 1, 7, 0, 0, 0, 1, 7, 0, 0)
HRT_ROT_TRAITS(OP_TRANSFORM_MUL, 2, 7, 7, 0, 0, 1, 7, 0, 0)
HRT_ROT_TRAITS(OP_TRANSFORM_APPLY, 2, 7, 3, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_ROT_MATRIX_DET, 1, 9, 0, 0, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_ROT_MATRIX_FROM_QUATERNION, 1, 4, 0, 0, 0, 1, 9, 0, 0)
HRT_ROT_TRAITS(OP_PROJECT_QUAT_TO_AXIS, 1, 4, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_EXTRACT_ROTATION_ALONG_AXIS, 1, 4, 0, 0, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_NORMALIZE_ANGLE, 1, 1, 0, 0, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_TO_ANGLE_AXIS, 1, 4, 0, 0, 0, 2, 1, 3, 0)
HRT_ROT_TRAITS(OP_QUAT_TO_EXP_MAP, 1, 4, 0, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_EXP_MAP_TO_ANGLE_AXIS, 1, 3, 0, 0, 0, 2, 1, 3, 0)
HRT_ROT_TRAITS(OP_EXP_MAP_TO_QUAT, 1, 3, 0, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_ANGLE_AXIS_TO_EXP_MAP, 2, 1, 3, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_BETWEEN_TWO_VECS, 2, 3, 3, 0, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_PROJ_IN_PLANE, 2, 3, 3, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_RADIANS_BETWEEN_VECS, 3, 3, 3, 3, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_SLERP, 3, 4, 4, 1, 0, 1, 4, 0, 0)
HRT_ROT_TRAITS(OP_QUAT_TO_DOF_POS, 2, 4, 1, 0, 0, 1, 1, 0, 0)
HRT_ROT_TRAITS(OP_EULER_SPLIT, 1, 4, 0, 0, 0, 3, 4, 4, 4)
HRT_ROT_TRAITS(OP_EULER_ANGLES_F64, 1, 4, 0, 0, 0, 1, 6, 0, 0)       // three doubles
HRT_ROT_TRAITS(OP_COORD_TRANSFORM, 2, 3, 3, 0, 0, 1, 3, 0, 0)
HRT_ROT_TRAITS(OP_CAL_SHOULDER_PR, 3, 3, 3, 4, 0, 2, 4, 4, 0)          // v1, v0, parent quat -> pitch quat, roll quat
HRT_ROT_TRAITS(OP_CAL_ELBOWP_SHOULDERY, 3, 3, 3, 4, 0, 2, 4, 4, 0)     // v1, v0, parent quat -> yaw quat, elbow-pitch quat
#undef HRT_ROT_TRAITS

HRT_DEV float4 ld4(const float* p) { return make_float4(p[0], p[1], p[2], p[3]); }
HRT_DEV void st4(float* p, const float4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
HRT_DEV void st3(float* p, const vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// in[k] / out[k] are the lane's rows (registers after unrolling)
template <int OP>
HRT_DEV void rot_op_body(const float (&in)[4][9], float (&out)[3][9], int ip, float fp) {
    if constexpr (OP == OP_QUAT_MUL) st4(out[0], quat_mul_x(ld4(in[0]), ld4(in[1])));
    else if constexpr (OP == OP_QUAT_MUL_NORM) st4(out[0], quat_mul_norm_x(ld4(in[0]), ld4(in[1])));
    else if constexpr (OP == OP_QUAT_MUL_THREE) st4(out[0], quat_mul_x(quat_mul_x(ld4(in[0]), ld4(in[1])), ld4(in[2])));
    else if constexpr (OP == OP_QUAT_MUL_FOUR)
        st4(out[0], quat_mul_x(quat_mul_x(quat_mul_x(ld4(in[0]), ld4(in[1])), ld4(in[2])), ld4(in[3])));
    else if constexpr (OP == OP_QUAT_POS) st4(out[0], quat_pos_x(ld4(in[0])));
    else if constexpr (OP == OP_QUAT_ABS) out[0][0] = norm4_x(ld4(in[0]));
    else if constexpr (OP == OP_QUAT_UNIT) st4(out[0], quat_unit_x(ld4(in[0])));
    else if constexpr (OP == OP_QUAT_NORMALIZE) st4(out[0], quat_normalize_x(ld4(in[0])));
    else if constexpr (OP == OP_QUAT_CONJUGATE) st4(out[0], quat_conj(ld4(in[0])));
    else if constexpr (OP == OP_QUAT_ROTATE) st3(out[0], quat_rotate_x(ld4(in[0]), ld3(in[1])));
    else if constexpr (OP == OP_QUAT_FROM_ANGLE_AXIS) {
        float a = in[0][0];
        if (ip) a = mul_rn(div_rn(a, 180.f), 3.14159265358979323846f);       // degree=True
        st4(out[0], quat_from_angle_axis_x(a, ld3(in[1])));
    } else if constexpr (OP == OP_QUAT_FROM_ROTATION_MATRIX) {
        const float m[3][3] = {{in[0][0], in[0][1], in[0][2]}, {in[0][3], in[0][4], in[0][5]}, {in[0][6], in[0][7], in[0][8]}};
        st4(out[0], quat_from_rotation_matrix_x(m));
    } else if constexpr (OP == OP_QUAT_ANGLE_AXIS) {                          // rotation3d.py:230-240
        const float4 q = ld4(in[0]);
        const float s = sub_rn(mul_rn(2.f, mul_rn(q.w, q.w)), 1.f);
        out[0][0] = acosf(fminf(fmaxf(s, -1.f), 1.f));
        const vec3 v = make_vec3(q.x, q.y, q.z);
        st3(out[1], div3_x(v, fmaxf(norm3_x(v), 1e-9f)));
    } else if constexpr (OP == OP_QUAT_YAW_ROTATION) {                        // :243-261
        const float4 q = ld4(in[0]);
        st4(out[0], quat_normalize_x(ip ? make_float4(0.f, 0.f, q.z, q.w) : make_float4(0.f, q.y, 0.f, q.w)));
    } else if constexpr (OP == OP_TRANSFORM_INVERSE) {                        // :300-306
        const float4 inv = quat_conj(ld4(in[0]));
        st4(out[0], inv);
        st3(out[0] + 4, quat_rotate_x(inv, make_vec3(-in[0][4], -in[0][5], -in[0][6])));
    } else if constexpr (OP == OP_TRANSFORM_MUL) {                            // :317-326
        const float4 rx = ld4(in[0]);
        st4(out[0], quat_mul_norm_x(rx, ld4(in[1])));
        const vec3 t = quat_rotate_x(rx, ld3(in[1] + 4));
        st3(out[0] + 4, make_vec3(add_rn(t.x, in[0][4]), add_rn(t.y, in[0][5]), add_rn(t.z, in[0][6])));
    } else if constexpr (OP == OP_TRANSFORM_APPLY) {                          // :329-335
        const vec3 t = quat_rotate_x(ld4(in[0]), ld3(in[1]));
        st3(out[0], make_vec3(add_rn(t.x, in[0][4]), add_rn(t.y, in[0][5]), add_rn(t.z, in[0][6])));
    } else if constexpr (OP == OP_ROT_MATRIX_DET) {                           // :338-350
        const float* m = in[0];
        const float t1 = mul_rn(m[0], sub_rn(mul_rn(m[4], m[8]), mul_rn(m[5], m[7])));
        const float t2 = mul_rn(m[1], sub_rn(mul_rn(m[3], m[8]), mul_rn(m[5], m[6])));
        const float t3 = mul_rn(m[2], sub_rn(mul_rn(m[3], m[7]), mul_rn(m[4], m[6])));
        out[0][0] = add_rn(sub_rn(t1, t2), t3);
    } else if constexpr (OP == OP_ROT_MATRIX_FROM_QUATERNION) {               // :398-427
        const float i = in[0][0], j = in[0][1], k = in[0][2], r = in[0][3];
        const float two_s = div_rn(2.f, add_rn(add_rn(add_rn(mul_rn(i, i), mul_rn(j, j)), mul_rn(k, k)), mul_rn(r, r)));
        out[0][0] = sub_rn(1.f, mul_rn(two_s, add_rn(mul_rn(j, j), mul_rn(k, k))));
        out[0][1] = mul_rn(two_s, sub_rn(mul_rn(i, j), mul_rn(k, r)));
        out[0][2] = mul_rn(two_s, add_rn(mul_rn(i, k), mul_rn(j, r)));
        out[0][3] = mul_rn(two_s, add_rn(mul_rn(i, j), mul_rn(k, r)));
        out[0][4] = sub_rn(1.f, mul_rn(two_s, add_rn(mul_rn(i, i), mul_rn(k, k))));
        out[0][5] = mul_rn(two_s, sub_rn(mul_rn(j, k), mul_rn(i, r)));
        out[0][6] = mul_rn(two_s, sub_rn(mul_rn(i, k), mul_rn(j, r)));
        out[0][7] = mul_rn(two_s, add_rn(mul_rn(j, k), mul_rn(i, r)));
        out[0][8] = sub_rn(1.f, mul_rn(two_s, add_rn(mul_rn(i, i), mul_rn(j, j))));
    } else if constexpr (OP == OP_PROJECT_QUAT_TO_AXIS) {                     // :479-530; ip: 0 x, 1 y, 2 z, 3 xy, 4 xz
        const float4 q = ld4(in[0]);
        if (ip <= 2) st4(out[0], axis_quat_x(axis_angle_of_x(q, ip), ip));
        else st4(out[0], quat_mul_x(axis_quat_x(axis_angle_of_x(q, 0), 0), axis_quat_x(axis_angle_of_x(q, ip == 3 ? 1 : 2), ip == 3 ? 1 : 2)));
    } else if constexpr (OP == OP_EXTRACT_ROTATION_ALONG_AXIS) out[0][0] = axis_angle_of_x(ld4(in[0]), ip);
    else if constexpr (OP == OP_NORMALIZE_ANGLE) out[0][0] = normalize_angle_x(in[0][0]);
    else if constexpr (OP == OP_QUAT_TO_ANGLE_AXIS) {
        vec3 ax;
        quat_to_angle_axis_x(ld4(in[0]), &out[0][0], &ax);
        st3(out[1], ax);
    } else if constexpr (OP == OP_QUAT_TO_EXP_MAP) {
        float a; vec3 ax;
        quat_to_angle_axis_x(ld4(in[0]), &a, &ax);
        st3(out[0], scale3_x(ax, a));
    } else if constexpr (OP == OP_EXP_MAP_TO_ANGLE_AXIS) {
        vec3 ax;
        exp_map_to_angle_axis_x(ld3(in[0]), &out[0][0], &ax);
        st3(out[1], ax);
    } else if constexpr (OP == OP_EXP_MAP_TO_QUAT) {
        float a; vec3 ax;
        exp_map_to_angle_axis_x(ld3(in[0]), &a, &ax);
        st4(out[0], quat_from_angle_axis_x(a, ax));
    } else if constexpr (OP == OP_ANGLE_AXIS_TO_EXP_MAP) st3(out[0], scale3_x(ld3(in[1]), in[0][0]));
    else if constexpr (OP == OP_QUAT_BETWEEN_TWO_VECS) {                      // transform3d.py:9-21, ip: whole-batch early-out
        if (ip) { st4(out[0], make_float4(0.f, 0.f, 0.f, 1.f)); return; }
        vec3 a = ld3(in[0]), b = ld3(in[1]);
        a = div3_x(a, norm3_x(a));
        b = div3_x(b, norm3_x(b));
        const vec3 c = cross3_x(a, b);
        st4(out[0], quat_normalize_x(make_float4(c.x, c.y, c.z, add_rn(1.f, dot3_x(a, b)))));
    } else if constexpr (OP == OP_PROJ_IN_PLANE) st3(out[0], proj_in_plane_x(ld3(in[0]), ld3(in[1])));
    else if constexpr (OP == OP_RADIANS_BETWEEN_VECS) out[0][0] = radians_between_x(ld3(in[0]), ld3(in[1]), ld3(in[2]));
    else if constexpr (OP == OP_QUAT_SLERP) {                                 // transform3d.py:152-174
        const float4 q0 = ld4(in[0]);
        float4 q1 = ld4(in[1]);
        const float t = in[2][0];
        float ch = add_rn(add_rn(add_rn(mul_rn(q0.x, q1.x), mul_rn(q0.y, q1.y)), mul_rn(q0.z, q1.z)), mul_rn(q0.w, q1.w));
        if (ch < 0.f) q1 = make_float4(-q1.x, -q1.y, -q1.z, -q1.w);
        ch = fabsf(ch);
        const float half = acosf(ch);
        const float sh = sqrt_rn(sub_rn(1.f, mul_rn(ch, ch)));
        const float ra = div_rn(sinf(mul_rn(sub_rn(1.f, t), half)), sh);
        const float rb = div_rn(sinf(mul_rn(t, half)), sh);
        float4 r = make_float4(add_rn(mul_rn(ra, q0.x), mul_rn(rb, q1.x)), add_rn(mul_rn(ra, q0.y), mul_rn(rb, q1.y)),
                               add_rn(mul_rn(ra, q0.z), mul_rn(rb, q1.z)), add_rn(mul_rn(ra, q0.w), mul_rn(rb, q1.w)));
        if (fabsf(sh) < 0.001f)
            r = make_float4(add_rn(mul_rn(0.5f, q0.x), mul_rn(0.5f, q1.x)), add_rn(mul_rn(0.5f, q0.y), mul_rn(0.5f, q1.y)),
                            add_rn(mul_rn(0.5f, q0.z), mul_rn(0.5f, q1.z)), add_rn(mul_rn(0.5f, q0.w), mul_rn(0.5f, q1.w)));
        if (fabsf(ch) >= 1.f) r = q0;
        st4(out[0], r);
    } else if constexpr (OP == OP_QUAT_TO_DOF_POS) {                          // transform3d.py:176-183; in[1] = hinge axis as a float
        out[0][0] = quat_to_dof_x(ld4(in[0]), (int)in[1][0]);
    } else if constexpr (OP == OP_EULER_SPLIT) {                              // transform3d.py:52-59
        // the three intrinsic sequences of the fused kernels go through the same half-angle routine they use
        double hs[3], hc[3];
        if (ip == (0 | (1 << 2) | (2 << 4))) {                                // 'XYZ'
            euler_intrinsic_half_sincos_f64<0, 1, 2>(ld4(in[0]), hs, hc);
            st4(out[0], axis_quat_from_sc(hs[0], hc[0], 0)); st4(out[1], axis_quat_from_sc(hs[1], hc[1], 1)); st4(out[2], axis_quat_from_sc(hs[2], hc[2], 2));
        } else if (ip == (1 | (0 << 2) | (2 << 4))) {                         // 'YXZ'
            euler_intrinsic_half_sincos_f64<1, 0, 2>(ld4(in[0]), hs, hc);
            st4(out[0], axis_quat_from_sc(hs[0], hc[0], 1)); st4(out[1], axis_quat_from_sc(hs[1], hc[1], 0)); st4(out[2], axis_quat_from_sc(hs[2], hc[2], 2));
        } else if (ip == (2 | (1 << 2) | (0 << 4))) {                         // 'ZYX'
            euler_intrinsic_half_sincos_f64<2, 1, 0>(ld4(in[0]), hs, hc);
            st4(out[0], axis_quat_from_sc(hs[0], hc[0], 2)); st4(out[1], axis_quat_from_sc(hs[1], hc[1], 1)); st4(out[2], axis_quat_from_sc(hs[2], hc[2], 0));
        } else {
            double e[3];
            euler_general_f64(ld4(in[0]), ip, e);
            st4(out[0], axis_quat_from_f64(e[0], ip & 3));
            st4(out[1], axis_quat_from_f64(e[1], (ip >> 2) & 3));
            st4(out[2], axis_quat_from_f64(e[2], (ip >> 4) & 3));
        }
    } else if constexpr (OP == OP_EULER_ANGLES_F64) {                         // rotation3d.py:658-661 (fp64 out)
        double e[3];
        euler_general_f64(ld4(in[0]), ip & 0x7f, e);
        const double k = (ip & 0x80) ? 180.0 / 3.141592653589793 : 1.0;       // degrees=True
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const double v = e[m] * k;
            out[0][2 * m] = __int_as_float(__double2loint(v));
            out[0][2 * m + 1] = __int_as_float(__double2hiint(v));
        }
    } else if constexpr (OP == OP_COORD_TRANSFORM) {                          // transform3d.py:24-29: p[..., order] * dir
        const int o0 = ip & 3, o1 = (ip >> 2) & 3, o2 = (ip >> 4) & 3;
        out[0][0] = mul_rn(in[0][o0], in[1][0]); out[0][1] = mul_rn(in[0][o1], in[1][1]); out[0][2] = mul_rn(in[0][o2], in[1][2]);
    }
    else if constexpr (OP == OP_CAL_SHOULDER_PR || OP == OP_CAL_ELBOWP_SHOULDERY) {
        // retarget_solver.py:103-158: the measured bone in the parent frame against the zero-pose bone, as angle
        // differences in the xz- (shoulder) or xy-plane (elbow); the same device code as pos_retarget_kernel
        constexpr int PLANE = (OP == OP_CAL_SHOULDER_PR) ? 1 : 2;
        const vec3 v = quat_rotate_x(quat_conj(ld4(in[2])), ld3(in[0]));
        float t1, p1, t0, p0;
        bone_angles_x<PLANE>(v, &t1, &p1);
        bone_angles_x<PLANE>(ld3(in[1]), &t0, &p0);
        st4(out[0], quat_from_angle_axis_k_x(sub_rn(t1, t0), PLANE == 1 ? 1 : 2));
        st4(out[1], quat_from_angle_axis_k_x(sub_rn(p1, p0), PLANE == 1 ? 0 : 1));
    }
    (void)fp;
}

// compile-time loop: f(std::integral_constant<int, K>) for K = 0 .. N-1
template <int K, int N, typename F>
HRT_DEV void static_for(F&& f) {
    if constexpr (K < N) {
        f(std::integral_constant<int, K>{});
        static_for<K + 1, N>(f);
    }
}
template <typename Tr, int K> __host__ __device__ constexpr int rot_in_off() {
    int o = 0;
    for (int k = 0; k < K; ++k) o += 32 * Tr::WI[k];
    return o;
}
template <typename Tr, int K> __host__ __device__ constexpr int rot_out_off() {
    int o = 32 * (Tr::WI[0] + Tr::WI[1] + Tr::WI[2] + Tr::WI[3]);
    for (int k = 0; k < K; ++k) o += 32 * Tr::WO[k];
    return o;
}

HRT_DEV bool ptr_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// contiguous span of n_words (a multiple of 4 when `vec`) global -> shared / shared -> global by the whole warp
HRT_DEV void span_load(float* tile, const float* src, int n_words, int lane, bool vec) {
    if (vec) {
        for (int w = lane; w < (n_words >> 2); w += 32)
            *reinterpret_cast<float4*>(tile + w * 4) = __ldcs(reinterpret_cast<const float4*>(src) + w);
    } else {
        for (int w = lane; w < n_words; w += 32) tile[w] = __ldcs(src + w);
    }
}
HRT_DEV void span_store(float* dst, const float* tile, int n_words, int lane, bool vec) {
    if (vec) {
        for (int w = lane; w < (n_words >> 2); w += 32)
            __stcs(reinterpret_cast<float4*>(dst) + w, *reinterpret_cast<const float4*>(tile + w * 4));
    } else {
        for (int w = lane; w < n_words; w += 32) __stcs(dst + w, tile[w]);
    }
}

// Tiles in flight per warp: a warp's tile is 32 rows, i.e. 0.4 - 1 KB of operands, and with one tile per warp and
// iteration the resident warps of an SM do not keep enough bytes in flight to cover the HBM latency when a row has ONE
// quaternion-sized operand (quat_normalize: 64 warps x 512 B = 32 KB per SM against the ~43 KB that 6.4 TB/s x 1 us /
// 148 SMs asks for).  Those ops (rows of 4 or 1 words: one LDG / STG per lane, nothing staged; at most 4 operand words)
// request the operands of ROT_UNROLL tiles before the first is consumed: quat_normalize 0.73 -> 0.80 / 0.81 / 0.83 of
// the HBM peak with 2 / 3 / 4 tiles.  Two-operand ops already have the bytes in flight (quat_mul 0.90 either way), and
// ops with 3- / 7- / 9-word rows go through the shared-memory staging, where a second tile in flight costs more than it
// hides (quat_rotate 0.59 -> 0.51): one tile.
#ifndef HRT_ROT_UNROLL
#define HRT_ROT_UNROLL 4
#endif
// a whole tile's span (32 rows of W words = 8 W sixteen-byte pieces) with compile-time trip counts: the rolled loops above
// cost ~10 instructions per piece and lane on the staged ops (quat_rotate ran 209 warp instructions per tile, 62 % of the
// issue slots: profiles/r02_notes.md)
template <int W>
HRT_DEV void span_load_tile(float* tile, const float* src, int lane) {
    constexpr int N4 = W * 8;
#pragma unroll
    for (int k = 0; k < (N4 + 31) / 32; ++k)
        if (k < N4 / 32 || lane < N4 % 32)
            reinterpret_cast<float4*>(tile)[lane + 32 * k] = __ldcs(reinterpret_cast<const float4*>(src) + lane + 32 * k);
}
template <int W>
HRT_DEV void span_store_tile(float* dst, const float* tile, int lane) {
    constexpr int N4 = W * 8;
#pragma unroll
    for (int k = 0; k < (N4 + 31) / 32; ++k)
        if (k < N4 / 32 || lane < N4 % 32)
            __stcs(reinterpret_cast<float4*>(dst) + lane + 32 * k, reinterpret_cast<const float4*>(tile)[lane + 32 * k]);
}

template <typename Tr> __host__ __device__ constexpr int rot_unroll() {
    for (int k = 0; k < 4; ++k) if (Tr::WI[k] != 0 && Tr::WI[k] != 1 && Tr::WI[k] != 4) return 1;
    for (int k = 0; k < 3; ++k) if (Tr::WO[k] != 0 && Tr::WO[k] != 1 && Tr::WO[k] != 4) return 1;
    if (Tr::WI[0] + Tr::WI[1] + Tr::WI[2] + Tr::WI[3] > 4) return 1;
    // the staging tiles stay declared (unaligned operands fall back to them): keep them inside the 48 KB of static shared memory
    const int tw = Tr::WI[0] + Tr::WI[1] + Tr::WI[2] + Tr::WI[3] + Tr::WO[0] + Tr::WO[1] + Tr::WO[2];
    const int fit = 48 * 1024 / (ROT_WARPS * 32 * tw * 4);
    return HRT_ROT_UNROLL < fit ? HRT_ROT_UNROLL : (fit < 1 ? 1 : fit);
}

template <int OP>
__global__ void __launch_bounds__(ROT_WARPS * 32)
rot_op_kernel(const RotOpArgs a) {
    using Tr = RotOpTraits<OP>;
    constexpr int TW = Tr::WI[0] + Tr::WI[1] + Tr::WI[2] + Tr::WI[3] + Tr::WO[0] + Tr::WO[1] + Tr::WO[2];
    constexpr int ROT_UNROLL = rot_unroll<Tr>();
    __shared__ __align__(16) float tiles[ROT_WARPS][ROT_UNROLL][32 * TW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n_tiles = (a.n + 31) / 32;
    const long long stride = (long long)gridDim.x * ROT_WARPS;
    // 16-byte accesses need 16-byte aligned bases (a warp's span starts at a multiple of 128 bytes from the base)
    bool al_in[4], al_out[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) al_in[k] = ptr_aligned16(a.in[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) al_out[k] = ptr_aligned16(a.out[k]);
    for (long long t0 = (long long)blockIdx.x * ROT_WARPS + warp; t0 < n_tiles; t0 += stride * ROT_UNROLL) {
        float in[ROT_UNROLL][4][9], out[3][9];
        // 1. operands of every tile of this iteration.  Width-4 rows (quaternions) and width-1 rows are coalesced as they
        //    are: one LDG.128 / LDG per lane.  Other widths: the warp's contiguous span goes through shared memory in
        //    16-byte pieces.
#pragma unroll
        for (int u = 0; u < ROT_UNROLL; ++u) {
            const long long t = t0 + u * stride;
            if (t < n_tiles) {
                float* tile = tiles[warp][u];
                const long long i0 = t * 32;
                const int cnt = (int)min(32LL, a.n - i0);
                const long long i = i0 + min(lane, cnt - 1);          // tail lanes shadow the last row
                static_for<0, Tr::NI>([&](auto K) {
                    constexpr int k = decltype(K)::value, W = Tr::WI[k], OFF = rot_in_off<Tr, k>();
                    if (a.period[k] != 0) {
                        const float* src = a.in[k] + (i % a.period[k]) * W;
#pragma unroll
                        for (int c = 0; c < W; ++c) in[u][k][c] = __ldg(src + c);
                    } else if (W == 4 && al_in[k]) {
                        const float4 v = __ldcs(reinterpret_cast<const float4*>(a.in[k]) + i);
                        in[u][k][0] = v.x; in[u][k][1] = v.y; in[u][k][2] = v.z; in[u][k][3] = v.w;
                    } else if (W == 1) {
                        in[u][k][0] = __ldcs(a.in[k] + i);
                    } else if (cnt == 32 && al_in[k]) {
                        span_load_tile<W>(tile + OFF, a.in[k] + i0 * W, lane);
                    } else {
                        span_load(tile + OFF, a.in[k] + i0 * W, cnt * W, lane, al_in[k] && ((cnt * W) & 3) == 0);
                    }
                });
            }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < ROT_UNROLL; ++u) {
            const long long t = t0 + u * stride;
            if (t >= n_tiles) break;
            float* tile = tiles[warp][u];
            const long long i0 = t * 32;
            const int cnt = (int)min(32LL, a.n - i0);
            const int row = min(lane, cnt - 1);
            const long long i = i0 + row;
            static_for<0, Tr::NI>([&](auto K) {
                constexpr int k = decltype(K)::value, W = Tr::WI[k], OFF = rot_in_off<Tr, k>();
                if (a.period[k] == 0 && !(W == 4 && al_in[k]) && W != 1) {
                    const float* r = tile + OFF + row * W;
#pragma unroll
                    for (int c = 0; c < W; ++c) in[u][k][c] = r[c];
                }
            });
            rot_op_body<OP>(in[u], out, a.iparam, a.fparam);
            // 2. results, the same way
            static_for<0, Tr::NO>([&](auto K) {
                constexpr int k = decltype(K)::value, W = Tr::WO[k], OFF = rot_out_off<Tr, k>();
                if (W == 4 && al_out[k]) {
                    if (lane < cnt) __stcs(reinterpret_cast<float4*>(a.out[k]) + i, make_float4(out[k][0], out[k][1], out[k][2], out[k][3]));
                } else if (W == 1) {
                    if (lane < cnt) __stcs(a.out[k] + i, out[k][0]);
                } else {
                    float* r = tile + OFF + lane * W;
#pragma unroll
                    for (int c = 0; c < W; ++c) r[c] = out[k][c];
                }
            });
            __syncwarp();
            static_for<0, Tr::NO>([&](auto K) {
                constexpr int k = decltype(K)::value, W = Tr::WO[k], OFF = rot_out_off<Tr, k>();
                if (!(W == 4 && al_out[k]) && W != 1) {
                    if (cnt == 32 && al_out[k]) span_store_tile<W>(a.out[k] + i0 * W, tile + OFF, lane);
                    else span_store(a.out[k] + i0 * W, tile + OFF, cnt * W, lane, al_out[k] && ((cnt * W) & 3) == 0);
                }
            });
        }
        __syncwarp();
    }
}

// max over rows of the 3-vector norm (torch.norm(v, dim=-1).max(), transform3d.py:11): the value is
// non-negative, so the float bit pattern orders like an unsigned integer
__global__ void __launch_bounds__(256)
max_norm3_kernel(const float* __restrict__ v, long long n, unsigned* __restrict__ out_bits) {
    float m = 0.f;
    bool nan = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float r = norm3_x(make_vec3(v[i * 3], v[i * 3 + 1], v[i * 3 + 2]));
        nan |= (r != r);
        m = fmaxf(m, r);
    }
    if (nan) m = __int_as_float(0x7fc00000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, m, o);
        m = (m != m || other != other) ? __int_as_float(0x7fc00000) : fmaxf(m, other);
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

// cal_joint_quat (transform3d.py:32-50), one thread per batch row, np points per row
__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ zero_t, long long zero_period, const float* __restrict__ motion_t,
              long long n, int np, float* __restrict__ out_q) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* z = zero_t + (zero_period ? i % zero_period : i) * np * 3;
        const float* m = motion_t + i * np * 3;
        double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int p = 0; p < np; ++p) {
            const double mv[3] = {(double)m[p * 3], (double)m[p * 3 + 1], (double)m[p * 3 + 2]};
            const double zv[3] = {(double)__ldg(z + p * 3), (double)__ldg(z + p * 3 + 1), (double)__ldg(z + p * 3 + 2)};
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) A[r][c] += mv[r] * zv[c];
        }
        const float4 q = kabsch_quat_from_A(A);
        *reinterpret_cast<float4*>(out_q + i * 4) = q;
    }
}

}  // namespace hrt
