// Packed fp32x2 arithmetic for sm_100a (Blackwell FFMA2 / FMUL2 / FADD2: `fma.rn.f32x2` & co.).
//
// One FFMA2 issues two fp32 FMAs per lane from 64-bit register pairs; operands may be negated, taken
// from a single 32-bit register broadcast to both halves (`R.F32`) or be immediates, all for free
// (checked in SASS).  ncu showed the fused retarget kernel issue-bound with 68 % of its dynamic
// instructions being FFMA/FMUL/FADD (profiles/r01_notes.md), so the IK refinement runs on TWO arms per
// thread: .x and .y of every value below belong to two different frames (same body side, hence the
// same per-arm tables, which enter as broadcast scalars).
#pragma once
#include "hrt_math.cuh"

namespace hrt {

typedef float2 f2;

HRT_DEV f2 dup2(float c) { return make_float2(c, c); }
HRT_DEV f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
HRT_DEV f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
HRT_DEV f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
HRT_DEV f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
HRT_DEV f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }            // a*b + c
HRT_DEV f2 fms2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, neg2(c)); }      // a*b - c
HRT_DEV f2 fnma2(f2 a, f2 b, f2 c) { return __ffma2_rn(neg2(a), b, c); }     // c - a*b
HRT_DEV f2 min2(f2 a, f2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
HRT_DEV f2 max2(f2 a, f2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
HRT_DEV f2 rsqrt2(f2 a) { return make_float2(rsqrtf(a.x), rsqrtf(a.y)); }

struct v3p { f2 x, y, z; };
struct q4p { f2 x, y, z, w; };

HRT_DEV v3p make_v3p(f2 x, f2 y, f2 z) { v3p v; v.x = x; v.y = y; v.z = z; return v; }
HRT_DEV v3p dup_v3p(const float* c) { return make_v3p(dup2(c[0]), dup2(c[1]), dup2(c[2])); }
HRT_DEV v3p add3p(const v3p a, const v3p b) { return make_v3p(add2(a.x, b.x), add2(a.y, b.y), add2(a.z, b.z)); }
HRT_DEV v3p sub3p(const v3p a, const v3p b) { return make_v3p(sub2(a.x, b.x), sub2(a.y, b.y), sub2(a.z, b.z)); }
HRT_DEV f2 dot3p(const v3p a, const v3p b) { return fma2(a.z, b.z, fma2(a.y, b.y, mul2(a.x, b.x))); }
HRT_DEV v3p cross3p(const v3p a, const v3p b) {
    return make_v3p(fms2(a.y, b.z, mul2(a.z, b.y)), fms2(a.z, b.x, mul2(a.x, b.z)), fms2(a.x, b.y, mul2(a.y, b.x)));
}

HRT_DEV q4p quat_mul_p(const q4p a, const q4p b) {
    q4p r;
    r.w = fnma2(a.z, b.z, fnma2(a.y, b.y, fnma2(a.x, b.x, mul2(a.w, b.w))));
    r.x = fnma2(a.z, b.y, fma2(a.y, b.z, fma2(a.x, b.w, mul2(a.w, b.x))));
    r.y = fnma2(a.x, b.z, fma2(a.z, b.x, fma2(a.y, b.w, mul2(a.w, b.y))));
    r.z = fnma2(a.y, b.x, fma2(a.x, b.y, fma2(a.z, b.w, mul2(a.w, b.z))));
    return r;
}
HRT_DEV q4p quat_conj_p(const q4p q) { q4p r; r.x = neg2(q.x); r.y = neg2(q.y); r.z = neg2(q.z); r.w = q.w; return r; }

// normalise and make w >= 0 (quat_normalize_f, two lanes)
HRT_DEV q4p quat_normalize_p(const q4p q) {
    const f2 n2 = fma2(q.w, q.w, fma2(q.z, q.z, fma2(q.y, q.y, mul2(q.x, q.x))));
    f2 inv = rsqrt2(max2(n2, dup2(1e-18f)));
    inv.x = (q.w.x < 0.f) ? -inv.x : inv.x;
    inv.y = (q.w.y < 0.f) ? -inv.y : inv.y;
    q4p r;
    r.x = mul2(q.x, inv); r.y = mul2(q.y, inv); r.z = mul2(q.z, inv); r.w = mul2(q.w, inv);
    return r;
}

// q * (axis-k rotation (s, c)), cf. quat_mul_axis_f
template <int K>
HRT_DEV q4p quat_mul_axis_p(const q4p a, const f2 s, const f2 c) {
    q4p r;
    if (K == 0) {
        r.w = fnma2(a.x, s, mul2(a.w, c)); r.x = fma2(a.w, s, mul2(a.x, c));
        r.y = fma2(a.z, s, mul2(a.y, c)); r.z = fnma2(a.y, s, mul2(a.z, c));
    } else if (K == 1) {
        r.w = fnma2(a.y, s, mul2(a.w, c)); r.x = fnma2(a.z, s, mul2(a.x, c));
        r.y = fma2(a.w, s, mul2(a.y, c)); r.z = fma2(a.x, s, mul2(a.z, c));
    } else {
        r.w = fnma2(a.z, s, mul2(a.w, c)); r.x = fma2(a.y, s, mul2(a.x, c));
        r.y = fnma2(a.x, s, mul2(a.y, c)); r.z = fma2(a.w, s, mul2(a.z, c));
    }
    return r;
}

// column K of R(q), cf. quat_axis_f
template <int K>
HRT_DEV v3p quat_axis_p(const q4p q) {
    const f2 two = dup2(2.f), one = dup2(1.f), m2 = dup2(-2.f);
    if (K == 0)
        return make_v3p(fma2(m2, fma2(q.z, q.z, mul2(q.y, q.y)), one), mul2(two, fma2(q.w, q.z, mul2(q.x, q.y))),
                        mul2(two, fnma2(q.w, q.y, mul2(q.x, q.z))));
    if (K == 1)
        return make_v3p(mul2(two, fnma2(q.w, q.z, mul2(q.x, q.y))), fma2(m2, fma2(q.z, q.z, mul2(q.x, q.x)), one),
                        mul2(two, fma2(q.w, q.x, mul2(q.y, q.z))));
    return make_v3p(mul2(two, fma2(q.w, q.y, mul2(q.x, q.z))), mul2(two, fnma2(q.w, q.x, mul2(q.y, q.z))),
                    fma2(m2, fma2(q.y, q.y, mul2(q.x, q.x)), one));
}

// base + R(q) v for a per-arm constant vector v (broadcast scalars): v + 2 (w t + u x t), t = u x v
HRT_DEV v3p quat_rotate_add_p(const q4p q, const float* v, const v3p base) {
    const f2 vx = dup2(v[0]), vy = dup2(v[1]), vz = dup2(v[2]);
    const f2 tx = fms2(q.y, vz, mul2(q.z, vy)), ty = fms2(q.z, vx, mul2(q.x, vz)), tz = fms2(q.x, vy, mul2(q.y, vx));
    const f2 sx = fma2(q.w, tx, fms2(q.y, tz, mul2(q.z, ty)));
    const f2 sy = fma2(q.w, ty, fms2(q.z, tx, mul2(q.x, tz)));
    const f2 sz = fma2(q.w, tz, fms2(q.x, ty, mul2(q.y, tx)));
    const f2 two = dup2(2.f);
    return make_v3p(fma2(two, sx, add2(base.x, vx)), fma2(two, sy, add2(base.y, vy)), fma2(two, sz, add2(base.z, vz)));
}
// the same for a packed vector
HRT_DEV v3p quat_rotate_p(const q4p q, const v3p v) {
    const v3p u = make_v3p(q.x, q.y, q.z);
    const v3p t = cross3p(u, v);
    const v3p c = cross3p(u, t);
    const f2 two = dup2(2.f);
    return make_v3p(fma2(two, fma2(q.w, t.x, c.x), v.x), fma2(two, fma2(q.w, t.y, c.y), v.y), fma2(two, fma2(q.w, t.z, c.z), v.z));
}

// sin / cos of a half joint angle, both lanes (sincos_half_nf): polynomial part packed, the fold is per lane
HRT_DEV void sincos_half_p(const f2 x, f2* s, f2* c) {
    const f2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const bool fx = ax.x > 0.78539816f, fy = ax.y > 0.78539816f;
    const f2 yf = add2(sub2(dup2(1.5707963705062866f), ax), dup2(-4.371139e-8f));
    const f2 y = make_float2(fx ? yf.x : ax.x, fy ? yf.y : ax.y);
    const f2 z = mul2(y, y);
    const f2 sp = fma2(mul2(fma2(fma2(dup2(-1.9515295891e-4f), z, dup2(8.3321608736e-3f)), z, dup2(-1.6666654611e-1f)), z), y, y);
    const f2 cp = fma2(mul2(fma2(fma2(dup2(2.443315711809948e-5f), z, dup2(-1.388731625493765e-3f)), z, dup2(4.166664568298827e-2f)), z), z,
                       fma2(dup2(-0.5f), z, dup2(1.f)));
    *s = make_float2(copysignf(fx ? cp.x : sp.x, x.x), copysignf(fy ? cp.y : sp.y, x.y));
    *c = make_float2(fx ? sp.x : cp.x, fy ? sp.y : cp.y);
}

// 7x7 SPD solve, packed-lower A, both lanes (chol_solve7)
HRT_DEV void chol_solve7_p(f2* A, f2* b) {
    f2 inv[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            f2 s = A[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fnma2(A[i * (i + 1) / 2 + k], A[j * (j + 1) / 2 + k], s);
            if (i == j) inv[i] = rsqrt2(s);
            else A[i * (i + 1) / 2 + j] = mul2(s, inv[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        f2 s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fnma2(A[i * (i + 1) / 2 + k], b[k], s);
        b[i] = mul2(s, inv[i]);
    }
#pragma unroll
    for (int i = 6; i >= 0; --i) {
        f2 s = b[i];
#pragma unroll
        for (int k = i + 1; k < 7; ++k) s = fnma2(A[k * (k + 1) / 2 + i], b[k], s);
        b[i] = mul2(s, inv[i]);
    }
}

}  // namespace hrt
