// Limb-parallel batched tree forward kinematics for sm_100a.
//
// Replaces the numeric bodies of
//   cal_forward_kinematics              robot_kinematics_model/kinematics.py:13-39
//   HuForwardModel.forward_kinematics   robot_kinematics_model/hu_forward_model.py:17-33
//
// Why this shape.  FK moves 156 B in / 924 B out per Hu configuration at ~2.4 flop/B: HBM should
// be the bound.  A thread-per-configuration walk of the 32-joint tree is one long dependent chain
// per warp and its AoS outputs (rows 528 B / 396 B apart) need per-row address arithmetic; ncu
// showed that version issue- and latency-bound at 29 % of the HBM roofline (profiles/r01_fk.md).
// Here FOUR lanes cooperate on one configuration, each walking one limb (the host list-schedules
// the tree's chains onto the four lanes: sched[step][lane], 10 steps for Hu instead of 32), so
//   * the dependent chain per warp is 3x shorter,
//   * a warp owns 16 WHOLE configurations, whose outputs are ONE contiguous span of HBM each
//     (16*J*16 B of quats, 16*J*12 B of positions): they are staged in a warp-private shared-memory
//     image of that span and leave with two TMA bulk stores (cp.async.bulk.global.shared::cta,
//     SASS UBLKCP) -- no per-row address math, no partially written sectors,
//   * a child finds its parent's global transform in that same staged image (one LDS.128 + three
//     LDS), so no transform is kept in registers across steps and there is no branch on topology,
//   * inputs (angles, root quats, root translations, or the local quaternions: contiguous spans too) arrive by
//     cp.async (LDGSTS) into a double buffer one task ahead of their use.
// All control flow is warp-uniform; idle lane-steps (18 % for Hu) are predicated off.
#pragma once
#include <type_traits>
#include "hrt_math.cuh"
#include "hrt_params.h"

namespace hrt {

// A lane can walk its limb for CPL configurations at once (c, c + 8, ...): the per-step schedule decode is paid once
// for CPL joint updates and the lane has CPL independent dependency chains in flight.  Measured (profiles/r01_notes.md):
// CPL = 2 lifts the local-quaternion variant from 0.86 to 0.95 of the HBM peak; the joint-angle variant (sin / cos, clamp
// and angle fetch per update, 135 instructions per lane-step of which ~60 are FP) is issue-bound and fastest with CPL = 1
// and twice the resident warps.
constexpr int FKL_GROUP = 32 / HRT_FK_LANES;      // 8 lanes share a limb
#ifndef HRT_FKL_LOCAL_CPL
#define HRT_FKL_LOCAL_CPL 2
#endif
#ifndef HRT_FKL_LOCAL_WARPS
#define HRT_FKL_LOCAL_WARPS 3
#endif
HRT_HD constexpr int fkl_cpl(bool from_angles, bool exact) { return (!from_angles && !exact) ? HRT_FKL_LOCAL_CPL : 1; }
HRT_HD constexpr int fkl_warps(bool from_angles, bool exact) { return (!from_angles && !exact) ? HRT_FKL_LOCAL_WARPS : 4; }

struct FkArgs {
    long long B;
    const float* __restrict__ angles;    // (B, J-1)      [FROM_ANGLES]
    const float* __restrict__ local_q;   // (B, J, 4)     [!FROM_ANGLES]
    const float* __restrict__ root_t;    // (B, 3) or nullptr (= 0)
    const float* __restrict__ root_q;    // (B, 4) [FROM_ANGLES] or nullptr (= identity)
    float* __restrict__ out_gq;          // (B, J, 4) or nullptr
    float* __restrict__ out_gt;          // (B, J, 3) or nullptr
    float* __restrict__ out_jac;         // (B, K, 6, J-1) or nullptr  (jacobian kernel)
    const float4* __restrict__ sched;    // device copy of the schedule: T * 4 entries of 32 bytes
    int T;
    int clip;
};

// shared-memory carve-up, in words (cfgs = configurations per warp task).  Every region is a multiple of 4 words.
// angle rows are staged with a padded stride (D rounded up to a multiple of 4, plus 4): 36 for D = 30 and 32, i.e.
// 4 banks between configurations, so the 8 configurations of a quarter-warp read one joint's angle conflict-free
HRT_HD inline int fkl_angle_stride(int J) { return ((J - 1) + 3) / 4 * 4 + 4; }
HRT_HD inline int fkl_in_words(int J, bool from_angles, int cfgs) {
    return from_angles ? cfgs * fkl_angle_stride(J) + cfgs * 4 + cfgs * 4    // angles | root_q | root_t (padded)
                       : cfgs * J * 4;                                       // local quaternions
}
HRT_HD inline int fkl_warp_words(int J, bool from_angles, int cfgs) {
    return cfgs * J * 4 + cfgs * J * 3 + 2 * fkl_in_words(J, from_angles, cfgs);
}
HRT_HD inline size_t fkl_smem_bytes(int J, int T, bool from_angles, int cfgs, int warps) {
    return (size_t)T * HRT_FK_LANES * 32 + (size_t)warps * fkl_warp_words(J, from_angles, cfgs) * 4;
}

HRT_DEV void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
HRT_DEV void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
HRT_DEV void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
HRT_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
HRT_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// contiguous span global -> shared by the whole warp: 16-byte pieces + a 4-byte tail
HRT_DEV void warp_span_g2s(float* dst, const float* src, int n_words, int lane) {
    const int n4 = n_words >> 2;
    for (int i = lane; i < n4; i += 32) cp_async16(dst + i * 4, src + i * 4);
    for (int i = (n4 << 2) + lane; i < n_words; i += 32) cp_async4(dst + i, src + i);
}

// the same for a span of exactly N4 sixteen-byte pieces: unrolled, one address computation (the rolled loop above costs
// ~8 instructions per piece and lane: 194 of the position kernel's 4,700 instructions per round went there)
template <int N4>
HRT_DEV void warp_span_g2s_n4(float* dst, const float* src, int lane) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst) + lane * 16;
    const char* g = reinterpret_cast<const char*>(src) + lane * 16;
#pragma unroll
    for (int k = 0; k < (N4 + 31) / 32; ++k)
        if (k < N4 / 32 || lane < N4 % 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + k * 512), "l"(g + k * 512) : "memory");
}

HRT_DEV void bulk_store_s2g(float* gmem_dst, const float* smem_src, unsigned bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(s), "r"(bytes) : "memory");
}
// hint: pull a contiguous span (16-byte aligned; whole 16-byte blocks of it) into L2 ahead of the cp.async that will read it
HRT_DEV void l2_prefetch_span(const float* gmem_src, int n_words) {
    const unsigned bytes = ((unsigned)n_words * 4u) & ~15u;
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem_src), "r"(bytes) : "memory");
}
HRT_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
HRT_DEV void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
HRT_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// Explicit shared-window accesses for the walk: it keeps 32-bit shared addresses of its rows and adds the schedule's BYTE
// offsets, one integer add per access (a generic pointer costs a word-index add, a scale and a window-base add each).
// volatile keeps them ordered among themselves and against the warp barriers; the loads carry no memory clobber so the
// arithmetic of a lane's independent chains can be scheduled around them.
HRT_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
HRT_DEV float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
HRT_DEV float lds32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}
HRT_DEV vec3 lds_vec3(unsigned addr) {
    vec3 v;
    asm volatile("ld.shared.f32 %0, [%3];\n\tld.shared.f32 %1, [%3+4];\n\tld.shared.f32 %2, [%3+8];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z) : "r"(addr));
    return v;
}
HRT_DEV void sts128(unsigned addr, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
HRT_DEV void sts_vec3(unsigned addr, const vec3 v) {
    asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+4], %2;\n\tst.shared.f32 [%0+8], %3;\n"
                 ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z) : "memory");
}

template <bool FROM_ANGLES, bool EXACT, bool BOUNDED>
__global__ void __launch_bounds__(fkl_warps(FROM_ANGLES, EXACT) * 32)
fk_limb_kernel(const int J, const FkArgs a) {
    constexpr int FKL_CPL = fkl_cpl(FROM_ANGLES, EXACT);
    constexpr int FKL_CFG = FKL_GROUP * FKL_CPL;           // configurations per warp task
    constexpr int FKL_WARPS_PER_CTA = fkl_warps(FROM_ANGLES, EXACT);
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // a quarter-warp (the unit of a 16-byte shared-memory access) = ONE limb lane over 8 configurations: with an odd
    // joint count the 8 rows of the staged images then fall into 8 different bank groups (ncu: the (cfg, limb) =
    // (lane/4, lane%4) mapping had 8-way conflicts on the quaternion tile, 48 % excess shared wavefronts)
    const int cfg = lane & 7;                   // first configuration of this lane within the warp task
    const int p = lane >> 3;                    // limb lane
    const int D = J - 1;
    const int T = a.T;

    // ---- the schedule goes to shared memory once per CTA -------------------------------------
    float4* sched_s = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < T * HRT_FK_LANES * 2; i += blockDim.x) sched_s[i] = __ldg(a.sched + i);
    __syncthreads();

    const int in_words = fkl_in_words(J, FROM_ANGLES, FKL_CFG);
    float* qtile = smem + T * HRT_FK_LANES * 8 + warp * fkl_warp_words(J, FROM_ANGLES, FKL_CFG);
    float* ptile = qtile + FKL_CFG * J * 4;
    float* inbuf = ptile + FKL_CFG * J * 3;
    const int AS = fkl_angle_stride(J);
    const int ang_words = FKL_CFG * AS;

    const long long n_tasks = (a.B + FKL_CFG - 1) / FKL_CFG;
    const long long stride = (long long)gridDim.x * FKL_WARPS_PER_CTA;
    long long task = (long long)blockIdx.x * FKL_WARPS_PER_CTA + warp;

    // angle rows arrive in pieces of `pw` words (16 / 8 / 4 bytes, what the row alignment allows); a lane's pieces are
    // 32 apart, so their (row, column) advance by constants: no division inside the task loop
    const int pw = (D & 3) == 0 ? 4 : ((D & 1) == 0 ? 2 : 1);
    const int ppr = D > 0 ? D / pw : 1;           // pieces per row
    const int row_step = 32 / ppr, col_step = 32 % ppr;
    const int row0 = lane / ppr, col0 = lane % ppr;

    // issue the asynchronous input copies of `tk` into buffer `b`
    auto stage_inputs = [&](long long tk, int b) {
        if (!FROM_ANGLES && tk < n_tasks) {
            const long long f0 = tk * FKL_CFG;
            const int rows = (int)min((long long)FKL_CFG, a.B - f0);
            warp_span_g2s(inbuf + b * in_words, a.local_q + f0 * J * 4, rows * J * 4, lane);
        }
        if (FROM_ANGLES && tk < n_tasks) {
            const long long f0 = tk * FKL_CFG;
            const int rows = (int)min((long long)FKL_CFG, a.B - f0);
            float* in = inbuf + b * in_words;
            const float* src = a.angles + f0 * D;
            int row = row0, col = col0;
            const int n_pieces = D > 0 ? rows * ppr : 0;
            for (int i = lane; i < n_pieces; i += 32) {
                float* dst = in + row * AS + col * pw;
                if (pw == 4) cp_async16(dst, src + i * 4);
                else if (pw == 2) cp_async8(dst, src + i * 2);
                else cp_async4(dst, src + i);
                row += row_step; col += col_step;
                if (col >= ppr) { col -= ppr; ++row; }
            }
            if (a.root_q) for (int i = lane; i < rows; i += 32) cp_async16(in + ang_words + i * 4, a.root_q + (f0 + i) * 4);
            if (a.root_t) warp_span_g2s(in + ang_words + FKL_CFG * 4, a.root_t + f0 * 3, rows * 3, lane);
        }
        cp_async_commit();
    };

    int buf = 0;
    stage_inputs(task, 0);
    bool pending_store = false;

    for (; task < n_tasks; task += stride, buf ^= 1) {
        const long long f0 = task * FKL_CFG;
        const int rows = (int)min((long long)FKL_CFG, a.B - f0);
        bool cfg_ok[FKL_CPL];
        int c[FKL_CPL];
        unsigned qrow[FKL_CPL], prow[FKL_CPL];          // shared-window byte addresses of this lane's rows
        char *qrow_g[FKL_CPL], *prow_g[FKL_CPL];        // the same as generic pointers (local-quaternion variant, see below)
#pragma unroll
        for (int u = 0; u < FKL_CPL; ++u) {
            cfg_ok[u] = cfg + u * FKL_GROUP < rows;
            c[u] = cfg_ok[u] ? cfg + u * FKL_GROUP : rows - 1;      // tail lanes shadow the last valid configuration
            qrow_g[u] = reinterpret_cast<char*>(qtile + c[u] * J * 4);
            prow_g[u] = reinterpret_cast<char*>(ptile + c[u] * J * 3);
            qrow[u] = smem_addr(qrow_g[u]);
            prow[u] = smem_addr(prow_g[u]);
        }

        // the previous task's bulk stores must have finished READING the tiles before we overwrite them
        if (pending_store) {
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
            pending_store = false;
        }
        stage_inputs(task + stride, buf ^ 1);             // next task's inputs, one task ahead
        cp_async_wait<1>();                               // ... and this task's have landed
        __syncwarp();
        const float* in = inbuf + buf * in_words;
        unsigned arow[FKL_CPL];                         // angle rows (shared-window addresses) / local-quaternion rows
        const char* lrow_g[FKL_CPL];
#pragma unroll
        for (int u = 0; u < FKL_CPL; ++u) {
            arow[u] = smem_addr(in + c[u] * AS);
            lrow_g[u] = reinterpret_cast<const char*>(in + c[u] * J * 4);
        }

        // ---- root (joint 0): G_r[0] = l[0] as given (NOT normalised), G_t[0] = root translation
        for (int r = lane; r < rows; r += 32) {
            float* qr = qtile + r * J * 4;
            float* pr = ptile + r * J * 3;
            if (FROM_ANGLES)
                *reinterpret_cast<float4*>(qr) = a.root_q ? *reinterpret_cast<const float4*>(in + ang_words + r * 4) : make_float4(0.f, 0.f, 0.f, 1.f);
            else
                *reinterpret_cast<float4*>(qr) = *reinterpret_cast<const float4*>(in + r * J * 4);
            if (a.root_t) {
                if (FROM_ANGLES) { const float* rt = in + ang_words + FKL_CFG * 4 + r * 3; pr[0] = rt[0]; pr[1] = rt[1]; pr[2] = rt[2]; }
                else { const float* g = a.root_t + (f0 + r) * 3; pr[0] = __ldg(g); pr[1] = __ldg(g + 1); pr[2] = __ldg(g + 2); }
            } else {
                pr[0] = 0.f; pr[1] = 0.f; pr[2] = 0.f;
            }
        }
        __syncwarp();

        // ---- the scheduled walk: step t, lane p -> joint sched[t][p] ---------------------------
        // The joint-angle variant is issue-bound and takes the explicit shared-window accesses (135 -> 113 instructions per
        // lane-step together with the byte offsets: 0.72 -> 0.83 of the HBM peak); the local-quaternion variant (two chains
        // per lane, 0.90 of the peak) keeps plain generic accesses, which measured the same.
        auto ld_q = [&](int u, unsigned o) {
            return FROM_ANGLES ? lds128(qrow[u] + o) : *reinterpret_cast<const float4*>(qrow_g[u] + o);
        };
        auto ld_p = [&](int u, unsigned o) {
            if (FROM_ANGLES) return lds_vec3(prow[u] + o);
            const float* q = reinterpret_cast<const float*>(prow_g[u] + o);
            return make_vec3(q[0], q[1], q[2]);
        };
        auto st_qp = [&](int u, unsigned qo, unsigned po, const float4 q, const vec3 v) {
            if (FROM_ANGLES) {
                sts128(qrow[u] + qo, q);
                sts_vec3(prow[u] + po, v);
            } else {
                *reinterpret_cast<float4*>(qrow_g[u] + qo) = q;
                float* d = reinterpret_cast<float*>(prow_g[u] + po);
                d[0] = v.x; d[1] = v.y; d[2] = v.z;
            }
        };
        bool wild = false;
        auto walk = [&](auto tame) {
            for (int t = 0; t < T; ++t) {
                const float4 r0 = sched_s[(t * HRT_FK_LANES + p) * 2];
                const float4 r1 = sched_s[(t * HRT_FK_LANES + p) * 2 + 1];
                const float2 lim = make_float2(r1.x, r1.y);
                const uint32_t w3 = __float_as_uint(r0.w), w6 = __float_as_uint(r1.z), w7 = __float_as_uint(r1.w);
                const bool joint_ok = (int)w3 < 0;
                const int k = (int)((w3 >> 16) & 3u);
                const unsigned ang_o = w3 & 0xFFFFu;                   // byte offsets into the angle row / the staged images
                const unsigned qpar_o = w6 & 0xFFFFu, ppar_o = w6 >> 16;
                const unsigned qj_o = w7 & 0xFFFFu, pj_o = w7 >> 16;
                const vec3 off = make_vec3(r0.x, r0.y, r0.z);
                float4 gq[FKL_CPL], pq[FKL_CPL], lq[FKL_CPL];
                vec3 gp[FKL_CPL], pp[FKL_CPL];
                float th[FKL_CPL];
                // all of the step's loads first (the lane's chains are independent), then the arithmetic
    #pragma unroll
                for (int u = 0; u < FKL_CPL; ++u) {
                    pq[u] = ld_q(u, qpar_o);
                    pp[u] = ld_p(u, ppar_o);
                    if (FROM_ANGLES) th[u] = lds32(arow[u] + ang_o);
                    else lq[u] = *reinterpret_cast<const float4*>(lrow_g[u] + qj_o);
                }
    #pragma unroll
                for (int u = 0; u < FKL_CPL; ++u) {
                    if (FROM_ANGLES) {
                        if (a.clip) {
                            // forward value of the straight-through clamp: (clamp(x) - x) + x
                            const float cl = fminf(fmaxf(th[u], lim.x), lim.y);
                            th[u] = add_rn(sub_rn(cl, th[u]), th[u]);
                        }
                        if (EXACT) {
                            gq[u] = quat_mul_norm_x(pq[u], quat_from_angle_axis_k_x(th[u], k));
                        } else {
                            float sn, cs;
                            if (decltype(tame)::value) {
                                wild |= fabsf(th[u]) > 4.7f;
                                sincos_half_nf(0.5f * th[u], &sn, &cs);
                            } else {
                                sincos_half_f(0.5f * th[u], &sn, &cs);
                            }
                            if (cs < 0.f) { sn = -sn; cs = -cs; }                // quat_normalize's sign flip
                            gq[u] = quat_normalize_f(quat_mul_axis_rt_f(pq[u], k, sn, cs));
                        }
                    } else {
                        gq[u] = EXACT ? quat_mul_norm_x(pq[u], lq[u]) : quat_mul_norm_f(pq[u], lq[u]);
                    }
                    if (EXACT) {
                        const vec3 r = quat_rotate_x(pq[u], off);
                        gp[u] = make_vec3(add_rn(r.x, pp[u].x), add_rn(r.y, pp[u].y), add_rn(r.z, pp[u].z));
                    } else {
                        gp[u] = add3(quat_rotate_f(pq[u], off), pp[u]);
                    }
                }
    #pragma unroll
                for (int u = 0; u < FKL_CPL; ++u) {
                    if (joint_ok && cfg_ok[u]) {
                        st_qp(u, qj_o, pj_o, gq[u], gp[u]);
                    }
                }
                __syncwarp();
            }
        };
        if constexpr (FROM_ANGLES && !EXACT && BOUNDED) {
            // Every limit of the tree is finite, so a clipped angle is inside the half-angle polynomials' range -- unless the
            // input is so large (|x| > ~1e7 rad) that (clamp(x) - x) + x loses the clamp to rounding.  The walk runs without
            // the per-update range test and branch and only accumulates "an angle was out of range" in a predicate (one
            // FSETP per update); one warp vote per task sends such a task through the generic walk again.
            walk(std::true_type{});
            if (__any_sync(0xffffffffu, wild)) {
                wild = false;
                walk(std::false_type{});
            }
        } else {
            walk(std::false_type{});
        }

        // ---- results leave as whole contiguous spans ---------------------------------------------
        if (rows == FKL_CFG) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (a.out_gq) bulk_store_s2g(a.out_gq + f0 * J * 4, qtile, (unsigned)(FKL_CFG * J * 16));
                if (a.out_gt) bulk_store_s2g(a.out_gt + f0 * J * 3, ptile, (unsigned)(FKL_CFG * J * 12));
                bulk_commit();
            }
            pending_store = true;
        } else {
            // ragged tail: plain stores
            if (a.out_gq) for (int i = lane; i < rows * J; i += 32)
                *reinterpret_cast<float4*>(a.out_gq + (f0 * J + i) * 4) = *reinterpret_cast<const float4*>(qtile + i * 4);
            if (a.out_gt) for (int i = lane; i < rows * J * 3; i += 32) a.out_gt[f0 * J * 3 + i] = ptile[i];
            __syncwarp();
        }
    }
    cp_async_wait<0>();
    if (pending_store && lane == 0) bulk_wait_read_all();
}

}  // namespace hrt
