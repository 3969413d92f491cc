// Batched tree forward kinematics (+ geometric Jacobian) for sm_100a.
//
// Replaces the numeric bodies of
//   cal_forward_kinematics              robot_kinematics_model/kinematics.py:13-39
//   HuForwardModel.forward_kinematics   robot_kinematics_model/hu_forward_model.py:17-33
//
// Mapping: one THREAD per kinematic configuration (frame); a warp owns 32 consecutive frames.
// The tree walk is serial per frame (a child needs its parent) but identical for every frame, so
// all control flow is warp-uniform.  HBM traffic is what should bound the kernel (156 B in, 924 B
// out per Hu config, ~2.4 flop/B), so everything else is kept off the issue slots:
//   * outputs are AoS (L,J,4)/(L,J,3) as the reference returns them; a thread's own rows are
//     528 B / 396 B apart, so results go through a WARP-PRIVATE shared-memory tile of 8 joints x
//     32 frames and leave as 16-byte vector stores covering whole 128-byte rows (quats) and
//     96-byte rows (positions).  No __syncthreads anywhere: only __syncwarp;
//   * the store patterns are precomputed per lane once (the (row, col) walk of the position tile
//     repeats every 3 iterations), so a flush iteration is LDS + STG + two integer adds;
//   * each thread reads its own row of joint angles, 32 bytes (= one whole sector) per 8-joint
//     chunk, software-pipelined one chunk ahead (the first chunk of the next task is issued during
//     the last chunk of this one), and the next task's rows are touched with prefetch.global.L2 a
//     whole task ahead, so load latency is never exposed;
//   * the parent transform stays in registers while the tree is a chain (parent == j-1); at a
//     branch point it is parked in a warp-private smem slot (liveness-allocated on the host);
//   * topology / offsets / axes / limits arrive as a __grid_constant__ parameter block, one
//     LDC.128 + one LDC.64 per joint (constant bank, broadcast).
#pragma once
#include "hrt_math.cuh"
#include "hrt_params.h"

namespace hrt {

constexpr int FK_CHUNK = 8;                       // joints per staged chunk
constexpr int FK_QROW = FK_CHUNK * 4 + 4;         // 36 words: conflict-free STS.128 per thread row
constexpr int FK_PROW = FK_CHUNK * 3 + 1;         // 25 words (odd stride)
constexpr int FK_QTILE = 32 * FK_QROW;            // 1152 words
constexpr int FK_PTILE = 32 * FK_PROW;            // 800 words
constexpr int FK_SLOT_WORDS = 7 * 32;             // one parked transform for 32 frames
constexpr int FK_WARPS_PER_CTA = 4;

HRT_HD inline int fk_warp_words(int n_slots) { return FK_QTILE + FK_PTILE + n_slots * FK_SLOT_WORDS; }

// Output stores use the default write-back policy on purpose: a 396-byte / 528-byte output row is
// not sector aligned, so consecutive 8-joint chunks of a row share 32-byte sectors; with an
// evict-first hint those half-written sectors were evicted between chunks and cost a DRAM
// read-modify-write each (ncu: 481 MB read for 164 MB of input).  Left in L2 they merge.
template <typename T>
HRT_DEV void HRT_ST(T* p, const T v) { *p = v; }

HRT_DEV void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
HRT_DEV void cp_async4(float* smem_dst, const float* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
HRT_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
HRT_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
HRT_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

struct FkArgs {
    long long B;
    const float* __restrict__ angles;    // (B, J-1)      [FROM_ANGLES]
    const float* __restrict__ local_q;   // (B, J, 4)     [!FROM_ANGLES]
    const float* __restrict__ root_t;    // (B, 3) or nullptr (= 0)
    const float* __restrict__ root_q;    // (B, 4) [FROM_ANGLES] or nullptr (= identity)
    float* __restrict__ out_gq;          // (B, J, 4) or nullptr
    float* __restrict__ out_gt;          // (B, J, 3) or nullptr
    float* __restrict__ out_jac;         // (B, K, 6, J-1) or nullptr
    int clip;
};

HRT_DEV void slot_store(float* s, const float4 q, const vec3 p) {
    s[0] = q.x; s[32] = q.y; s[64] = q.z; s[96] = q.w; s[128] = p.x; s[160] = p.y; s[192] = p.z;
}

// one chunk (8 dofs starting at d0) of a frame's joint angles; each lane reads its own row, one
// 32-byte sector per chunk, so no byte is fetched twice
HRT_DEV void load_angle_chunk(const float* row, int d0, int D, bool vec, float4& v0, float4& v1) {
    if (vec) {
        v0 = __ldcs(reinterpret_cast<const float4*>(row + d0));
        v1 = __ldcs(reinterpret_cast<const float4*>(row + d0 + 4));
    } else {
        const int m = D - 1;
        v0 = make_float4(__ldg(row + min(d0, m)), __ldg(row + min(d0 + 1, m)), __ldg(row + min(d0 + 2, m)), __ldg(row + min(d0 + 3, m)));
        v1 = make_float4(__ldg(row + min(d0 + 4, m)), __ldg(row + min(d0 + 5, m)), __ldg(row + min(d0 + 6, m)), __ldg(row + min(d0 + 7, m)));
    }
}

template <bool FROM_ANGLES, bool EXACT>
__global__ void __launch_bounds__(FK_WARPS_PER_CTA * 32, 5)
fk_kernel(const __grid_constant__ TreeParams tp, const FkArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* qtile = smem + warp * fk_warp_words(tp.n_slots);
    float* ptile = qtile + FK_QTILE;
    float* slots = ptile + FK_PTILE + lane;
    const int J = tp.J;
    const int D = J - 1;
    const long long n_groups = (a.B + 31) / 32;

    // ---- per-lane flush patterns, computed once --------------------------------------------
    // quats: iteration it covers rows 4*it + (lane >> 3), float4 column lane & 7
    const int q_src = (lane >> 3) * FK_QROW + (lane & 7) * 4;
    const int q_dst = (lane >> 3) * J * 4 + (lane & 7) * 4;
    // positions: 24 floats per row; flat index 32*it + lane = 96*m + (32*r + lane), r = it % 3
    int p_src[3], p_dst[3], p_row[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int idx = 32 * r + lane;
        const int row = idx / 24, col = idx - row * 24;
        p_row[r] = row;
        p_src[r] = row * FK_PROW + col;
        p_dst[r] = row * J * 3 + col;
    }
    const bool vec_angles = FROM_ANGLES && ((D & 3) == 0);
    float4 pre0 = make_float4(0.f, 0.f, 0.f, 0.f), pre1 = pre0;   // angles of the upcoming chunk (software pipeline)
    bool have_pre = false;

    for (long long grp = (long long)blockIdx.x * FK_WARPS_PER_CTA + warp; grp < n_groups;
         grp += (long long)gridDim.x * FK_WARPS_PER_CTA) {
        const long long f0 = grp * 32;
        const int rows = (int)min(32LL, a.B - f0);
        const int my = min(lane, rows - 1);                 // tail lanes shadow the last valid frame
        const bool valid = lane < rows;
        float* gq_w = a.out_gq ? a.out_gq + f0 * J * 4 : nullptr;
        float* gt_w = a.out_gt ? a.out_gt + f0 * J * 3 : nullptr;
        const float* ang_row = FROM_ANGLES ? a.angles + (f0 + my) * D : nullptr;
        const float* lq_w = FROM_ANGLES ? nullptr : a.local_q + f0 * J * 4;

        // ---- warm L2 with the next task's rows (a whole task ahead of their use)
        const long long f0n = f0 + (long long)gridDim.x * FK_WARPS_PER_CTA * 32;
        if (FROM_ANGLES) {
            const long long fn = f0n + lane;
            if (fn < a.B) {
                prefetch_l2(a.angles + fn * D);
                if (a.root_q && (lane & 7) == 0) prefetch_l2(a.root_q + fn * 4);
                if (a.root_t && (lane & 7) == 4) prefetch_l2(a.root_t + fn * 3);
            }
            if (!have_pre) load_angle_chunk(ang_row, 0, D, vec_angles && D >= 8, pre0, pre1);
        }
        // ---- root (joint 0): G_r[0] = l[0] as given (NOT normalised), G_t[0] = root translation
        float4 gq;
        vec3 gp;
        if (FROM_ANGLES) gq = a.root_q ? __ldg(reinterpret_cast<const float4*>(a.root_q) + f0 + my) : make_float4(0.f, 0.f, 0.f, 1.f);
        else gq = __ldg(reinterpret_cast<const float4*>(lq_w + my * J * 4));
        if (a.root_t) {
            const float* r = a.root_t + (f0 + my) * 3;
            gp = make_vec3(__ldg(r), __ldg(r + 1), __ldg(r + 2));
        } else {
            gp = make_vec3(0.f, 0.f, 0.f);
        }
        if (valid) {
            if (gq_w) HRT_ST(reinterpret_cast<float4*>(gq_w + lane * J * 4), gq);
            if (gt_w) { float* o = gt_w + lane * J * 3; HRT_ST(o, gp.x); HRT_ST(o + 1, gp.y); HRT_ST(o + 2, gp.z); }
        }
        {
            const int sv = jr_save(tp.jr[0].meta);
            if (sv >= 0) slot_store(slots + sv * FK_SLOT_WORDS, gq, gp);
        }

        // ---- joints 1..J-1 in chunks of 8 ----------------------------------------------------
        for (int j0 = 1; j0 < J; j0 += FK_CHUNK) {
            const int nj = min(FK_CHUNK, J - j0);
            float ang[FK_CHUNK];
            float4 lq[FK_CHUNK];
            if (FROM_ANGLES) {
                ang[0] = pre0.x; ang[1] = pre0.y; ang[2] = pre0.z; ang[3] = pre0.w;
                ang[4] = pre1.x; ang[5] = pre1.y; ang[6] = pre1.z; ang[7] = pre1.w;
                // issue the loads of the next chunk (or of the next task's first chunk) now
                have_pre = false;
                if (j0 + FK_CHUNK < J) {
                    const int d0 = j0 - 1 + FK_CHUNK;
                    load_angle_chunk(ang_row, d0, D, vec_angles && d0 + 8 <= D, pre0, pre1);
                } else if (f0n < a.B) {
                    const long long fr = min(f0n + lane, a.B - 1);
                    load_angle_chunk(a.angles + fr * D, 0, D, vec_angles && D >= 8, pre0, pre1);
                    have_pre = true;
                }
            } else {
                // coalesced float4 load of this chunk's local quats through the quat tile
                for (int it = 0; it < nj; ++it) {
                    const int idx = it * 32 + lane;
                    const int row = idx / nj, col = idx - row * nj;
                    const float4 v = __ldcs(reinterpret_cast<const float4*>(lq_w + (min(row, rows - 1) * J + j0 + col) * 4));
                    *reinterpret_cast<float4*>(qtile + row * FK_QROW + col * 4) = v;
                }
                __syncwarp();
#pragma unroll
                for (int jj = 0; jj < FK_CHUNK; ++jj)
                    lq[jj] = *reinterpret_cast<const float4*>(qtile + lane * FK_QROW + (jj < nj ? jj : 0) * 4);
                __syncwarp();
            }

#pragma unroll
            for (int jj = 0; jj < FK_CHUNK; ++jj) {
                if (jj < nj) {
                    const int j = j0 + jj;
                    const float4 rec = *reinterpret_cast<const float4*>(&tp.jr[j]);      // LDC.128
                    const uint32_t meta = __float_as_uint(rec.w);
                    float4 pq = gq;
                    vec3 pp = gp;
                    const int src = jr_src(meta);
                    if (src >= 0) {
                        const float* s = slots + src * FK_SLOT_WORDS;
                        pq = make_float4(s[0], s[32], s[64], s[96]);
                        pp = make_vec3(s[128], s[160], s[192]);
                    }
                    const vec3 off = make_vec3(rec.x, rec.y, rec.z);
                    if (FROM_ANGLES) {
                        float th = ang[jj];
                        if (a.clip) {
                            // forward value of the straight-through clamp: (clamp(x) - x) + x
                            const float2 lim = *reinterpret_cast<const float2*>(tp.lim[j]);
                            const float c = fminf(fmaxf(th, lim.x), lim.y);
                            th = add_rn(sub_rn(c, th), th);
                        }
                        const int k = jr_axis(meta);
                        if (EXACT) {
                            gq = quat_mul_norm_x(pq, quat_from_angle_axis_k_x(th, k));
                        } else {
                            float s, c;
                            sincos_half_f(0.5f * th, &s, &c);
                            if (c < 0.f) { s = -s; c = -c; }              // quat_normalize's sign flip
                            gq = quat_normalize_f(quat_mul_axis_f(pq, k, s, c));
                        }
                    } else {
                        gq = EXACT ? quat_mul_norm_x(pq, lq[jj]) : quat_mul_norm_f(pq, lq[jj]);
                    }
                    const vec3 r = EXACT ? quat_rotate_x(pq, off) : quat_rotate_f(pq, off);
                    gp = EXACT ? make_vec3(add_rn(r.x, pp.x), add_rn(r.y, pp.y), add_rn(r.z, pp.z)) : add3(r, pp);
                    *reinterpret_cast<float4*>(qtile + lane * FK_QROW + jj * 4) = gq;
                    float* pt = ptile + lane * FK_PROW + jj * 3;
                    pt[0] = gp.x; pt[1] = gp.y; pt[2] = gp.z;
                    const int sv = jr_save(meta);
                    if (sv >= 0) slot_store(slots + sv * FK_SLOT_WORDS, gq, gp);
                }
            }
            __syncwarp();
            // ---- flush the chunk ----------------------------------------------------------------
            if (nj == FK_CHUNK) {
                if (gq_w) {
                    float* dst = gq_w + q_dst + j0 * 4;
                    const int row0 = lane >> 3;
                    float4 v[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) v[it] = *reinterpret_cast<const float4*>(qtile + q_src + it * 4 * FK_QROW);
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        if (row0 + 4 * it < rows) HRT_ST(reinterpret_cast<float4*>(dst + it * 4 * J * 4), v[it]);
                }
                if (gt_w) {
                    float* dst = gt_w + j0 * 3;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float v[12];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int r = 0; r < 3; ++r) v[m * 3 + r] = ptile[p_src[r] + (h * 4 + m) * 4 * FK_PROW];
#pragma unroll
                        for (int m = 0; m < 4; ++m)
#pragma unroll
                            for (int r = 0; r < 3; ++r)
                                if (p_row[r] + 4 * (h * 4 + m) < rows) HRT_ST(dst + p_dst[r] + (h * 4 + m) * 4 * J * 3, v[m * 3 + r]);
                    }
                }
            } else {
                if (gq_w)
                    for (int idx = lane; idx < 32 * nj; idx += 32) {
                        const int row = idx / nj, col = idx - row * nj;
                        if (row < rows)
                            HRT_ST(reinterpret_cast<float4*>(gq_w + (row * J + j0 + col) * 4),
                                   *reinterpret_cast<const float4*>(qtile + row * FK_QROW + col * 4));
                    }
                if (gt_w) {
                    const int w = nj * 3;
                    for (int idx = lane; idx < 32 * w; idx += 32) {
                        const int row = idx / w, col = idx - row * w;
                        if (row < rows) HRT_ST(gt_w + (row * J + j0) * 3 + col, ptile[row * FK_PROW + col]);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Geometric Jacobian (no reference implementation: SURVEY.md F2; spec in DESIGN.md section 5).
// For requested link k and hinge i:  a_i = R_parent(i) e_i,  J_v = a_i x (p_k - p_i),
// J_w = a_i  when i is an ancestor-or-self of k, else 0.   Output (B, K, 6, D).
// One thread per configuration walks only the chain root -> link (<= HRT_MAX_CHAIN joints);
// the 6 x D block of one (frame, link) is contiguous in HBM, staged in a warp-private tile.
// ---------------------------------------------------------------------------------------------
constexpr int JAC_WARPS_PER_CTA = 2;

__global__ void __launch_bounds__(JAC_WARPS_PER_CTA * 32)
jacobian_kernel(const __grid_constant__ TreeParams tp, const __grid_constant__ JacParams jp, const FkArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int D = tp.J - 1;
    const int blk = 6 * D;                 // floats per (frame, link)
    const int row_words = blk + 1;         // odd stride -> conflict-free per-thread rows
    float* tile = smem + warp * 32 * row_words;
    const long long n_groups = (a.B + 31) / 32;
    for (long long grp = (long long)blockIdx.x * JAC_WARPS_PER_CTA + warp; grp < n_groups;
         grp += (long long)gridDim.x * JAC_WARPS_PER_CTA) {
        const long long f0 = grp * 32;
        const int rows = (int)min(32LL, a.B - f0);
        const long long fc = f0 + min(lane, rows - 1);
        const float4 rq = a.root_q ? __ldg(reinterpret_cast<const float4*>(a.root_q) + fc) : make_float4(0.f, 0.f, 0.f, 1.f);
        vec3 rp = make_vec3(0.f, 0.f, 0.f);
        if (a.root_t) rp = make_vec3(__ldg(a.root_t + fc * 3), __ldg(a.root_t + fc * 3 + 1), __ldg(a.root_t + fc * 3 + 2));
        for (int k = 0; k < jp.K; ++k) {
            const int depth = jp.depth[k];
            // zero the tile (columns off the chain stay zero)
            for (int i = lane; i < 32 * row_words; i += 32) tile[i] = 0.f;
            __syncwarp();
            float* row = tile + lane * row_words;
            // pass 1: walk the chain, remember world axis and position of every chain joint
            float4 gq = rq;
            vec3 gp = rp;
            vec3 ax[HRT_MAX_CHAIN], pj[HRT_MAX_CHAIN];
#pragma unroll
            for (int c = 0; c < HRT_MAX_CHAIN; ++c) {
                if (c < depth) {
                    const int j = jp.chain[k][c];
                    const float4 rec = *reinterpret_cast<const float4*>(&tp.jr[j]);
                    float ang = __ldg(a.angles + fc * D + (j - 1));
                    if (a.clip) {
                        const float cl = fminf(fmaxf(ang, tp.lim[j][0]), tp.lim[j][1]);
                        ang = add_rn(sub_rn(cl, ang), ang);
                    }
                    const int kx = jr_axis(__float_as_uint(rec.w));
                    const vec3 e = make_vec3(kx == 0 ? 1.f : 0.f, kx == 1 ? 1.f : 0.f, kx == 2 ? 1.f : 0.f);
                    ax[c] = quat_rotate_f(gq, e);
                    gp = add3(quat_rotate_f(gq, make_vec3(rec.x, rec.y, rec.z)), gp);
                    pj[c] = gp;
                    float s, cs;
                    sincos_half_f(0.5f * ang, &s, &cs);
                    if (cs < 0.f) { s = -s; cs = -cs; }
                    gq = quat_normalize_f(quat_mul_axis_f(gq, kx, s, cs));
                }
            }
            const vec3 pk = gp;
#pragma unroll
            for (int c = 0; c < HRT_MAX_CHAIN; ++c) {
                if (c < depth) {
                    const int col = jp.chain[k][c] - 1;
                    const vec3 jv = cross3_f(ax[c], sub3(pk, pj[c]));
                    row[0 * D + col] = jv.x; row[1 * D + col] = jv.y; row[2 * D + col] = jv.z;
                    row[3 * D + col] = ax[c].x; row[4 * D + col] = ax[c].y; row[5 * D + col] = ax[c].z;
                }
            }
            __syncwarp();
            // flush: `rows` contiguous blocks of 6*D floats at (f, k)
            for (int r = 0; r < rows; ++r) {
                float* dst = a.out_jac + ((f0 + r) * jp.K + k) * blk;
                for (int i = lane; i < blk; i += 32) HRT_ST(dst + i, tile[r * row_words + i]);
            }
            __syncwarp();
        }
    }
}

}  // namespace hrt
