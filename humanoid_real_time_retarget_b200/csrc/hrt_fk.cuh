// Batched tree forward kinematics (+ geometric Jacobian) for sm_100a.
//
// Replaces the numeric bodies of
//   cal_forward_kinematics              robot_kinematics_model/kinematics.py:13-39
//   HuForwardModel.forward_kinematics   robot_kinematics_model/hu_forward_model.py:17-33
//
// Mapping: one THREAD per kinematic configuration (frame); a warp owns 32 consecutive frames.
// The tree walk is serial per frame (a child needs its parent), but identical for every frame,
// so all control flow is warp-uniform.  HBM traffic is what bounds the kernel (156 B in, 924 B
// out per Hu config, ~2.4 flop/B), so the design is about moving bytes:
//   * outputs are AoS (L,J,4)/(L,J,3) as the reference returns them; a thread's own rows are
//     528 B / 396 B apart, so results go through a WARP-PRIVATE shared-memory tile of 8 joints
//     x 32 frames and leave as 16-byte vector stores that cover whole 128-byte rows
//     (no __syncthreads anywhere: only __syncwarp);
//   * the parent transform stays in registers while the tree is a chain (parent == j-1); at a
//     branch point it is parked in a warp-private smem slot (liveness-allocated on the host);
//   * tree topology / offsets / axes / limits arrive as a __grid_constant__ parameter block
//     (constant bank, broadcast reads).
#pragma once
#include "hrt_math.cuh"
#include "hrt_params.h"

namespace hrt {

constexpr int FK_CHUNK = 8;                       // joints per staged chunk
constexpr int FK_QROW = FK_CHUNK * 4 + 4;         // 36 words: conflict-free STS.128 per thread row
constexpr int FK_PROW = FK_CHUNK * 3 + 1;         // 25 words (odd stride)
constexpr int FK_AROW = FK_CHUNK + 1;             // 9 words
constexpr int FK_TILE_WORDS = 32 * FK_QROW;       // 1152 words, reused for quats then positions
constexpr int FK_SLOT_WORDS = HRT_MAX_SLOTS * 7 * 32;
constexpr int FK_ANG_WORDS = 32 * FK_AROW;
constexpr int FK_WARP_WORDS = FK_TILE_WORDS + FK_SLOT_WORDS + FK_ANG_WORDS;
constexpr int FK_WARPS_PER_CTA = 4;

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

// flush a staged chunk of quats: tile rows (32 frames) x nj float4 -> global (B,J,4)
HRT_DEV void fk_flush_quats(const float* tile, float* __restrict__ gq, long long f0, long long B,
                            int J, int j0, int nj, int lane) {
    if (nj == FK_CHUNK) {
#pragma unroll
        for (int it = 0; it < FK_CHUNK; ++it) {
            int idx = it * 32 + lane;
            int row = idx >> 3, col = idx & 7;
            long long f = f0 + row;
            if (f < B) {
                float4 v = *reinterpret_cast<const float4*>(tile + row * FK_QROW + col * 4);
                __stcs(reinterpret_cast<float4*>(gq + (f * J + j0 + col) * 4), v);
            }
        }
    } else {
        for (int idx = lane; idx < 32 * nj; idx += 32) {
            int row = idx / nj, col = idx - row * nj;
            long long f = f0 + row;
            if (f < B) {
                float4 v = *reinterpret_cast<const float4*>(tile + row * FK_QROW + col * 4);
                __stcs(reinterpret_cast<float4*>(gq + (f * J + j0 + col) * 4), v);
            }
        }
    }
}

// flush a staged chunk of positions: tile rows x (nj*3) floats -> global (B,J,3)
HRT_DEV void fk_flush_pos(const float* tile, float* __restrict__ gt, long long f0, long long B,
                          int J, int j0, int nj, int lane) {
    const int w = nj * 3;
    if (nj == FK_CHUNK) {
#pragma unroll 4
        for (int it = 0; it < 24; ++it) {
            int idx = it * 32 + lane;
            int row = idx / 24, col = idx - row * 24;
            long long f = f0 + row;
            if (f < B) __stcs(gt + (f * J + j0) * 3 + col, tile[row * FK_PROW + col]);
        }
    } else {
        for (int idx = lane; idx < 32 * w; idx += 32) {
            int row = idx / w, col = idx - row * w;
            long long f = f0 + row;
            if (f < B) __stcs(gt + (f * J + j0) * 3 + col, tile[row * FK_PROW + col]);
        }
    }
}

template <bool FROM_ANGLES, bool EXACT>
__global__ void __launch_bounds__(FK_WARPS_PER_CTA * 32)
fk_kernel(const __grid_constant__ TreeParams tp, const FkArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* tile = smem + warp * FK_WARP_WORDS;
    float* slots = tile + FK_TILE_WORDS;
    float* angt = slots + FK_SLOT_WORDS;
    const int J = tp.J;
    const int D = J - 1;
    const long long n_groups = (a.B + 31) / 32;

    for (long long grp = (long long)blockIdx.x * FK_WARPS_PER_CTA + warp; grp < n_groups;
         grp += (long long)gridDim.x * FK_WARPS_PER_CTA) {
        const long long f0 = grp * 32;
        const long long f = f0 + lane;
        const bool valid = f < a.B;
        const long long fc = valid ? f : a.B - 1;     // clamp so that tail lanes read legal memory

        // ---- root (joint 0): G_r[0] = l[0] as given (NOT normalised), G_t[0] = root translation
        float4 gq;
        vec3 gp;
        if (FROM_ANGLES) {
            gq = a.root_q ? __ldg(reinterpret_cast<const float4*>(a.root_q) + fc) : make_float4(0.f, 0.f, 0.f, 1.f);
        } else {
            gq = __ldg(reinterpret_cast<const float4*>(a.local_q) + fc * J);
        }
        if (a.root_t) gp = make_vec3(__ldg(a.root_t + fc * 3), __ldg(a.root_t + fc * 3 + 1), __ldg(a.root_t + fc * 3 + 2));
        else gp = make_vec3(0.f, 0.f, 0.f);
        if (valid) {
            if (a.out_gq) __stcs(reinterpret_cast<float4*>(a.out_gq) + f * J, gq);
            if (a.out_gt) { __stcs(a.out_gt + f * J * 3, gp.x); __stcs(a.out_gt + f * J * 3 + 1, gp.y); __stcs(a.out_gt + f * J * 3 + 2, gp.z); }
        }
        if (tp.save_slot[0] >= 0) {
            float* s = slots + tp.save_slot[0] * 7 * 32 + lane;
            s[0] = gq.x; s[32] = gq.y; s[64] = gq.z; s[96] = gq.w; s[128] = gp.x; s[160] = gp.y; s[192] = gp.z;
        }

        // ---- joints 1..J-1 in chunks of 8
        for (int j0 = 1; j0 < J; j0 += FK_CHUNK) {
            const int nj = min(FK_CHUNK, J - j0);
            float4 lq[FK_CHUNK];
            if (FROM_ANGLES) {
                // coalesced load of this chunk's angles: rows = frames, cols = dofs j0-1 .. j0-1+nj
                for (int idx = lane; idx < 32 * nj; idx += 32) {
                    int row = idx / nj, col = idx - row * nj;
                    long long fr = min(f0 + row, a.B - 1);
                    angt[row * FK_AROW + col] = __ldcs(a.angles + fr * D + (j0 - 1) + col);
                }
            } else {
                // coalesced float4 load of this chunk's local quats into the tile
                for (int idx = lane; idx < 32 * nj; idx += 32) {
                    int row = idx / nj, col = idx - row * nj;
                    long long fr = min(f0 + row, a.B - 1);
                    float4 v = __ldcs(reinterpret_cast<const float4*>(a.local_q + (fr * J + j0 + col) * 4));
                    *reinterpret_cast<float4*>(tile + row * FK_QROW + col * 4) = v;
                }
            }
            __syncwarp();
            if (!FROM_ANGLES) {
#pragma unroll
                for (int jj = 0; jj < FK_CHUNK; ++jj)
                    if (jj < nj) lq[jj] = *reinterpret_cast<const float4*>(tile + lane * FK_QROW + jj * 4);
                __syncwarp();
            }

            vec3 pos[FK_CHUNK];
#pragma unroll
            for (int jj = 0; jj < FK_CHUNK; ++jj) {
                if (jj < nj) {
                    const int j = j0 + jj;
                    float4 pq = gq;
                    vec3 pp = gp;
                    const int src = tp.src_slot[j];
                    if (src >= 0) {
                        const float* s = slots + src * 7 * 32 + lane;
                        pq = make_float4(s[0], s[32], s[64], s[96]);
                        pp = make_vec3(s[128], s[160], s[192]);
                    }
                    const vec3 off = make_vec3(tp.off[j * 3], tp.off[j * 3 + 1], tp.off[j * 3 + 2]);
                    if (FROM_ANGLES) {
                        float ang = angt[lane * FK_AROW + jj];
                        if (a.clip) {
                            // forward value of the straight-through clamp: (clamp(x) - x) + x
                            float c = fminf(fmaxf(ang, tp.lower[j]), tp.upper[j]);
                            ang = add_rn(sub_rn(c, ang), ang);
                        }
                        const int k = tp.axis[j];
                        if (EXACT) {
                            gq = quat_mul_norm_x(pq, quat_from_angle_axis_k_x(ang, k));
                        } else {
                            float s, c;
                            sincosf(0.5f * ang, &s, &c);
                            if (c < 0.f) { s = -s; c = -c; }              // quat_normalize's sign flip
                            float inv = rsqrtf(s * s + c * c);
                            gq = quat_normalize_f(quat_mul_axis_f(pq, k, s * inv, c * inv));
                        }
                    } else {
                        gq = EXACT ? quat_mul_norm_x(pq, lq[jj]) : quat_mul_norm_f(pq, lq[jj]);
                    }
                    vec3 r = EXACT ? quat_rotate_x(pq, off) : quat_rotate_f(pq, off);
                    gp = EXACT ? make_vec3(add_rn(r.x, pp.x), add_rn(r.y, pp.y), add_rn(r.z, pp.z)) : add3(r, pp);
                    pos[jj] = gp;
                    *reinterpret_cast<float4*>(tile + lane * FK_QROW + jj * 4) = gq;
                    const int sv = tp.save_slot[j];
                    if (sv >= 0) {
                        float* s = slots + sv * 7 * 32 + lane;
                        s[0] = gq.x; s[32] = gq.y; s[64] = gq.z; s[96] = gq.w; s[128] = gp.x; s[160] = gp.y; s[192] = gp.z;
                    }
                }
            }
            __syncwarp();
            if (a.out_gq) fk_flush_quats(tile, a.out_gq, f0, a.B, J, j0, nj, lane);
            __syncwarp();
            if (a.out_gt) {
#pragma unroll
                for (int jj = 0; jj < FK_CHUNK; ++jj)
                    if (jj < nj) {
                        tile[lane * FK_PROW + jj * 3] = pos[jj].x;
                        tile[lane * FK_PROW + jj * 3 + 1] = pos[jj].y;
                        tile[lane * FK_PROW + jj * 3 + 2] = pos[jj].z;
                    }
                __syncwarp();
                fk_flush_pos(tile, a.out_gt, f0, a.B, J, j0, nj, lane);
                __syncwarp();
            }
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
        const long long fc = min(f0 + lane, a.B - 1);
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
                    float ang = __ldg(a.angles + fc * D + (j - 1));
                    if (a.clip) {
                        float cl = fminf(fmaxf(ang, tp.lower[j]), tp.upper[j]);
                        ang = add_rn(sub_rn(cl, ang), ang);
                    }
                    const int kx = tp.axis[j];
                    const vec3 e = make_vec3(kx == 0 ? 1.f : 0.f, kx == 1 ? 1.f : 0.f, kx == 2 ? 1.f : 0.f);
                    const vec3 off = make_vec3(tp.off[j * 3], tp.off[j * 3 + 1], tp.off[j * 3 + 2]);
                    ax[c] = quat_rotate_f(gq, e);
                    gp = add3(quat_rotate_f(gq, off), gp);
                    pj[c] = gp;
                    float s, cs;
                    sincosf(0.5f * ang, &s, &cs);
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
            // flush: 32 contiguous blocks of 6*D floats at (f, k)
            for (int r = 0; r < 32; ++r) {
                const long long f = f0 + r;
                if (f >= a.B) break;
                float* dst = a.out_jac + (f * jp.K + k) * blk;
                for (int i = lane; i < blk; i += 32) __stcs(dst + i, tile[r * row_words + i]);
            }
            __syncwarp();
        }
    }
}

}  // namespace hrt
