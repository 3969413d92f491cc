// Geometric Jacobian kernel (the FK kernel itself lives in hrt_fk_limb.cuh).
#pragma once
#include "hrt_fk_limb.cuh"

namespace hrt {

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
                for (int i = lane; i < blk; i += 32) dst[i] = tile[r * row_words + i];
            }
            __syncwarp();
        }
    }
}

}  // namespace hrt
