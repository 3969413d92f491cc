// Geometric Jacobian kernel (the FK kernel itself lives in hrt_fk_limb.cuh).
#pragma once
#include "hrt_fk_limb.cuh"

namespace hrt {

// ---------------------------------------------------------------------------------------------
// Geometric Jacobian (no reference implementation: SURVEY.md F2; spec in DESIGN.md section 5).
// For requested link k and hinge i:  a_i = R_parent(i) e_i,  J_v = a_i x (p_k - p_i),
// J_w = a_i  when i is an ancestor-or-self of k, else 0.   Output (B, K, 6, D).
// One lane per (configuration, link) walks only the chain root -> link (<= HRT_MAX_CHAIN joints).
// The K 6 x D blocks of consecutive configurations are ONE contiguous span of HBM: the lanes assemble them in a
// shared-memory image and bundles of 8 configurations leave with one TMA bulk store each (per-copy overhead of the
// bulk-store path dominated with one 1.5 KB copy per configuration: 0.61 -> 0.73 of the HBM peak, profiles/r01_notes.md).
// Columns off the chains are zero and never change: the image is zeroed once per kernel, every tile overwrites the
// same chain entries.  A tile's angles and root transform are fetched one tile ahead with all loads in flight.
// ---------------------------------------------------------------------------------------------
constexpr int JAC_WARPS_PER_CTA = 4;
// JAC_BUNDLE consecutive rows (configurations) lie back to back; a 4-word pad follows each bundle (16-byte alignment
// for the bulk copy; shifts the next bundle's column writes to other banks).  Measured 1 / 2 / 4 / 8 rows per copy:
// 0.65 / 0.67 / 0.70 / 0.73 of the HBM peak; one copy per tile (all 32 lanes on one bank) drops to 0.55.
#ifndef HRT_JAC_BUNDLE
#define HRT_JAC_BUNDLE 8
#endif
constexpr int JAC_BUNDLE = HRT_JAC_BUNDLE;
HRT_HD inline int jac_bundle_words(int K, int D) { return JAC_BUNDLE * K * 6 * D + 4; }
HRT_HD inline int jac_tile_words(int K, int D, int cpw) { return cpw / JAC_BUNDLE * jac_bundle_words(K, D); }
HRT_HD inline int jac_lanes_per_cfg(int K) { return K <= 1 ? 1 : (K <= 2 ? 2 : 4); }   // K <= HRT_MAX_LINKS = 4

// lanes of a warp = (configuration, link): the K chains of a configuration are walked in parallel.
// MAXC = unroll bound of the chain loops (8 or HRT_MAX_CHAIN): the walk keeps every chain joint's axis and position in
// registers, so the loops are fully unrolled; the host takes the small instantiation when every requested chain fits
// (Hu wrists: 8 joints) -- half the code (the 16-step body missed the instruction cache: 0.8 warps per issue without an
// instruction) and 96 registers fewer.
template <int MAXC>
__global__ void __launch_bounds__(JAC_WARPS_PER_CTA * 32)
jacobian_kernel(const __grid_constant__ TreeParams tp, const __grid_constant__ JacParams jp, const FkArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int D = tp.J - 1;
    const int blk = 6 * D;                 // floats per (configuration, link)
    const int tile_words = jac_tile_words(jp.K, D, 32 / jac_lanes_per_cfg(jp.K));
    const int lpc = jac_lanes_per_cfg(jp.K);
    const int cpw = 32 / lpc;              // configurations per warp tile
    const int ci = lane / lpc, k = lane % lpc;
    const bool has_link = k < jp.K;
    for (int i = threadIdx.x; i < JAC_WARPS_PER_CTA * tile_words / 4; i += blockDim.x)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float* row = smem + warp * tile_words + (ci / JAC_BUNDLE) * jac_bundle_words(jp.K, D) + (ci % JAC_BUNDLE) * jp.K * blk;
    bool pending = false;
    const int kk = has_link ? k : 0;
    const int depth = jp.depth[kk];
    // this lane's chain, packed four joints to a word (the per-lane index into the constant bank is loop invariant)
    uint32_t chain_w[MAXC / 4];
#pragma unroll
    for (int w = 0; w < MAXC / 4; ++w) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(uint8_t)jp.chain[kk][w * 4 + b] << (8 * b);
        chain_w[w] = v;
    }
    auto chain_joint = [&](int c) { return (int)((chain_w[c >> 2] >> (8 * (c & 3))) & 0xFFu); };
    const long long n_groups = (a.B + cpw - 1) / cpw;
    const long long grp_stride = (long long)gridDim.x * JAC_WARPS_PER_CTA;

    // A tile's inputs (the chain's angles, root transform) are fetched one tile ahead, all loads in flight together:
    // ncu showed the first version latency-bound (issue slots 25 % busy, 3.2 warps per issue waiting on global loads,
    // 8 resident warps per SM) with one dependent load per chain step.
    struct TileIn { float th[MAXC]; float4 rq; vec3 rt; };
    auto fetch = [&](long long grp, TileIn& in) {
        const long long f0 = grp * cpw;
        const int rows = (int)min((long long)cpw, a.B - f0);
        const long long fc = f0 + min(ci, rows - 1);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) in.th[c] = (c < depth) ? __ldg(a.angles + fc * D + (chain_joint(c) - 1)) : 0.f;
        in.rq = a.root_q ? __ldg(reinterpret_cast<const float4*>(a.root_q) + fc) : make_float4(0.f, 0.f, 0.f, 1.f);
        in.rt = make_vec3(0.f, 0.f, 0.f);
        if (a.root_t) in.rt = make_vec3(__ldg(a.root_t + fc * 3), __ldg(a.root_t + fc * 3 + 1), __ldg(a.root_t + fc * 3 + 2));
    };
    TileIn cur, nxt;
    long long grp = (long long)blockIdx.x * JAC_WARPS_PER_CTA + warp;
    if (grp < n_groups) fetch(grp, cur);
    for (; grp < n_groups; grp += grp_stride) {
        const long long f0 = grp * cpw;
        const int rows = (int)min((long long)cpw, a.B - f0);
        if (grp + grp_stride < n_groups) fetch(grp + grp_stride, nxt);
        vec3 gp = cur.rt;
        // walk this lane's chain, remember world axis and position of every chain joint
        float4 gq = cur.rq;
        vec3 ax[MAXC], pj[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            if (c < depth) {
                const int j = chain_joint(c);
                const float4 rec = *reinterpret_cast<const float4*>(&tp.jr[j]);
                float ang = cur.th[c];
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
        // the previous tile's bulk stores (issued by the k == 0 lanes) must have finished reading the rows
        if (pending) { bulk_wait_read_all(); pending = false; }
        __syncwarp();
        if (has_link) {
            float* r = row + k * blk;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                if (c < depth) {
                    const int col = chain_joint(c) - 1;
                    const vec3 jv = cross3_f(ax[c], sub3(pk, pj[c]));
                    r[0 * D + col] = jv.x; r[1 * D + col] = jv.y; r[2 * D + col] = jv.z;
                    r[3 * D + col] = ax[c].x; r[4 * D + col] = ax[c].y; r[5 * D + col] = ax[c].z;
                }
            }
        }
        // one bulk store per configuration: K*6*D contiguous floats at (f, 0)
        fence_proxy_async_smem();
        __syncwarp();
        if (k == 0 && ci % JAC_BUNDLE == 0 && ci < rows) {
            bulk_store_s2g(a.out_jac + (f0 + ci) * jp.K * blk, row, (unsigned)(min(JAC_BUNDLE, rows - ci) * jp.K * blk * 4));
            bulk_commit();
            pending = true;
        }
        cur = nxt;
    }
    if (pending) bulk_wait_read_all();
}

}  // namespace hrt
