// Clip-level stages around the per-frame solvers (sm_100a):
//   Retarget.rescale_motion_to_standard_size          retarget/main.py:36-47
//   RetargetHuV5fromMocap._rebuild_with_vtrdyn_zero_pose (global-rotation rebuild)   retarget/main.py:116-152
//   SkeletonMotion._compute_velocity / _compute_angular_velocity                      poselib/poselib/skeleton/skeleton3d.py:1126-1146
//     (np.gradient + scipy.ndimage.gaussian_filter1d(sigma=2, mode="nearest"))
// Frames lead; a warp owns 32 consecutive frames of a clip, staged as one contiguous span.
#pragma once
#include "hrt_fk_limb.cuh"
#include "hrt_math.cuh"
#include "hrt_ops.cuh"
#include "hrt_pos.cuh"

namespace hrt {

constexpr int MOT_WARPS = 4;

struct RescaleParams {
    int J;
    int8_t parent[HRT_MAX_JOINTS];
    float off[HRT_MAX_JOINTS][3];      // zero-pose local translations
    float dir[3];                      // coord_transform(dir=...) applied first (main.py:170), 1,1,1 = none
};

// p'[j] = p'[parent] + (p[j] - p[parent]) / (||p[j] - p[parent]|| / ||off[j]||), joints in index order
__global__ void __launch_bounds__(MOT_WARPS * 32)
rescale_motion_kernel(const __grid_constant__ RescaleParams rp, const float* __restrict__ gt, long long B,
                      float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = rp.J * 3;
    float* src = smem + (size_t)warp * 2 * 32 * W;
    float* dst = src + 32 * W;
    const long long n_tiles = (B + 31) / 32;
    // zero-pose bone lengths: the same value for every frame, computed once per CTA (it was a square root per joint and frame)
    __shared__ float off_norm[HRT_MAX_JOINTS];
    for (int j = threadIdx.x; j < rp.J; j += blockDim.x) off_norm[j] = norm3_x(make_vec3(rp.off[j][0], rp.off[j][1], rp.off[j][2]));
    __syncthreads();
    for (long long t = (long long)blockIdx.x * MOT_WARPS + warp; t < n_tiles; t += (long long)gridDim.x * MOT_WARPS) {
        const long long f0 = t * 32;
        const int cnt = (int)min(32LL, B - f0);
        for (int w = lane; w < cnt * W; w += 32) src[w] = __ldcs(gt + f0 * W + w);
        __syncwarp();
        if (lane < cnt) {
            float* s = src + lane * W;
            float* d = dst + lane * W;
            for (int j = 0; j < rp.J; ++j) {
                s[j * 3] = mul_rn(s[j * 3], rp.dir[0]); s[j * 3 + 1] = mul_rn(s[j * 3 + 1], rp.dir[1]); s[j * 3 + 2] = mul_rn(s[j * 3 + 2], rp.dir[2]);
            }
            for (int j = 0; j < rp.J; ++j) {
                const int p = rp.parent[j];
                if (p < 0) { d[j * 3] = s[j * 3]; d[j * 3 + 1] = s[j * 3 + 1]; d[j * 3 + 2] = s[j * 3 + 2]; continue; }
                const vec3 v = sub3_x(ld3(s + j * 3), ld3(s + p * 3));
                const float scale = div_rn(norm3_x(v), off_norm[j]);
                d[j * 3] = add_rn(d[p * 3], div_rn(v.x, scale));
                d[j * 3 + 1] = add_rn(d[p * 3 + 1], div_rn(v.y, scale));
                d[j * 3 + 2] = add_rn(d[p * 3 + 2], div_rn(v.z, scale));
            }
        }
        __syncwarp();
        for (int w = lane; w < cnt * W; w += 32) __stcs(out + f0 * W + w, dst[w]);
        __syncwarp();
    }
}

struct RebuildParams {
    int J;
    int8_t parent[HRT_MAX_JOINTS];
    float off[HRT_MAX_JOINTS][3];
    int n_kabsch;                      // joints whose rotation comes from a 3-point Kabsch fit
    int kabsch_joint[2];               // main.py:126-136: joints 0 and 10
    int kabsch_pts[2][3];              // [4,1,7] and [17,13,11]
    uint8_t skip[HRT_MAX_JOINTS];      // main.py:146: joint 0, children of joint 0 / joint 10
};

// per-bone max over the clip of ||p[j] - p[parent]|| (quat_between_two_vecs' whole-batch early-out)
// max over the clip of every bone's length (the early-out test of main.py's rotation rebuild), ONE pass over the clip:
// a lane takes a frame, the warp reduces each bone's length with fmaxf, one shared-memory maximum per CTA and bone, one
// global atomic per CTA and bone.  (It was one pass PER BONE, each reading two 12-byte pieces of every 252-byte row:
// the clip 20 times over in 32-byte sectors.)
__global__ void __launch_bounds__(256)
bone_max_norm_kernel(const __grid_constant__ RebuildParams rp, const float* __restrict__ gt, long long B,
                     unsigned* __restrict__ out_bits) {
    __shared__ unsigned smax[HRT_MAX_JOINTS];
    for (int j = threadIdx.x; j < rp.J; j += blockDim.x) smax[j] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long f0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); f0 < B; f0 += stride) {
        const long long f = f0 + lane;
        const bool valid = f < B;
        const float* r = gt + (valid ? f : B - 1) * rp.J * 3;
        for (int j = 0; j < rp.J; ++j) {
            const int p = rp.parent[j];
            if (p < 0) continue;
            float m = valid ? norm3_x(sub3_x(ld3(r + j * 3), ld3(r + p * 3))) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) atomicMax(smax + j, __float_as_uint(m));
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < rp.J; j += blockDim.x)
        if (rp.parent[j] >= 0) atomicMax(out_bits + j, smax[j]);
}

__global__ void __launch_bounds__(MOT_WARPS * 32)
rebuild_rotation_kernel(const __grid_constant__ RebuildParams rp, const float* __restrict__ gt, long long B,
                        const unsigned* __restrict__ bone_max_bits, float* __restrict__ out_gq) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = rp.J * 3, Q = rp.J * 4;
    float* src = smem + (size_t)warp * 32 * (W + Q);
    float* dst = src + 32 * W;
    const long long n_tiles = (B + 31) / 32;
    for (long long t = (long long)blockIdx.x * MOT_WARPS + warp; t < n_tiles; t += (long long)gridDim.x * MOT_WARPS) {
        const long long f0 = t * 32;
        const int cnt = (int)min(32LL, B - f0);
        for (int w = lane; w < cnt * W; w += 32) src[w] = __ldcs(gt + f0 * W + w);
        for (int w = lane; w < cnt * rp.J; w += 32) *reinterpret_cast<float4*>(dst + w * 4) = make_float4(0.f, 0.f, 0.f, 1.f);
        __syncwarp();
        if (lane < cnt) {
            const float* s = src + lane * W;
            float* d = dst + lane * Q;
            for (int k = 0; k < rp.n_kabsch; ++k) {
                vec3 M[3], Z[3];
                const int jk = rp.kabsch_joint[k];
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    const int pj = rp.kabsch_pts[k][n];
                    M[n] = sub3_x(ld3(s + pj * 3), ld3(s + jk * 3));
                    Z[n] = make_vec3(rp.off[pj][0], rp.off[pj][1], rp.off[pj][2]);
                }
                *reinterpret_cast<float4*>(d + jk * 4) = kabsch_quat<3>(M, Z);
            }
            for (int j = 0; j < rp.J; ++j) {
                if (rp.skip[j]) continue;
                const int p = rp.parent[j];
                vec3 a = make_vec3(rp.off[j][0], rp.off[j][1], rp.off[j][2]);
                vec3 b = sub3_x(ld3(s + j * 3), ld3(s + p * 3));
                float4 q = make_float4(0.f, 0.f, 0.f, 1.f);
                const bool early = !(norm3_x(a) > 1e-6f) || !(__uint_as_float(bone_max_bits[j]) > 1e-6f);
                if (!early) {
                    a = div3_x(a, norm3_x(a));
                    b = div3_x(b, norm3_x(b));
                    const vec3 c = cross3_x(a, b);
                    q = quat_normalize_x(make_float4(c.x, c.y, c.z, add_rn(1.f, dot3_x(a, b))));
                }
                *reinterpret_cast<float4*>(d + p * 4) = q;       // written at the PARENT index (main.py:149)
            }
        }
        __syncwarp();
        for (int w = lane; w < cnt * Q; w += 32) __stcs(out_gq + f0 * Q + w, dst[w]);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// velocities along the frame axis
// ---------------------------------------------------------------------------------------------
// np.gradient(p, axis=frames) / dt in fp32: central differences inside, one-sided at both ends
// (frame, component) of a grid-stride element index advanced by additions: a 64-bit division per element was most of
// these kernels' instructions
struct FrameIdx {
    long long t, dt;
    int c, dc, C;
    __device__ FrameIdx(long long i, long long stride, long long C_) : t(i / C_), dt(stride / C_), c((int)(i % C_)), dc((int)(stride % C_)), C((int)C_) {}
    __device__ void next() {
        t += dt; c += dc;
        if (c >= C) { c -= C; ++t; }
    }
};

__global__ void __launch_bounds__(256)
frame_gradient_kernel(const float* __restrict__ p, long long T, long long C, float dt, float* __restrict__ out) {
    const long long n = T * C, stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (FrameIdx f(i, stride, C); i < n; i += stride, f.next()) {
        const long long t = f.t;
        float g;
        if (t == 0) g = div_rn(sub_rn(__ldg(p + i + C), __ldg(p + i)), 1.f);
        else if (t == T - 1) g = div_rn(sub_rn(__ldg(p + i), __ldg(p + i - C)), 1.f);
        else g = div_rn(sub_rn(__ldg(p + i + C), __ldg(p + i - C)), 2.f);
        out[i] = div_rn(g, dt);
    }
}

__global__ void __launch_bounds__(256)
angular_velocity_raw_kernel(const float4* __restrict__ r, long long T, long long J, float dt, float* __restrict__ out) {
    const long long n = T * J, stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (FrameIdx f(i, stride, J); i < n; i += stride, f.next()) {
        const long long t = f.t;
        float4 q = make_float4(0.f, 0.f, 0.f, 1.f);
        if (t < T - 1) q = quat_mul_norm_x(__ldg(r + i + J), quat_conj(__ldg(r + i)));
        const float s = sub_rn(mul_rn(2.f, mul_rn(q.w, q.w)), 1.f);
        const float angle = acosf(fminf(fmaxf(s, -1.f), 1.f));
        const vec3 v = make_vec3(q.x, q.y, q.z);
        const vec3 ax = div3_x(v, fmaxf(norm3_x(v), 1e-9f));
        out[i * 3] = div_rn(mul_rn(ax.x, angle), dt);
        out[i * 3 + 1] = div_rn(mul_rn(ax.y, angle), dt);
        out[i * 3 + 2] = div_rn(mul_rn(ax.z, angle), dt);
    }
}

constexpr int HRT_GAUSS_MAX_RADIUS = 128;
struct GaussParams {
    int radius;                 // int(4 * sigma + 0.5): 8 for sigma = 2 (velocities), 80 for sigma = 20 (forward vector)
    double w[2 * HRT_GAUSS_MAX_RADIUS + 1];   // w[k] = weight of offset k - radius (normalised, fp64, as scipy builds it)
};

// scipy.ndimage.gaussian_filter1d along frames, mode="nearest": fp64 accumulation in correlate1d's
// symmetric order (centre, then the +-j pairs from the outside in), result rounded to the output type
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
gauss_filter_frames_kernel(const __grid_constant__ GaussParams gp, const TI* __restrict__ x, long long T, long long C,
                           TO* __restrict__ out) {
    const long long n = T * C, stride = (long long)gridDim.x * blockDim.x;
    const int R = gp.radius;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (FrameIdx f(i, stride, C); i < n; i += stride, f.next()) {
        const long long t = f.t;
        const TI* xi = x + i;
        double acc = (double)__ldg(xi) * gp.w[R];
        if (t >= R && t + R < T) {
            // interior frames (all but 2 R of the clip): the taps are at fixed 32-bit offsets from the element
            const int Ci = (int)C;
            for (int j = -R; j < 0; ++j) acc += ((double)__ldg(xi + j * Ci) + (double)__ldg(xi - j * Ci)) * gp.w[j + R];
        } else {
            const long long c = f.c;
            for (int j = -R; j < 0; ++j) {
                const long long ta = min(max(t + j, 0LL), T - 1), tb = min(max(t - j, 0LL), T - 1);
                acc += ((double)__ldg(x + ta * C + c) + (double)__ldg(x + tb * C + c)) * gp.w[j + R];
            }
        }
        out[i] = (TO)acc;
    }
}

// The same filter for the velocities' radius (8) as a sliding window: a thread owns one component over a run of
// consecutive frames and keeps the 17 taps as doubles in registers, so every input element is loaded and widened ONCE
// per run instead of 17 times (the fp32 -> fp64 conversions, 34 per output on the quarter-rate conversion pipe, were the
// generic kernel's bound).  Same accumulation order, same bits.  Consecutive threads = consecutive components of a run
// (coalesced rows); a run re-reads 2 R frames of its neighbours (L2 hits).
constexpr int GAUSS_RUN = 128;
template <typename TI, typename TO, int R>
__global__ void __launch_bounds__(256)
gauss_filter_frames_window_kernel(const __grid_constant__ GaussParams gp, const TI* __restrict__ x, long long T, long long C,
                                  TO* __restrict__ out) {
    const long long n_runs = (T + GAUSS_RUN - 1) / GAUSS_RUN;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_runs * C) return;
    const long long k = idx / C, c = idx - k * C;
    const long long t0 = k * GAUSS_RUN, t1 = min(T, t0 + GAUSS_RUN);
    double gw[R + 1];
#pragma unroll
    for (int j = 0; j <= R; ++j) gw[j] = gp.w[j];
    double win[2 * R + 1];
#pragma unroll
    for (int m = 0; m <= 2 * R; ++m) win[m] = (double)__ldg(x + min(max(t0 - R + m, 0LL), T - 1) * C + c);
    const TI* nxt = x + min(t0 + R + 1, T - 1) * C + c;              // the element that enters the window next
    for (long long t = t0; t < t1; ++t) {
        double acc = win[R] * gw[R];
#pragma unroll
        for (int j = -R; j < 0; ++j) acc += (win[R + j] + win[R - j]) * gw[j + R];
        out[t * C + c] = (TO)acc;
        const double in = (double)__ldg(nxt);
        if (t + R + 2 < T) nxt += C;                                  // mode "nearest": the last frame repeats
#pragma unroll
        for (int m = 0; m < 2 * R; ++m) win[m] = win[m + 1];
        win[2 * R] = in;
    }
}

// SkeletonState.compute_forward_vector, poselib/poselib/skeleton/skeleton3d.py:542-566, first half: side direction
// from shoulders and hips (fp32, numpy's left-to-right order), normalised, crossed with the up vector (0,1,0) in
// fp64 (numpy promotes float32 x int64 to float64).  One thread per frame.
__global__ void __launch_bounds__(256)
forward_raw_kernel(const float* __restrict__ gt, long long T, int J, int ls, int rs, int lh, int rh, double* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
        const float* f = gt + t * J * 3;
        float s[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            s[k] = sub_rn(add_rn(sub_rn(__ldg(f + ls * 3 + k), __ldg(f + rs * 3 + k)), __ldg(f + lh * 3 + k)), __ldg(f + rh * 3 + k));
        const float n = __fsqrt_rn(add_rn(add_rn(mul_rn(s[0], s[0]), mul_rn(s[1], s[1])), mul_rn(s[2], s[2])));
        const double sx = (double)div_rn(s[0], n), sy = (double)div_rn(s[1], n), sz = (double)div_rn(s[2], n);
        out[t * 3 + 0] = __dsub_rn(__dmul_rn(sy, 0.0), __dmul_rn(sz, 1.0));
        out[t * 3 + 1] = __dsub_rn(__dmul_rn(sz, 0.0), __dmul_rn(sx, 0.0));
        out[t * 3 + 2] = __dsub_rn(__dmul_rn(sx, 1.0), __dmul_rn(sy, 0.0));
    }
}

// second half: rows of the smoothed (T,3) array scaled to unit length, fp64, in place
__global__ void __launch_bounds__(256)
normalize_rows3_f64_kernel(double* __restrict__ x, long long T) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
        const double a = x[t * 3], b = x[t * 3 + 1], c = x[t * 3 + 2];
        const double n = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
        x[t * 3] = __ddiv_rn(a, n); x[t * 3 + 1] = __ddiv_rn(b, n); x[t * 3 + 2] = __ddiv_rn(c, n);
    }
}

}  // namespace hrt
