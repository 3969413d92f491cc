// Bit-for-bit check of the slow-path-free division / square root of csrc/hrt_math.cuh against the IEEE intrinsics.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -I humanoid_real_time_retarget_b200/csrc \
//        -o /tmp/exact_ops_check tools/ubench/exact_ops_check.cu && /tmp/exact_ops_check
// 2^30 random operand pairs drawn from the ranges the fused kernels divide in (|x| in 2^-60 .. 2^60, random mantissas,
// both signs); sqrt over 2^-100 .. 2^120 (below 2^-101 __fsqrt_rn itself takes its scaled slow path; the kernels take
// square roots of squared norms and of 1 - w^2, i.e. of exact zeros or values above 1e-16).
// Prints the mismatch counts (expected 0) and the worst relative error of drsqrt_n against 1/sqrt in fp64.
#include <cstdio>
#include <cstdint>
#include "hrt_math.cuh"

using namespace hrt;

__device__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ float rnd_float(uint32_t h, int emin, int emax) {
    const uint32_t man = h & 0x7fffffu, sign = (h >> 31) << 31;
    const int e = emin + (int)((h >> 23) & 0xff) % (emax - emin + 1);
    return __uint_as_float(sign | ((uint32_t)(e + 127) << 23) | man);
}

__global__ void check(unsigned long long n, unsigned long long* bad_div, unsigned long long* bad_div3, unsigned long long* bad_sqrt,
                      double* worst_rsqrt) {
    unsigned long long b_div = 0, b_div3 = 0, b_sqrt = 0;
    double worst = 0.0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t h0 = mix((uint32_t)i * 2654435761u + 1u), h1 = mix(h0 + (uint32_t)(i >> 32) + 0x9e3779b9u), h2 = mix(h1 ^ 0x85ebca6bu);
        const float a = rnd_float(h0, -60, 60), b = rnd_float(h1, -60, 60), c = rnd_float(h2, -60, 60);
        const float y = rcp_refined(b);
        if (__float_as_uint(div_by_rn(a, b, y)) != __float_as_uint(__fdiv_rn(a, b))) ++b_div;
        if (__float_as_uint(div_by_rn(c, b, y)) != __float_as_uint(__fdiv_rn(c, b))) ++b_div3;
        const float x = fabsf(rnd_float(h2, -100, 120));      // __fsqrt_rn's own fast-path range starts at 2^-101
        if (__float_as_uint(sqrtn_rn(x)) != __float_as_uint(__fsqrt_rn(x))) ++b_sqrt;
        const double xd = (double)fabsf(a) * (double)fabsf(c) + 1e-30;
        const double r = drsqrt_n(xd), ref = 1.0 / sqrt(xd);
        worst = fmax(worst, fabs(r - ref) / ref);
    }
    // values the kernels meet constantly: exact zeros, ones, one-hot axes
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (sqrtn_rn(0.f) != 0.f || sqrtn_rn(1.f) != 1.f || divn_rn(0.f, 3.f) != 0.f || divn_rn(1.f, 1.f) != 1.f) ++b_sqrt;
        if (!(sqrtn_rn(-1.f) != sqrtn_rn(-1.f))) ++b_sqrt;                       // NaN for negative input
        if (!(divn_rn(NAN, 2.f) != divn_rn(NAN, 2.f))) ++b_div;                  // NaN in -> NaN out
    }
    atomicAdd(bad_div, b_div); atomicAdd(bad_div3, b_div3); atomicAdd(bad_sqrt, b_sqrt);
    // positive doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(worst_rsqrt), (unsigned long long)__double_as_longlong(worst));
}

int main() {
    unsigned long long *d, h[3];
    double* dw, hw;
    cudaMalloc(&d, 3 * sizeof(unsigned long long));
    cudaMalloc(&dw, sizeof(double));
    cudaMemset(d, 0, 3 * sizeof(unsigned long long));
    cudaMemset(dw, 0, sizeof(double));
    const unsigned long long n = 1ull << 30;
    check<<<148 * 8, 256>>>(n, d, d + 1, d + 2, dw);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(&hw, dw, sizeof(hw), cudaMemcpyDeviceToHost);
    printf("{\"pairs\": %llu, \"div_mismatch\": %llu, \"div_shared_reciprocal_mismatch\": %llu, \"sqrt_mismatch\": %llu, "
           "\"drsqrt_worst_rel_err\": %.3e}\n", n, h[0], h[1], h[2], hw);
    return (h[0] | h[1] | h[2]) ? 1 : 0;
}
