// Microbenchmark: issue throughput of FFMA (3 register operands) vs FFMA2 (packed fp32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o ffma2_tput ffma2_tput.cu && ./ffma2_tput   (cudart linked dynamically; the binary is git-ignored)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {            // scalar FFMA x2 (same flops as one FFMA2)
                    acc[i].x = fmaf(acc[i].x, av.x, bv.x);
                    acc[i].y = fmaf(acc[i].y, av.y, bv.y);
                } else if (MODE == 1) {     // FFMA2, all three operands register pairs
                    acc[i] = __ffma2_rn(acc[i], av, bv);
                } else if (MODE == 2) {     // FFMA2 with broadcast scalar operands
                    acc[i] = __ffma2_rn(acc[i], make_float2(a, a), make_float2(b, b));
                } else {                    // FFMA2 with three distinct pair operands per instruction
                    acc[i] = __ffma2_rn(acc[i], acc[(i + 3) & 7], acc[(i + 5) & 7]);
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sms * 1024 * 4);
    const int iters = 20000;
    k<MODE><<<sms, warps_per_sm * 32>>>(out, 100, 0.999f, 0.001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_per_lane = (double)iters * 4 * 8 * 2;                       // fp32 FMAs per thread
    double cycles = ms * 1e-3 * clk * 1e3;
    double fma_per_clk_sm = fma_per_lane * warps_per_sm * 32 / cycles;
    printf("%-34s warps/SM %2d  %.3f ms  fp32 FMA/clk/SM %.1f  (TFLOP/s %.1f)\n", name, warps_per_sm, ms, fma_per_clk_sm,
           fma_per_clk_sm * 2 * sms * clk * 1e3 / 1e12);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("FFMA (scalar, 3 reg operands)", w);
        run<1>("FFMA2 (3 pair operands)", w);
        run<2>("FFMA2 (2 broadcast operands)", w);
        run<3>("FFMA2 (3 distinct pairs)", w);
    }
    return 0;
}
