// Micro-benchmark (dev tool): FFMA vs packed FFMA2 (fma.rn.f32x2) throughput per SM on sm_100a -- decides which one the
// register-tiled GEMM micro-kernels (epoch2.cu, rollout_fused.cu) should use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_bench tools/fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0 scalar FFMA, 1 FFMA2
__global__ void k(float* out, long long* cyc, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 1) acc[i] = __ffma2_rn(aa, acc[i], bb);
                else { acc[i].x = fmaf(a, acc[i].x, b); acc[i].y = fmaf(a, acc[i].y, b); }
            }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 148 * 8);
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; ++mode) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.999f, 0.001f);
                else k<1><<<1, warps * 32>>>(d_out, d_cyc, iters, 0.999f, 0.001f);
                cudaDeviceSynchronize();
                cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
            }
            const double fma = (double)iters * 4 * 16 * 2 * warps * 32;
            printf("%2d warps/SM  %-6s  %8lld cycles  %6.1f FMA/clk/SM\n", warps, mode ? "FFMA2" : "FFMA", c, fma / (double)c);
        }
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
