// Microbenchmark: FP32 FMA throughput per SM as scalar FFMA vs packed FFMA2 (sm_100a), alone and
// with a broadcast LDS.128 every 8 FMAs (the mix of the feature-transformer kernels).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0 scalar, 1 packed, 2 scalar+LDS, 3 packed+LDS
__global__ void __launch_bounds__(512) k(float *out, int iters, float a, float b) {
    __shared__ float4 sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = make_float4(a, b, a, b);
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 m = make_float2(a, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 2) {
            const float4 g = sm[it & 255];
            m = make_float2(g.x, g.y);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE & 1) acc[i] = __ffma2_rn(acc[i], m, make_float2(b, a));
            else {
                acc[i].x = fmaf(acc[i].x, m.x, b);
                acc[i].y = fmaf(acc[i].y, m.y, a);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *out) {
    const int iters = 20000, grid = 148 * 2, blk = 512;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, blk>>>(out, 100, 0.999f, 1e-3f);
    cudaEventRecord(e0);
    k<MODE><<<grid, blk>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 16.0 * iters * grid * blk;
    printf("%-14s %8.3f ms  %7.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float *out; cudaMalloc(&out, 148 * 2 * 512 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FFMA+LDS", out); run<3>("FFMA2+LDS", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
