// Microbenchmark: legacy warp-level mma.sync throughput on sm_100a (bf16 m16n8k16, tf32 m16n8k8),
// FMA per clock per SM, at 4 / 8 / 16 warps per SM with 4 independent accumulator tiles per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k(float *out, int iters, long long *cycles) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    float c[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = c[t][2] = c[t][3] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + t), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + t), "r"(b1));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps : {4, 8, 16, 32}) {
            if (kind == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double fma = (kind == 0 ? 16 * 8 * 16 : 16 * 8 * 8) * 8.0 * iters * warps;
            printf("%s warps/SM %2d: %8.1f FMA/clk/SM  (%.2f cycles per mma per warp)\n", kind == 0 ? "bf16 m16n8k16" : "tf32 m16n8k8 ",
                   warps, fma / c, (double)c / (8.0 * iters));
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
