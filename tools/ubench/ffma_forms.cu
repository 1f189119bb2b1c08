// Microbenchmark: FP32 FMA issue rate per SM for the operand patterns the feature-transformer
// kernels use, at 8 / 16 / 32 warps per SM.  Prints warp-FFMAs per cycle per SM (peak 4).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma_forms ffma_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: acc[i] = fma(on, g[i], acc[i])   (one operand shared by all: reuse-friendly), g in registers
// MODE 1: d[c]  = fma(w[i], g[i], d[c])     (two fresh operands per FFMA, 4 chains)
// MODE 2: MODE 0 with g from a broadcast LDS.128 per 4 FFMAs
// MODE 3: MODE 0 + MODE 1 interleaved, g from LDS.128 per 8 FFMAs (the merged kernel's mix)
// MODE 4: 4-row register tile: d[r] = fma(w[i], g[r][i], d[r]) (w reused 4x), g from LDS.128
template <int MODE>
__global__ void k(float *out, int iters, float a, long long *cycles) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(a, a * 0.5f, a * 0.25f, a * 2.f);
    __syncthreads();
    float acc[64], w[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) { acc[i] = i * 1e-3f; w[i] = a + i * 1e-4f; }
    float d0 = 0, d1 = 0, d2 = 0, d3 = 0;
    float on = a > 0 ? 1.f : 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float4 *row = sm + (it & 31) * 16;
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i] = fmaf(on, w[i], acc[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                d0 = fmaf(w[i], acc[i], d0); d1 = fmaf(w[i + 1], acc[i + 1], d1);
                d2 = fmaf(w[i + 2], acc[i + 2], d2); d3 = fmaf(w[i + 3], acc[i + 3], d3);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                const float4 g = row[v];
                acc[4 * v] = fmaf(on, g.x, acc[4 * v]); acc[4 * v + 1] = fmaf(on, g.y, acc[4 * v + 1]);
                acc[4 * v + 2] = fmaf(on, g.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on, g.w, acc[4 * v + 3]);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                const float4 g = row[v];
                acc[4 * v] = fmaf(on, g.x, acc[4 * v]); acc[4 * v + 1] = fmaf(on, g.y, acc[4 * v + 1]);
                acc[4 * v + 2] = fmaf(on, g.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on, g.w, acc[4 * v + 3]);
                d0 = fmaf(w[4 * v], g.x, d0); d1 = fmaf(w[4 * v + 1], g.y, d1);
                d2 = fmaf(w[4 * v + 2], g.z, d2); d3 = fmaf(w[4 * v + 3], g.w, d3);
            }
        } else {
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                const float4 g0 = row[v], g1 = row[16 + v], g2 = row[32 + v], g3 = row[48 + v];
                d0 = fmaf(w[4 * v], g0.x, d0); d1 = fmaf(w[4 * v], g1.x, d1); d2 = fmaf(w[4 * v], g2.x, d2); d3 = fmaf(w[4 * v], g3.x, d3);
                d0 = fmaf(w[4 * v + 1], g0.y, d0); d1 = fmaf(w[4 * v + 1], g1.y, d1); d2 = fmaf(w[4 * v + 1], g2.y, d2); d3 = fmaf(w[4 * v + 1], g3.y, d3);
                d0 = fmaf(w[4 * v + 2], g0.z, d0); d1 = fmaf(w[4 * v + 2], g1.z, d1); d2 = fmaf(w[4 * v + 2], g2.z, d2); d3 = fmaf(w[4 * v + 2], g3.z, d3);
                d0 = fmaf(w[4 * v + 3], g0.w, d0); d1 = fmaf(w[4 * v + 3], g1.w, d1); d2 = fmaf(w[4 * v + 3], g2.w, d2); d3 = fmaf(w[4 * v + 3], g3.w, d3);
            }
        }
    }
    long long t1 = clock64();
    float s = d0 + d1 + d2 + d3;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, float *out, long long *cyc, int warps) {
    const int iters = 4000;
    k<MODE><<<148, warps * 32>>>(out, 10, 0.999f, cyc);
    k<MODE><<<148, warps * 32>>>(out, iters, 0.999f, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ffma_per_iter = MODE == 3 ? 128 : 64;
    const double lds = MODE == 2 || MODE == 3 ? 16 : MODE == 4 ? 64 : 0;
    const double f = MODE == 4 ? 256 : ffma_per_iter;
    printf("%-34s warps/SM %2d: %6.2f warp-FFMA/clk/SM  (%.0f cycles per iteration of %g FFMA + %g LDS.128 per warp)\n", name,
           warps, f * iters * warps / (double)c, (double)c / iters, f, lds);
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int warps : {8, 16, 32}) {
        run<0>("acc += on*w (shared operand)", out, cyc, warps);
        run<1>("d += w*g (two fresh operands)", out, cyc, warps);
        run<2>("acc += on*g, g by LDS.128/4", out, cyc, warps);
        run<3>("merged mix, LDS.128/8", out, cyc, warps);
        run<4>("4-row tile, w reused, LDS.128/4", out, cyc, warps);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
