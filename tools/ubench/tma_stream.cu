// Microbenchmark: HBM -> shared-memory streaming rate of a bulk-TMA ring (cp.async.bulk + mbarrier), one
// CTA per SM, as a function of the copy size and of the ring depth.  Consumers only wait and release.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void expect_tx(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void *dst, const void *src, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(n), "r"(s32(b)) : "memory");
}

__global__ void __launch_bounds__(256, 1) stream(const char *src, size_t total, int chunk, int copies, int ST, float *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *full = (uint64_t *)sm, *empty = full + 32;
    char *stages = (char *)sm + 512;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const size_t stage_bytes = (size_t)chunk * copies;
    const long long n_stage = total / stage_bytes;
    const long long mine = (n_stage - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto produce = [&](long long ii) {
        const int st = ii % ST;
        if (ii >= ST) wait(&empty[st], ((ii / ST) - 1) & 1);
        expect_tx(&full[st], (uint32_t)stage_bytes);
        const char *g = src + (blockIdx.x + ii * gridDim.x) * stage_bytes;
        for (int c = 0; c < copies; ++c) bulk(stages + (size_t)st * stage_bytes + (size_t)c * chunk, g + (size_t)c * chunk, chunk, &full[st]);
    };
    const int ahead = ST - 1;
    if (threadIdx.x == 0) for (long long ii = 0; ii < ahead && ii < mine; ++ii) produce(ii);
    float acc = 0.f;
    for (long long i = 0; i < mine; ++i) {
        if (threadIdx.x == 0 && i + ahead < mine) produce(i + ahead);
        __syncwarp();
        const int st = i % ST;
        wait(&full[st], (i / ST) & 1);
        acc += ((float *)(stages + (size_t)st * stage_bytes))[threadIdx.x];
        __syncwarp();
        if (lane == 0) arrive(&empty[st]);
    }
    if (acc == 123.456f) sink[0] = acc;
    (void)warp;
}

int main() {
    const size_t total = (size_t)402 << 20;
    char *src; float *sink;
    cudaMalloc(&src, total); cudaMalloc(&sink, 4);
    cudaMemset(src, 1, total);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int chunk, copies, ST; } cfgs[] = {{4096, 5, 8}, {4096, 5, 10}, {4096, 1, 32}, {12288, 1, 16}, {20480, 1, 10}, {32768, 1, 6}, {65536, 1, 3}, {4096, 3, 16}, {2048, 10, 8}};
    for (auto c : cfgs) {
        const size_t smem = 512 + (size_t)c.chunk * c.copies * c.ST;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            stream<<<148, 256, smem>>>(src, total, c.chunk, c.copies, c.ST, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("chunk %6d x %2d per stage, %2d stages (%3zu KB smem): %7.1f us  %6.2f TB/s  %s\n", c.chunk, c.copies, c.ST, smem >> 10,
               ms * 1e3, total / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
