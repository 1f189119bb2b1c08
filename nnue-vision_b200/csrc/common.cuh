// common.cuh -- shared device/host helpers for libnnue_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nnue_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnnue_b200 is written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace nnue {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Records the failing CUDA call for nnue_last_cuda_error(); defined in api.cu.
void note_cuda_error(cudaError_t e, const char *what);

#define NNUE_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) {                              \
            ::nnue::note_cuda_error(_e, #expr);               \
            return NNUE_ERR_CUDA;                             \
        }                                                     \
    } while (0)

// every kernel launch goes through this macro, which also feeds nnue_launch_count()
void note_launch();
#define NNUE_CHECK_LAUNCH(name)                               \
    do {                                                      \
        ::nnue::note_launch();                                \
        cudaError_t _e = cudaGetLastError();                  \
        if (_e != cudaSuccess) {                              \
            ::nnue::note_cuda_error(_e, name);                \
            return NNUE_ERR_CUDA;                             \
        }                                                     \
    } while (0)

// Tuning knobs settable through nnue_set_option (api.cu); every value has a working default.
enum Option { kOptFtFwdStaging = 0, kOptDwOwner, kOptInputFused, kOptInputVariant, kOptHeadFused, kOptFtBwdBoth, kOptFtMma, kOptExtractTma, kOptExtractFixed, kOptFtUmma, kOptInputSwizzle, kOptHeadUmma, kOptQTcMinBatch, kOptFtForm, kOptFtDensity, kOptFtGather, kOptFtGatherVariant, kOptInputRows, kOptGatherUnits, kOptInputFusedGbin, kOptQCtaMaxBatch, kOptQConvFixed, kOptQStackFused, kOptHeadMid, kOptHeadPairEpi, kOptInlineA, kOptConvBwdPacked, kNumOptions };
int get_option(int which);

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// 32x32 bit-matrix transpose across a warp: lane i passes row i, lane i returns column i
// (bit r of the result = bit i of lane r's input).  Five butterfly stages of one shuffle each.
__device__ __forceinline__ unsigned warp_bit_transpose(unsigned x, int lane) {
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        const unsigned m = sft == 16 ? 0xFFFF0000u : sft == 8 ? 0xFF00FF00u : sft == 4 ? 0xF0F0F0F0u
                         : sft == 2 ? 0xCCCCCCCCu : 0xAAAAAAAAu;
        const unsigned o = __shfl_xor_sync(kFull, x, sft);
        x = (lane & sft) ? ((x & m) | ((o >> sft) & ~m)) : ((x & ~m) | ((o << sft) & m));
    }
    return x;
}

__device__ __forceinline__ float4 f4_add(float4 a, const float4 b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    return a;
}

// ---- mbarrier + 1-D bulk TMA copy (cp.async.bulk -> SASS UBLKCP) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the barrier init visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "NNUE_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NNUE_DONE_%=;\n\t"
        "bra NNUE_WAIT_%=;\n\t"
        "NNUE_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage `bytes` (multiple of 16) from global into shared with bulk TMA copies issued by one
// thread, completion on `bar` (phase `parity`).  All threads of the CTA must call it.
__device__ __forceinline__ void tma_stage(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                          uint32_t parity) {
    constexpr uint32_t kChunk = 32768;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
            tma_bulk_g2s(static_cast<char *>(smem_dst) + off, static_cast<const char *>(gmem_src) + off, n, bar);
        }
    }
    mbar_wait(bar, parity);
}

}  // namespace nnue
