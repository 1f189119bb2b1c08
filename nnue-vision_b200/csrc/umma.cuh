// umma.cuh -- device helpers shared by the tcgen05 kernels (ft_umma.cu, gemm_umma.cu): descriptors, UMMA issue and
// commit, tensor-memory allocation and loads, and the exact fp32 -> 3 x bf16 split.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace nnue {

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// eight fp32 values -> the three bf16 terms, each packed as one 16-byte chunk (k % 8 ascending)
__device__ __forceinline__ void split3x8(const float (&v)[8], uint4 (&o)[3]) {
    uint32_t h[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float r = v[j];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const __nv_bfloat16 b = __float2bfloat16_rn(r);
            h[t][j] = (uint32_t)__bfloat16_as_ushort(b);
            r -= __bfloat162float(b);
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t)
        o[t] = make_uint4(h[t][0] | (h[t][1] << 16), h[t][2] | (h[t][3] << 16), h[t][4] | (h[t][5] << 16), h[t][6] | (h[t][7] << 16));
}

// ---- an A operand produced inside the GEMM kernel ------------------------------------------------------------------
// The [128 rows x 16 k] x 3-term stage of a K-major operand built straight from a row-major fp32 source by the kernel's
// (otherwise idle) epilogue warps, instead of by a formatting kernel that writes 6 bytes per element to HBM for the
// GEMM to read back: thread r owns tile row r, reads its 16 consecutive k (64 bytes), splits them exactly and stores the
// six 16-byte chunks (canonical no-swizzle layout: chunk (k / 8) at (k / 8) * 2048 + r * 16, terms 4096 bytes apart).
// pair_half = h > 0: the source row is x[2 h] and the operand is l0[k] = k < h ? x[k] * x[k + h] : x[k - h]
// (nnue.py:660-666; h % 16 == 0 so that a k-step never straddles the halves).
struct RowChunk { float4 a[4], b[4]; };  // 16 floats of the row (+ the partner half's 16 for a product k-step)
__device__ __forceinline__ void row_chunk_load(RowChunk &c, const float *row, int k0, int h, bool live) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *p = reinterpret_cast<const float4 *>(row + (h > 0 && k0 >= h ? k0 - h : k0));
    const float4 *q = reinterpret_cast<const float4 *>(row + k0 + h);
    const bool prod = h > 0 && k0 < h;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c.a[i] = live ? __ldg(p + i) : z;
        c.b[i] = (live && prod) ? __ldg(q + i) : z;
    }
}
__device__ __forceinline__ void row_chunk_store(const RowChunk &c, bool prod, unsigned char *stage_row, uint32_t term_bytes) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4 x0 = c.a[2 * half], x1 = c.a[2 * half + 1], y0 = c.b[2 * half], y1 = c.b[2 * half + 1];
        float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        if (prod) {
            v[0] *= y0.x; v[1] *= y0.y; v[2] *= y0.z; v[3] *= y0.w; v[4] *= y1.x; v[5] *= y1.y; v[6] *= y1.z; v[7] *= y1.w;
        }
        uint4 o[3];
        split3x8(v, o);
#pragma unroll
        for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(stage_row + (uint32_t)t * term_bytes + (uint32_t)half * 2048u) = o[t];
    }
}

}  // namespace nnue
