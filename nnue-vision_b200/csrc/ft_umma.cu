// ft_umma.cu -- the feature-transformer forward on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same arithmetic contract as ft_mma.cu: out = bias + bits . (W1 + W2 + W3), the bitmask exact in bf16,
// the table split exactly into three bf16 terms, exact products, fp32 accumulation -- but issued as
// tcgen05.mma (UMMA) 128 x 64 x 16 tiles by ONE thread, with the accumulator in tensor memory:
//
//   * CTA = one tile of 128 samples x all 64 columns; K = padded positions, 64 per pipeline stage.
//   * A operand (the bitmask as bf16): eight producer warps expand 32 bits -> 64 B per (sample, word) and
//     store them as the canonical K-major / no-swizzle UMMA layout (8 x 16-byte core matrices);
//     every bit is expanded exactly once per step (the warp-level MMA path re-expands it per column
//     group and per lane: ncu showed that integer work, not the tensor pipe, as its limit).
//   * B operand (the split table, pre-formatted in the same canonical layout by a small kernel): one
//     bulk TMA copy of 24 KB per stage (4 k-steps x 3 terms x 2 KB).
//   * One elected thread of the issuer warp waits for both, issues 12 UMMAs per stage and commits
//     them to the stage's `empty` mbarrier; after the last stage it commits to `done`.
//   * Epilogue: warps 0-3 read their 32 TMEM lanes (tcgen05.ld 32x32b), add the bias and store rows.
//
// Canonical K-major layout without swizzle, element (row r, k) of a [rows x 16] bf16 tile:
//   byte = (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2,  SBO = 128, LBO = rows * 16.
#include <cuda_bf16.h>

#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kUmmaM = 128;          // samples per CTA tile (UMMA M)
constexpr int kUmmaN = 64;           // columns (UMMA N) = L1
constexpr int kUmmaKStage = 64;      // positions per pipeline stage (two bitmask words, four k16 steps)
constexpr int kUmmaStages = 4;
constexpr int kUmmaProducerWarps = 8;
constexpr int kUmmaThreads = (kUmmaProducerWarps + 1) * 32;
constexpr uint32_t kUmmaABytes = kUmmaM * kUmmaKStage * 2;         // 16 KB: four [128 x 16] tiles
constexpr uint32_t kUmmaBBytes = 3 * 4 * kUmmaN * 16 * 2;          // 24 KB: (k-step, term) tiles of 2 KB
constexpr uint32_t kUmmaTmemCols = 64;

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, K-major both, N = 64, M = 128
constexpr uint32_t kUmmaIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((kUmmaN >> 3) << 17) | ((kUmmaM >> 4) << 24);

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

// row of the table hit by padded position pp, or -1 for padding cells (same rule as ft_mma.cu)
__device__ __forceinline__ int umma_table_row(const nnue_shape &s, int pp) {
    const int w = pp >> 5, c = w / s.CW, cell = (w % s.CW) * 32 + (pp & 31);
    if (c >= s.C || cell >= s.Gh * s.Gw) return -1;
    return min(c * s.Gh * s.Gw + cell, s.F - 1);
}

// B operand: out[stage][k-step (4)][term (3)] tiles of [64 columns x 16 positions] bf16 in the canonical layout
__global__ void umma_format_table_kernel(const nnue_shape s, const float *__restrict__ w, uint16_t *__restrict__ out) {
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;  // one (position, column)
    if (i >= 1LL * s.PP * kUmmaN) return;
    const int pp = (int)(i / kUmmaN), n = (int)(i % kUmmaN);
    const int row = umma_table_row(s, pp);
    float r = row >= 0 ? __ldg(w + (size_t)row * s.L1 + n) : 0.0f;
    const int kstep = pp / 16, k = pp % 16;
    const size_t elem = (size_t)(k / 8) * (kUmmaN * 8) + (size_t)(n / 8) * 64 + (n % 8) * 8 + (k % 8);  // in bf16 units
#pragma unroll
    for (int sp = 0; sp < 3; ++sp) {
        const __nv_bfloat16 b = __float2bfloat16_rn(r);
        out[((size_t)kstep * 3 + sp) * (kUmmaN * 16) + elem] = __bfloat16_as_ushort(b);
        r -= __bfloat162float(b);
    }
}

__global__ void __launch_bounds__(kUmmaThreads, 1)
ft_fwd_umma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint16_t *__restrict__ wtiles,
                   const float *__restrict__ bias, float *__restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full_a = reinterpret_cast<uint64_t *>(smem_raw);       // [ST] producers -> issuer
    uint64_t *full_b = full_a + kUmmaStages;                         // [ST] TMA -> issuer
    uint64_t *empty = full_b + kUmmaStages;                          // [ST] UMMA commit -> producers / TMA
    uint64_t *done = empty + kUmmaStages;                            // accumulator complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    unsigned char *sa = smem_raw + 1024;                             // [ST][16 KB]
    unsigned char *sb = sa + kUmmaStages * kUmmaABytes;              // [ST][24 KB]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = blockIdx.x * kUmmaM;
    const int n_stage = s.PP / kUmmaKStage;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kUmmaStages; ++i) {
            mbar_init(&full_a[i], kUmmaProducerWarps);
            mbar_init(&full_b[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == kUmmaProducerWarps) {  // the issuer warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kUmmaTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp < kUmmaProducerWarps) {
        // ---- A producers: thread = (sample row m, word half wh) of the stage ----
        const int m = threadIdx.x & (kUmmaM - 1), wh = threadIdx.x >> 7;  // 256 threads: 128 rows x 2 words
        const int b = b0 + m;
        const uint32_t *brow = bits_s + (size_t)min(b, s.B - 1) * s.NW;
        // my four 16-byte chunks inside a stage: word wh -> k-steps 2 wh, 2 wh + 1; chunk (k-step, k half)
        unsigned char *dst0 = sa + (uint32_t)(2 * wh) * (kUmmaM * 32) + (uint32_t)(m / 8) * 128 + (uint32_t)(m % 8) * 16;
        for (int st_i = 0; st_i < n_stage; ++st_i) {
            const int st = st_i % kUmmaStages;
            if (st_i >= kUmmaStages) mbar_wait(&empty[st], ((st_i / kUmmaStages) - 1) & 1);
            const uint32_t word = b < s.B ? __ldg(brow + st_i * 2 + wh) : 0u;
            unsigned char *dst = dst0 + (uint32_t)st * kUmmaABytes;
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // chunk c: bits 8c .. 8c+7 -> k-step c / 2, k half c % 2
                const uint32_t byte = (word >> (8 * c)) & 0xFFu;
                uint4 v;
                v.x = ((byte & 1u) ? 0x3F80u : 0u) | ((byte & 2u) ? 0x3F800000u : 0u);
                v.y = ((byte & 4u) ? 0x3F80u : 0u) | ((byte & 8u) ? 0x3F800000u : 0u);
                v.z = ((byte & 16u) ? 0x3F80u : 0u) | ((byte & 32u) ? 0x3F800000u : 0u);
                v.w = ((byte & 64u) ? 0x3F80u : 0u) | ((byte & 128u) ? 0x3F800000u : 0u);
                *reinterpret_cast<uint4 *>(dst + (uint32_t)(c / 2) * (kUmmaM * 32) + (uint32_t)(c % 2) * (kUmmaM * 16)) = v;
            }
            fence_async_smem();  // my generic-proxy stores must be visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[st]);
        }
        // ---- epilogue: warps 0-3 own TMEM lanes 32 w .. 32 w + 31 (= sample rows) ----
        if (warp < 4) {
            mbar_wait(done, 0);
            tcgen05_fence_after();
            const int row = b0 + warp * 32 + lane;
            float *orow = out + (size_t)row * s.L1;
#pragma unroll
            for (int c0 = 0; c0 < kUmmaN; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < s.B) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + c0) + q);
                        reinterpret_cast<float4 *>(orow + c0)[q] =
                            make_float4(bv.x + __uint_as_float(v[4 * q]), bv.y + __uint_as_float(v[4 * q + 1]),
                                        bv.z + __uint_as_float(v[4 * q + 2]), bv.w + __uint_as_float(v[4 * q + 3]));
                    }
                }
            }
        }
    } else if (lane == 0) {
        // ---- issuer: TMA for B, then the UMMAs of each stage ----
        const int ahead = kUmmaStages - 1;
        auto load_b = [&](int ii) {
            const int st = ii % kUmmaStages;
            if (ii >= kUmmaStages) mbar_wait(&empty[st], ((ii / kUmmaStages) - 1) & 1);
            mbar_arrive_expect_tx(&full_b[st], kUmmaBBytes);
            tma_bulk_g2s(sb + (uint32_t)st * kUmmaBBytes, reinterpret_cast<const unsigned char *>(wtiles) + (size_t)ii * kUmmaBBytes,
                         kUmmaBBytes, &full_b[st]);
        };
        for (int ii = 0; ii < ahead && ii < n_stage; ++ii) load_b(ii);
        for (int st_i = 0; st_i < n_stage; ++st_i) {
            const int st = st_i % kUmmaStages;
            const uint32_t ph = (st_i / kUmmaStages) & 1;
            mbar_wait(&full_a[st], ph);
            mbar_wait(&full_b[st], ph);
            tcgen05_fence_after();
            const uint32_t a_base = smem_u32(sa + (uint32_t)st * kUmmaABytes), b_base = smem_u32(sb + (uint32_t)st * kUmmaBBytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t adesc = umma_smem_desc(a_base + (uint32_t)ks * (kUmmaM * 32), kUmmaM * 16, 128);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) {
                    const uint64_t bdesc = umma_smem_desc(b_base + (uint32_t)(ks * 3 + sp) * (kUmmaN * 32), kUmmaN * 16, 128);
                    umma_bf16(tmem_acc, adesc, bdesc, kUmmaIdesc, (st_i | ks | sp) ? 1u : 0u);
                }
            }
            umma_commit(&empty[st]);  // arrives when the UMMAs above have read their shared-memory operands
            if (st_i + ahead < n_stage) load_b(st_i + ahead);  // (waits for the commit of the previous stage)
        }
        umma_commit(done);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == kUmmaProducerWarps) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "n"(kUmmaTmemCols) : "memory");
    }
}

int launch_ft_fwd_umma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *bias, float *out,
                       void *workspace, cudaStream_t st) {
    uint16_t *wtiles = static_cast<uint16_t *>(workspace);
    const long long n = 1LL * s.PP * kUmmaN;
    umma_format_table_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, wtiles);
    NNUE_CHECK_LAUNCH("umma_format_table_kernel");
    const size_t smem = 1024 + (size_t)kUmmaStages * (kUmmaABytes + kUmmaBBytes);
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ft_fwd_umma_kernel<<<ceil_div(s.B, kUmmaM), kUmmaThreads, smem, st>>>(s, bits_s, wtiles, bias, out);
    NNUE_CHECK_LAUNCH("ft_fwd_umma_kernel");
    return NNUE_OK;
}

}  // namespace nnue
