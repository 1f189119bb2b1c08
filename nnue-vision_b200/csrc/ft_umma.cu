// ft_umma.cu -- the three feature-transformer contractions on the 5th-generation tensor cores
// (tcgen05.mma issued by one thread, accumulators in tensor memory, operands fed by bulk TMA).
//
// Same arithmetic contract as ft_mma.cu: the bitmask operand is exact in bf16 (0.0 / 1.0), an fp32
// operand is split exactly into three bf16 terms, bf16 x bf16 products are exact in fp32 and the
// accumulation is fp32 -- only the summation order differs from an fp32 FMA chain.  Any L1 that is a
// multiple of 64 is served (config D: 64, the reference's "real" config: 1024).
//
//   forward          out[b, n]   = bias[n] + sum_pp bit[b, pp] W[row(pp), n]      M = samples,   K = positions
//   weight gradient  dW[pp, n]   = sum_b bit[b, pp] g[b, n]                        M = positions, K = samples
//   value gradient   gbin[b, pp] = bit[b, pp] ? sum_k g[b, k] W[row(pp), k] : 0    M = samples,   K = L1
//
// Forward and weight gradient are ONE kernel (`ft_bitgemm_umma_kernel`): a CTA owns a 128-row M tile and a
// 64-column N tile; the three bf16 terms of the fp32 operand are stacked along N (UMMA 128 x 192 x 16) and
// summed in fp32 in the epilogue.  The A operand is expanded from the bitmask by four producer warps
// straight into the canonical K-major shared-memory layout, every bit exactly once per CTA (for the
// weight gradient each 32 x 32 (sample, position) bit block is transposed in registers first so that k
// runs over samples); the B operand is pre-formatted by a small kernel and arrives by bulk TMA.
// The value gradient (`ft_gbin_umma_kernel`) is a TMA-fed GEMM of the pre-split g_ft against the
// pre-split table with the six term pairs (i, j), i + j <= 4, per k-step (UMMA 128 x 256 x 16) and a
// masked epilogue.  Two CTAs are resident per SM (<= 110 KB shared, 256 TMEM columns each), so one CTA's
// epilogue overlaps the other's main loop.
//
// Warp roles: warp 0 lane 0 = TMA producer, warp 1 = TMEM owner, lane 0 issues the UMMAs, warps 2.. = bitmask
// expansion (two groups of four warps on alternating stages: the expansion is latency-bound, not bandwidth-bound),
// then the epilogue (warp w reads TMEM lanes 32 (w % 4) .. + 31).
//
// Canonical K-major layout without swizzle, element (row r, k) of a [rows x 16] bf16 tile:
//   byte = (k / 8) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 8) * 2,  SBO = 128, LBO = rows * 16.
#include <cuda_bf16.h>

#include "common.cuh"
#include "plan.cuh"
#include "umma.cuh"

namespace nnue {

constexpr int kUM = 128;                                 // rows of every M tile (UMMA M)
constexpr int kBgThreads = 320;                          // bit GEMM: TMA warp, issuer warp, two groups of four producer warps
constexpr int kGbThreads = 192;                          // value gradient: TMA warp, issuer warp, four epilogue warps

// ---- bit GEMM (forward / weight gradient) ----
constexpr int kBgStages = 5;
constexpr uint32_t kBgABytes = kUM * 32 * 2;             // one bitmask word per row: two [128 x 16] tiles
constexpr uint32_t kBgNRows = 3 * kUmmaNCols;            // 192: the three terms of 64 columns stacked along N
constexpr uint32_t kBgBTile = kBgNRows * 32;             // 6 KB: one [192 x 16] tile
constexpr uint32_t kBgBBytes = 2 * kBgBTile;             // two k-steps per stage
constexpr uint32_t kBgNRowsInt = 2 * kUmmaNCols;         // integer accumulate: high and low byte of the int16 rows (128)
constexpr uint32_t kBgTmemCols = 256;
// ---- value gradient ----
constexpr int kGbStages = 3;
constexpr uint32_t kGbATile = kUM * 32;                  // 4 KB  [128 x 16]
constexpr uint32_t kGbBTile = kUmmaGbinN * 32;           // 8 KB  [256 x 16]
constexpr uint32_t kGbABytes = 3 * kGbATile, kGbBBytes = 3 * kGbBTile;
constexpr uint32_t kGbTmemCols = 256;
constexpr int kGbGroupM = 16;                            // sample tiles per rasterisation group (see the kernel)
constexpr int kGbRowPitch = 132;                         // floats per staged epilogue row (128 + 4: 16-byte aligned, conflict-free)

// row of the table hit by padded position pp, or -1 for padding cells (same rule as ft_mma.cu)
__device__ __forceinline__ int umma_table_row(const nnue_shape &s, int pp) {
    if (pp >= s.PP) return -1;
    const int w = pp >> 5, c = w / s.CW, cell = (w % s.CW) * 32 + (pp & 31);
    if (c >= s.C || cell >= s.Gh * s.Gw) return -1;
    return min(c * s.Gh * s.Gw + cell, s.F - 1);  // clamp of nnue.py:701
}
// ---- operand formatting ------------------------------------------------------------------------------------
// "K over rows": src[K rows][L1] fp32 -> tiles [nt = L1 / 64][ks] of [192 x 16] with N row = term * 64 + column % 64
// and k = source row (TABLE: padded position through the clamp rule; else sample).  A thread packs eight
// consecutive k of one column.
template <bool TABLE>
__global__ void umma_format_kt_kernel(const nnue_shape s, const float *__restrict__ src, int nrows, int n_ks,
                                      unsigned char *__restrict__ out) {
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2LL * n_ks * s.L1) return;
    const int col = (int)(i % s.L1), kc = (int)(i / s.L1);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kc * 8 + j;
        const int row = TABLE ? umma_table_row(s, k) : (k < nrows ? k : -1);
        v[j] = row >= 0 ? __ldg(src + (size_t)row * s.L1 + col) : 0.0f;
    }
    uint4 o[3];
    split3x8(v, o);
    const int nt = col / kUmmaNCols, c = col % kUmmaNCols, ks = kc >> 1;
    unsigned char *tile = out + ((size_t)nt * n_ks + ks) * kBgBTile + (uint32_t)(kc & 1) * (kBgNRows * 16) + (uint32_t)(c >> 3) * 128 +
                          (uint32_t)(c & 7) * 16;
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(tile + (uint32_t)t * (kUmmaNCols / 8) * 128) = o[t];
}
// "K over columns": src[rows][L1] fp32 -> tiles [rt][ks = L1 / 16][term] of [RT x 16] with k = source column.
// TABLE: row = padded position through the clamp rule, zero rows for padding.
template <bool TABLE, int RT>
__global__ void umma_format_rows_kernel(const nnue_shape s, const float *__restrict__ src, int nrows, int n_rt,
                                        unsigned char *__restrict__ out) {
    const int kcs = s.L1 / 8, n_ks = s.L1 / 16;
    // neighbouring threads take neighbouring rows of the same 8-wide k chunk: contiguous 16-byte outputs in the tile
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * n_rt * RT * kcs) return;
    const int r = (int)(i % (n_rt * RT)), kc = (int)(i / (n_rt * RT));
    const int row = TABLE ? umma_table_row(s, r) : (r < nrows ? r : -1);
    float v[8];
    if (row >= 0) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(src + (size_t)row * s.L1 + kc * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(src + (size_t)row * s.L1 + kc * 8) + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    }
    uint4 o[3];
    split3x8(v, o);
    const int rt = r / RT, rr = r % RT, ks = kc >> 1;
    unsigned char *tile = out + ((size_t)rt * n_ks + ks) * 3 * (RT * 32) + (uint32_t)(kc & 1) * (RT * 16) + (uint32_t)(rr >> 3) * 128 +
                          (uint32_t)(rr & 7) * 16;
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(tile + (uint32_t)t * (RT * 32)) = o[t];
}

// eight bits -> eight bf16 (0.0 / 1.0), k ascending
__device__ __forceinline__ uint4 bits8_to_bf16x8(uint32_t byte) {
    uint4 v;
    v.x = ((byte & 1u) ? 0x3F80u : 0u) | ((byte & 2u) ? 0x3F800000u : 0u);
    v.y = ((byte & 4u) ? 0x3F80u : 0u) | ((byte & 8u) ? 0x3F800000u : 0u);
    v.z = ((byte & 16u) ? 0x3F80u : 0u) | ((byte & 32u) ? 0x3F800000u : 0u);
    v.w = ((byte & 64u) ? 0x3F80u : 0u) | ((byte & 128u) ? 0x3F800000u : 0u);
    return v;
}

// ---- forward / weight gradient ---------------------------------------------------------------------------------
// grid = (M tiles, L1 / 64, K chunks).  DW = false: M = samples, stage j = bitmask word j of the sample rows,
// out = ft_out [B][L1] (+ bias).  DW = true: M = padded positions (bitmask words 4 x .. 4 x + 3), stage j = the
// 32-sample block chunk * chunk_blocks + j, out = partial[chunk][P + 1][L1] (row P = bias gradient) for the fold kernels.
// MODE 2 (integer inference, qinfer.cu): same as the forward but the B operand is the int16 table as TWO exact bf16
// terms (high byte, signed; low byte, unsigned: UMMA 128 x 128 x 16), bias is int32 [L1] and out is int16 [B][L1]:
// out = (int16)(bias + 256 * sum_hi + sum_lo), the engine's wrap-around accumulate (simd_scalar.cpp:78-95).
// MODE 3: MODE 2 with L1 = 64 (one N tile holds the whole accumulator row) and the rest of the integer path in the epilogue:
// clipped ReLU, pairwise, the three dense layers (dp4a), logits / 64 -- `out` is unused, qa.logits is written.
// MODE 2 / 3 walk only the bitmask words that can hold a bit (qa: the words of a channel past the conv raster are skipped
// when the threshold is not negative): stage jj = word (jj / cw_used) * cw_all + jj % cw_used.
template <int MODE>
__global__ void __launch_bounds__(kBgThreads, 2)
ft_bitgemm_umma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const unsigned char *__restrict__ btiles,
                       const float *__restrict__ bias, float *__restrict__ out, int chunk_blocks, const QAccArgs qa) {
    constexpr bool DW = MODE == 1;
    constexpr bool INT = MODE >= 2;
    constexpr uint32_t NR = INT ? kBgNRowsInt : kBgNRows, BTile = NR * 32, BBytes = 2 * BTile;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full_a = reinterpret_cast<uint64_t *>(smem_raw);     // [ST] producer warps -> issuer
    uint64_t *full_b = full_a + kBgStages;                         // [ST] TMA -> issuer
    uint64_t *empty = full_b + kBgStages;                          // [ST] UMMA commit -> producers / TMA
    uint64_t *done = empty + kBgStages;                            // accumulator complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    unsigned char *sa = smem_raw + 1024;                           // [ST][8 KB]
    unsigned char *sb = sa + kBgStages * kBgABytes;                // [ST][12 KB]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int mt = blockIdx.x, nt = blockIdx.y, chunk = blockIdx.z;
    int first, n_stage, n_ks_total;
    if (DW) {
        first = chunk * chunk_blocks;
        n_stage = min(s.BW, first + chunk_blocks) - first;
        n_ks_total = 2 * s.BW;
    } else {
        first = 0;
        n_stage = INT ? qa.n_words_used : s.NW;
        n_ks_total = 2 * s.NW;
    }
    auto word_of = [&](int jj) -> int { return INT ? (jj / qa.cw_used) * qa.cw_all + jj % qa.cw_used : jj; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < kBgStages; ++i) {
            mbar_init(&full_a[i], 4);
            mbar_init(&full_b[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kBgTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA: the B tiles of every stage ----
            const unsigned char *src = btiles + ((size_t)nt * n_ks_total + 2 * (size_t)first) * BTile;
            for (int j = 0; j < n_stage; ++j) {
                const int st = j % kBgStages;
                if (j >= kBgStages) mbar_wait(&empty[st], ((j / kBgStages) - 1) & 1);
                mbar_arrive_expect_tx(&full_b[st], BBytes);
                tma_bulk_g2s(sb + (uint32_t)st * BBytes, src + (size_t)word_of(j) * BBytes, BBytes, &full_b[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- issuer: two UMMAs (128 x 192 x 16) per stage ----
            constexpr uint32_t idesc = umma_idesc(kUM, NR);
            for (int j = 0; j < n_stage; ++j) {
                const int st = j % kBgStages;
                const uint32_t ph = (j / kBgStages) & 1;
                mbar_wait(&full_a[st], ph);
                mbar_wait(&full_b[st], ph);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(sa + (uint32_t)st * kBgABytes), b_base = smem_u32(sb + (uint32_t)st * BBytes);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma_bf16(tmem_acc, umma_smem_desc(a_base + (uint32_t)ks * (kUM * 32), kUM * 16, 128),
                              umma_smem_desc(b_base + (uint32_t)ks * BTile, NR * 16, 128), idesc, (j | ks) ? 1u : 0u);
                umma_commit(&empty[st]);  // arrives when the UMMAs above have read their shared-memory operands
            }
            umma_commit(done);
        }
    } else {
        // ---- A producers: warp quadrant q owns tile rows 32 q .. 32 q + 31, lane = row (after the transpose);
        //      group grp takes the stages j = grp, grp + 2, ... ----
        const int q = warp & 3, row = q * 32 + lane, grp = (warp - 2) >> 2;
        const int wi = mt * 4 + q;  // DW: my bitmask word
        // DW: one padding row of the M tiles is all ones, so that its output row is the column sum of g_ft (bias gradient)
        const bool ones_row = DW && (mt * kUM + row == umma_bias_pp(s));
        auto load = [&](int j) -> uint32_t {
            if (j >= n_stage) return 0u;
            if (DW) {
                const int b = (first + j) * 32 + lane;
                return (wi < s.NW && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + wi) : 0u;
            }
            const int b = mt * kUM + row;
            return b < s.B ? __ldg(bits_s + (size_t)b * s.NW + word_of(j)) : 0u;
        };
        unsigned char *dst0 = sa + (uint32_t)row * 16;
        uint32_t pre[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pre[u] = load(grp + 2 * u);
        for (int j0 = grp; j0 < n_stage; j0 += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 2 * u;
                if (j < n_stage) {  // warp-uniform
                    const int st = j % kBgStages;
                    if (j >= kBgStages) mbar_wait(&empty[st], ((j / kBgStages) - 1) & 1);
                    uint32_t word = pre[u];
                    pre[u] = load(j + 8);
                    if (DW) {
                        word = warp_bit_transpose(word, lane);  // lane L: the 32 samples of position 32 wi + L
                        if (ones_row) word = 0xFFFFFFFFu;       // (samples past B are zero rows of the B operand)
                    }
                    unsigned char *dst = dst0 + (uint32_t)st * kBgABytes;
#pragma unroll
                    for (int c = 0; c < 4; ++c)  // chunk c: k 8c .. 8c+7 -> k-step c / 2, k half c % 2
                        *reinterpret_cast<uint4 *>(dst + (uint32_t)(c >> 1) * (kUM * 32) + (uint32_t)(c & 1) * (kUM * 16)) =
                            bits8_to_bf16x8((word >> (8 * c)) & 0xFFu);
                    fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_a[st]);
                }
            }
        }
        // ---- epilogue: (term 2 + term 1) + term 0 per column; group grp takes 32 of the 64 columns of its rows ----
        mbar_wait(done, 0);
        tcgen05_fence_after();
        const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
        float *orow = nullptr;
        if (DW) {
            const int pp = mt * kUM + row, w = pp >> 5, cells = s.Gh * s.Gw;
            if (ones_row) {
                orow = out + ((size_t)chunk * (s.P + 1) + s.P) * s.L1 + nt * kUmmaNCols;
            } else if (w < s.NW) {
                const int c = w / s.CW, cell = (w % s.CW) * 32 + (pp & 31);
                if (cell < cells) orow = out + ((size_t)chunk * (s.P + 1) + (size_t)c * cells + cell) * s.L1 + nt * kUmmaNCols;
            }
        } else {
            const int b = mt * kUM + row;
            if (b < s.B) orow = out + (size_t)b * s.L1 + nt * kUmmaNCols;
        }
        if (MODE == 3) {
            // ---- the layer stack on the accumulator row (nnue_engine.cpp:726-729, 490-533).  The two column groups of a
            // row meet through shared memory (the A ring is idle: every UMMA has completed); per row 65 words (odd pitch:
            // the 32 rows of a warp hit 32 banks): clipped int16 x 64 | pairwise int8 x 64 | layer-1 out | layer-2 out
            constexpr int kPitch = 65, oPw = 32, oH1 = 48, oH2 = 56;
            uint32_t *srow = reinterpret_cast<uint32_t *>(sa) + row * kPitch;
            const int b = mt * kUM + row;
            const int32_t *bias32 = reinterpret_cast<const int32_t *>(bias);
            auto qbar = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };  // the eight epilogue warps
#pragma unroll
            for (int cc = 0; cc < kUmmaNCols / 2; cc += 16) {
                const int c0 = grp * (kUmmaNCols / 2) + cc;
                float hi[16], lo[16];
                tmem_ld16(tbase + (uint32_t)c0, hi);
                tmem_ld16(tbase + (uint32_t)(kUmmaNCols + c0), lo);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int v0 = __ldg(bias32 + c0 + 2 * u) + 256 * __float2int_rn(hi[2 * u]) + __float2int_rn(lo[2 * u]);
                    const int v1 = __ldg(bias32 + c0 + 2 * u + 1) + 256 * __float2int_rn(hi[2 * u + 1]) + __float2int_rn(lo[2 * u + 1]);
                    const int x0 = max(0, min(qa.qone, (int)(int16_t)v0)), x1 = max(0, min(qa.qone, (int)(int16_t)v1));  // int16 wrap, clipped ReLU
                    srow[(c0 >> 1) + u] = (uint32_t)x0 | ((uint32_t)x1 << 16);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) srow[(grp ? oH2 : oH1) + k] = 0u;  // zero padding of the dense layers' inputs
            qbar();
            {   // pairwise words: group 0 the 32 products, group 1 the 32 clipped first-half values
                const int16_t *c16 = reinterpret_cast<const int16_t *>(srow);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t wd = 0u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * k + e;
                        const int x = c16[i];
                        const int v = grp == 0 ? max(0, min(127, (x * (int)c16[i + 32]) / 128)) : min(127, x);
                        wd |= (uint32_t)v << (8 * e);
                    }
                    srow[oPw + 8 * grp + k] = wd;
                }
            }
            qbar();
            {   // layer 1: float divide then truncate (simd_scalar.cpp:131-133), clamp 0..127; each group half of the outputs
                int pw[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) pw[k] = (int)srow[oPw + k];
                const int hn = (qa.L2 + 1) / 2, o1 = min(qa.L2, (grp + 1) * hn);
                int8_t *h1 = reinterpret_cast<int8_t *>(srow + oH1);
                for (int o = grp * hn; o < o1; ++o) {
                    int a = __ldg(qa.b1 + o);
#pragma unroll
                    for (int k = 0; k < 16; ++k) a = __dp4a(pw[k], __ldg(qa.w1 + k * qa.L2 + o), a);
                    h1[o] = (int8_t)max(0, min(127, __float2int_rz(__fdiv_rn(__int2float_rn(a), qa.l1_scale))));
                }
            }
            qbar();
            {   // layer 2: integer divide (truncating), clamp +-127, ReLU (nnue_engine.cpp:512-523)
                const int hn = (qa.L3 + 1) / 2, o1 = min(qa.L3, (grp + 1) * hn);
                int8_t *h2 = reinterpret_cast<int8_t *>(srow + oH2);
                for (int o = grp * hn; o < o1; ++o) {
                    int a = __ldg(qa.b2 + o);
                    for (int k = 0; k < qa.K2; ++k) a = __dp4a((int)srow[oH1 + k], __ldg(qa.w2 + k * qa.L3 + o), a);
                    h2[o] = (int8_t)max(0, max(-127, min(127, a / qa.l2_iscale)));
                }
            }
            qbar();
            {   // output: (float)acc / output_scale (nnue_engine.cpp:526-533)
                const int hn = (qa.NC + 1) / 2, c1 = min(qa.NC, (grp + 1) * hn);
                for (int c = grp * hn; c < c1; ++c) {
                    int a = __ldg(qa.bo + c);
                    for (int k = 0; k < qa.K3; ++k) a = __dp4a((int)srow[oH2 + k], __ldg(qa.wo + k * qa.NC + c), a);
                    if (b < s.B) qa.logits[(size_t)b * qa.NC + c] = __fdiv_rn(__int2float_rn(a), qa.out_scale);
                }
            }
        } else if (MODE == 2) {
            const int b = mt * kUM + row;
            int16_t *orow16 = reinterpret_cast<int16_t *>(out) + (size_t)min(b, s.B - 1) * s.L1 + nt * kUmmaNCols;
            const int32_t *bias32 = reinterpret_cast<const int32_t *>(bias) + nt * kUmmaNCols;
#pragma unroll
            for (int cc = 0; cc < kUmmaNCols / 2; cc += 16) {
                const int c0 = grp * (kUmmaNCols / 2) + cc;
                float hi[16], lo[16];
                tmem_ld16(tbase + (uint32_t)c0, hi);
                tmem_ld16(tbase + (uint32_t)(kUmmaNCols + c0), lo);
                tmem_ld_wait();
                uint32_t pk[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int v0 = __ldg(bias32 + c0 + 2 * u) + 256 * __float2int_rn(hi[2 * u]) + __float2int_rn(lo[2 * u]);
                    const int v1 = __ldg(bias32 + c0 + 2 * u + 1) + 256 * __float2int_rn(hi[2 * u + 1]) + __float2int_rn(lo[2 * u + 1]);
                    pk[u] = ((uint32_t)v0 & 0xFFFFu) | ((uint32_t)v1 << 16);
                }
                if (b < s.B) {
                    reinterpret_cast<uint4 *>(orow16 + c0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    reinterpret_cast<uint4 *>(orow16 + c0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        } else
#pragma unroll
        for (int cc = 0; cc < kUmmaNCols / 2; cc += 16) {
            const int c0 = grp * (kUmmaNCols / 2) + cc;
            float t0[16], t1[16], t2[16];
            tmem_ld16(tbase + (uint32_t)c0, t0);
            tmem_ld16(tbase + (uint32_t)(kUmmaNCols + c0), t1);
            tmem_ld16(tbase + (uint32_t)(2 * kUmmaNCols + c0), t2);
            tmem_ld_wait();
            if (orow) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float4 r = make_float4((t2[4 * v] + t1[4 * v]) + t0[4 * v], (t2[4 * v + 1] + t1[4 * v + 1]) + t0[4 * v + 1],
                                           (t2[4 * v + 2] + t1[4 * v + 2]) + t0[4 * v + 2], (t2[4 * v + 3] + t1[4 * v + 3]) + t0[4 * v + 3]);
                    if (!DW) r = f4_add(r, __ldg(reinterpret_cast<const float4 *>(bias + nt * kUmmaNCols + c0) + v));
                    reinterpret_cast<float4 *>(orow + c0)[v] = r;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kBgTmemCols>(tmem_acc);
}

// ---- value gradient -----------------------------------------------------------------------------------------------
// grid = (256-position N tiles, 128-sample M tiles).  K = L1 in k-steps of 16: a stage holds the three terms of the
// g_ft tile (12 KB) and of the table tile (24 KB); six UMMAs (128 x 256 x 16) per stage.
__global__ void __launch_bounds__(kGbThreads, 2)
ft_gbin_umma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const unsigned char *__restrict__ atiles,
                    const unsigned char *__restrict__ btiles, float *__restrict__ gbin, const float *__restrict__ g_ft_src) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kGbStages;
    uint64_t *done = empty + kGbStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    unsigned char *sa = smem_raw + 1024;
    unsigned char *sb = sa + kGbStages * kGbABytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_ks = s.L1 / 16;
    // Tile order: CTAs are dispatched in linear order, and the ~300 resident ones should share operand tiles through L2.
    // Groups of kGbGroupM sample tiles are walked position-tile by position-tile, so a wave touches ~16 A tiles and ~19
    // B tiles (tens of MB) instead of 296 different B tiles (ncu at SURVEY config I, one position tile per CTA in x
    // order: 12.8 GB of DRAM reads per launch for 0.4 GB of table tiles, 80 % of the HBM peak beside an 88 % busy
    // tensor pipe).
    int nt, mt;
    {
        const int n_nt = gridDim.x, n_mt = gridDim.y;
        const int lid = blockIdx.x + n_nt * blockIdx.y, group = kGbGroupM * n_nt;
        const int first_m = (lid / group) * kGbGroupM, gm = min(kGbGroupM, n_mt - first_m);
        mt = first_m + (lid % group) % gm;
        nt = (lid % group) / gm;
    }

    // g_ft_src != null: the A operand (g_ft as three bf16 terms) is built in the kernel by the epilogue warps from the fp32
    // rows (umma.cuh: row_chunk_*) instead of arriving as pre-formatted tiles
    const bool a_inline = g_ft_src != nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kGbStages; ++i) {
            mbar_init(&full[i], a_inline ? 5 : 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<kGbTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char *asrc = atiles + (size_t)mt * n_ks * kGbABytes, *bsrc = btiles + (size_t)nt * n_ks * kGbBBytes;
            for (int j = 0; j < n_ks; ++j) {
                const int st = j % kGbStages;
                if (j >= kGbStages) mbar_wait(&empty[st], ((j / kGbStages) - 1) & 1);
                mbar_arrive_expect_tx(&full[st], a_inline ? kGbBBytes : kGbABytes + kGbBBytes);
                if (!a_inline) tma_bulk_g2s(sa + (uint32_t)st * kGbABytes, asrc + (size_t)j * kGbABytes, kGbABytes, &full[st]);
                tma_bulk_g2s(sb + (uint32_t)st * kGbBBytes, bsrc + (size_t)j * kGbBBytes, kGbBBytes, &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kUM, kUmmaGbinN);
            for (int j = 0; j < n_ks; ++j) {
                const int st = j % kGbStages;
                mbar_wait(&full[st], (j / kGbStages) & 1);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(sa + (uint32_t)st * kGbABytes), b_base = smem_u32(sb + (uint32_t)st * kGbBBytes);
                // term pairs (split of g, split of W) with i + j <= 2 (0-based): the rest is below 2^-27 relative
#pragma unroll
                for (int ta = 0; ta < 3; ++ta)
#pragma unroll
                    for (int tw = 0; ta + tw < 3; ++tw)
                        umma_bf16(tmem_acc, umma_smem_desc(a_base + (uint32_t)ta * kGbATile, kUM * 16, 128),
                                  umma_smem_desc(b_base + (uint32_t)tw * kGbBTile, kUmmaGbinN * 16, 128), idesc, (j | ta | tw) ? 1u : 0u);
                umma_commit(&empty[st]);
            }
            umma_commit(done);
        }
    } else {
        // ---- epilogue: mask with the sample's bits, transpose through shared memory (the operand ring is idle
        //      once `done` fires) so that every global store instruction writes 512 contiguous bytes of one row ----
        const int q = warp & 3, b = mt * kUM + q * 32 + lane;
        if (a_inline) {  // A producers (thread = tile row): k-step j + 1 is loading while k-step j is split and stored
            const bool live = b < s.B;
            const float *row = g_ft_src + (size_t)min(b, s.B - 1) * s.L1;
            unsigned char *dst0 = sa + (uint32_t)(q * 32 + lane) * 16;
            RowChunk nxt;
            row_chunk_load(nxt, row, 0, 0, live);
            for (int j = 0; j < n_ks; ++j) {
                const int st = j % kGbStages;
                const RowChunk cur = nxt;
                if (j + 1 < n_ks) row_chunk_load(nxt, row, (j + 1) * 16, 0, live);
                if (j >= kGbStages) mbar_wait(&empty[st], ((j / kGbStages) - 1) & 1);
                row_chunk_store(cur, false, dst0 + (uint32_t)st * kGbABytes, kGbATile);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st]);
            }
        }
        uint32_t mw[kUmmaGbinN / 32];
#pragma unroll
        for (int i = 0; i < kUmmaGbinN / 32; ++i) {
            const int w = nt * (kUmmaGbinN / 32) + i;
            mw[i] = (b < s.B && w < s.NW) ? __ldg(bits_s + (size_t)b * s.NW + w) : 0u;
        }
        mbar_wait(done, 0);
        tcgen05_fence_after();
        const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
        float *stg = reinterpret_cast<float *>(smem_raw + 1024) + q * (32 * kGbRowPitch);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[2][16];
                tmem_ld16(tbase + (uint32_t)(half * 128 + c0), v[0]);
                tmem_ld16(tbase + (uint32_t)(half * 128 + c0 + 16), v[1]);
                tmem_ld_wait();
                const uint32_t m = mw[(half * 128 + c0) >> 5];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float *x = &v[u >> 2][4 * (u & 3)];
                    const uint32_t mm = m >> (4 * u);
                    *reinterpret_cast<float4 *>(stg + lane * kGbRowPitch + c0 + 4 * u) =
                        make_float4(mm & 1u ? x[0] : 0.0f, mm & 2u ? x[1] : 0.0f, mm & 4u ? x[2] : 0.0f, mm & 8u ? x[3] : 0.0f);
                }
            }
            __syncwarp();
            const int pp = nt * kUmmaGbinN + half * 128 + lane * 4;
            if (pp < s.PP) {
                float *o = gbin + (size_t)(mt * kUM + q * 32) * s.PP + pp;
                const int nrow = min(32, s.B - (mt * kUM + q * 32));
#pragma unroll 8
                for (int r = 0; r < nrow; ++r)
                    *reinterpret_cast<float4 *>(o + (size_t)r * s.PP) = *reinterpret_cast<const float4 *>(stg + r * kGbRowPitch + lane * 4);
            }
            __syncwarp();
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kGbTmemCols>(tmem_acc);
}

// ---- host side ---------------------------------------------------------------------------------------------------
constexpr size_t kBgSmem = 1024 + (size_t)kBgStages * (kBgABytes + kBgBBytes);
constexpr size_t kGbSmem = 1024 + (size_t)kGbStages * (kGbABytes + kGbBBytes);

// The table in both tile orders, formatted once per step: [forward / kt tiles: umma_kt_bytes(PP)] [value-gradient /
// rows tiles: ceil(PP / 256) * 256 * L1 * 6].  It depends only on the weights, so a caller may run it on a side stream
// while the images are being extracted (nnue_ft_format_tables).
int launch_ft_format_tables(const nnue_shape &s, const float *w, void *tables, int which, cudaStream_t st) {
    unsigned char *wt = static_cast<unsigned char *>(tables);
    if (which & 1) {
        const int n_ks = 2 * s.NW;
        const long long n = 2LL * n_ks * s.L1;
        umma_format_kt_kernel<true><<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, s.PP, n_ks, wt);
        NNUE_CHECK_LAUNCH("umma_format_kt_kernel");
    }
    if (which & 2) {
        const int n_nt = ceil_div(s.PP, kUmmaGbinN);
        const long long n = 1LL * n_nt * kUmmaGbinN * (s.L1 / 8);
        umma_format_rows_kernel<true, kUmmaGbinN><<<(int)((n + 255) / 256), 256, 0, st>>>(
            s, w, s.PP, n_nt, wt + align_up(umma_kt_bytes((size_t)s.PP, s), 256));
        NNUE_CHECK_LAUNCH("umma_format_rows_kernel");
    }
    return NNUE_OK;
}

// the table as [PP / 128][L1 / 16][3][128 x 16] tiles: the A operand of the one-kernel input gradient (input_bwd_fused.cu)
int launch_format_table_rows128(const nnue_shape &s, const float *w, unsigned char *out, cudaStream_t st) {
    const int n_rt = ceil_div(s.PP, kUM);
    const long long n = 1LL * n_rt * kUM * (s.L1 / 8);
    umma_format_rows_kernel<true, kUM><<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, s.PP, n_rt, out);
    NNUE_CHECK_LAUNCH("umma_format_rows_kernel");
    return NNUE_OK;
}

// out = bias + bits . W from pre-formatted forward tiles
int launch_ft_fwd_umma_tiles(const nnue_shape &s, const uint32_t *bits_s, const void *tables, const float *bias, float *out,
                             cudaStream_t st) {
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_bitgemm_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBgSmem));
    ft_bitgemm_umma_kernel<0><<<dim3(ceil_div(s.B, kUM), s.L1 / kUmmaNCols, 1), kBgThreads, kBgSmem, st>>>(
        s, bits_s, static_cast<const unsigned char *>(tables), bias, out, 0, QAccArgs{});
    NNUE_CHECK_LAUNCH("ft_bitgemm_umma_kernel");
    return NNUE_OK;
}

// out = bias + bits . W;  workspace: the table as [L1 / 64][PP / 16] tiles (umma_kt_bytes(PP, L1))
int launch_ft_fwd_umma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *bias, float *out,
                       void *workspace, cudaStream_t st) {
    const int rc = launch_ft_format_tables(s, w, workspace, 1, st);
    if (rc != NNUE_OK) return rc;
    return launch_ft_fwd_umma_tiles(s, bits_s, workspace, bias, out, st);
}

// Fold stage 1: every position p sums its partials over the K chunks (loads batched eight deep: the loop is
// latency-bound).  Positions below F-1 are table rows and go straight to g_w; positions >= F-1 all alias onto the
// last row (the clamp of nnue.py:701) and are parked in `alias` [P-(F-1)][L1] for stage 2; rows in [P, F-1) get
// zeros; the extra row P of every partial block is the bias gradient.
__global__ void __launch_bounds__(256)
umma_dw_fold_kernel(const nnue_shape s, int n_chunks, const float *__restrict__ partial, float *__restrict__ g_w,
                    float *__restrict__ g_b, float *__restrict__ alias) {
    const int v4 = s.L1 / 4;
    const int nrows = max(s.P, s.F - 1);
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * (nrows + 1) * v4) return;
    const int r = (int)(i / v4), c4 = (int)(i % v4);
    const int p = r == nrows ? s.P : r;  // row of the partial blocks
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r == nrows || r < s.P) {
        const float4 *src = reinterpret_cast<const float4 *>(partial + (size_t)p * s.L1) + c4;
        const size_t pitch = (size_t)(s.P + 1) * v4;
        int c = 0;
        for (; c + 8 <= n_chunks; c += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(c + u) * pitch);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = f4_add(acc, v[u]);
        }
        for (; c < n_chunks; ++c) acc = f4_add(acc, __ldg(src + (size_t)c * pitch));
    }
    if (r == nrows) reinterpret_cast<float4 *>(g_b)[c4] = acc;
    else if (r < s.F - 1) reinterpret_cast<float4 *>(g_w + (size_t)r * s.L1)[c4] = acc;
    else reinterpret_cast<float4 *>(alias + (size_t)(r - (s.F - 1)) * s.L1)[c4] = acc;
}
// Fold stage 2: g_w[F-1] = sum of the parked rows, 32 slices per column combined in a fixed order.
__global__ void __launch_bounds__(1024)
umma_dw_fold_last_kernel(const nnue_shape s, const float *__restrict__ alias, float *__restrict__ g_w) {
    __shared__ float4 red[1024];
    const int v4 = s.L1 / 4;
    const int nalias = max(0, s.P - (s.F - 1));
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c4 = blockIdx.x * 32 + col;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < v4)
#pragma unroll 4
        for (int i = sl; i < nalias; i += 32)
            acc = f4_add(acc, __ldg(reinterpret_cast<const float4 *>(alias + (size_t)i * s.L1) + c4));
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sl == 0 && c4 < v4) {
        for (int k = 1; k < 32; ++k) acc = f4_add(acc, red[k * 32 + col]);
        reinterpret_cast<float4 *>(g_w + (size_t)(s.F - 1) * s.L1)[c4] = acc;
    }
}

// g_w = bits^T . g_ft, g_b = column sums of g_ft.  `ws`: split g_ft^T tiles | partial[chunk][P + 1][L1] | alias rows
// (ws_ft_dw_umma bytes)
int launch_ft_bwd_dw_umma(const nnue_shape &s, const uint32_t *bits_s, const float *g_ft, void *ws, float *g_w, float *g_b,
                          cudaStream_t st) {
    const UmmaDwPlan p = plan_ft_dw_umma(s);
    unsigned char *gt = static_cast<unsigned char *>(ws);
    float *partial = reinterpret_cast<float *>(gt + align_up(umma_kt_bytes((size_t)s.BW * 32, s), 256));
    float *alias = partial + (size_t)p.n_chunks * (s.P + 1) * s.L1;
    const int n_ks = 2 * s.BW;
    long long n = 2LL * n_ks * s.L1;
    umma_format_kt_kernel<false><<<(int)((n + 255) / 256), 256, 0, st>>>(s, g_ft, s.B, n_ks, gt);
    NNUE_CHECK_LAUNCH("umma_format_kt_kernel");
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_bitgemm_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBgSmem));
    const int m_rows = umma_bias_pp(s) + 1 > s.PP ? s.PP + 1 : s.PP;  // the all-ones row may need one more M tile
    ft_bitgemm_umma_kernel<1><<<dim3(ceil_div(m_rows, kUM), s.L1 / kUmmaNCols, p.n_chunks), kBgThreads, kBgSmem, st>>>(
        s, bits_s, gt, nullptr, partial, p.chunk_blocks, QAccArgs{});
    NNUE_CHECK_LAUNCH("ft_bitgemm_umma_kernel");
    n = 1LL * ((s.P > s.F - 1 ? s.P : s.F - 1) + 1) * (s.L1 / 4);
    umma_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s, p.n_chunks, partial, g_w, g_b, alias);
    NNUE_CHECK_LAUNCH("umma_dw_fold_kernel");
    umma_dw_fold_last_kernel<<<ceil_div(s.L1 / 4, 32), 1024, 0, st>>>(s, alias, g_w);
    NNUE_CHECK_LAUNCH("umma_dw_fold_last_kernel");
    return NNUE_OK;
}

// Integer accumulate of qinfer.cu: acc16[b] = (int16)(bias + sum over set bits of the int16 table rows).
// bits [B][NW] (word w = 32 k-indices), tiles [L1 / 64][2 NW][128 x 16] (high / low byte terms, built at load time).
// qa: which bitmask words are walked and, with qa.logits set (L1 = 64, small stacks: q_stack_fused_ok), the layer stack
// in the epilogue -- acc16 is then unused.
int launch_q_accumulate_umma(int B, int NW, int L1, const uint32_t *bits, const unsigned char *tiles, const int32_t *bias,
                             int16_t *acc16, const QAccArgs &qa, cudaStream_t st) {
    nnue_shape s{};
    s.B = B; s.NW = NW; s.L1 = L1; s.PP = 32 * NW; s.BW = ceil_div(B, 32);
    constexpr size_t smem = 1024 + (size_t)kBgStages * (kBgABytes + 2 * kBgNRowsInt * 32);
    if (qa.logits) {
        if (L1 != kUmmaNCols) return NNUE_ERR_UNSUPPORTED;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_bitgemm_umma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ft_bitgemm_umma_kernel<3><<<dim3(ceil_div(B, kUM), 1, 1), kBgThreads, smem, st>>>(
            s, bits, tiles, reinterpret_cast<const float *>(bias), nullptr, 0, qa);
    } else {
        NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_bitgemm_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ft_bitgemm_umma_kernel<2><<<dim3(ceil_div(B, kUM), L1 / kUmmaNCols, 1), kBgThreads, smem, st>>>(
            s, bits, tiles, reinterpret_cast<const float *>(bias), reinterpret_cast<float *>(acc16), 0, qa);
    }
    NNUE_CHECK_LAUNCH("ft_bitgemm_umma_kernel");
    return NNUE_OK;
}

// gbin = bits ? g_ft . W^T : 0;  workspace: split g_ft tiles | split table tiles (ws_ft_gbin_umma)
// table_tiles != null: the value-gradient table tiles are already formatted (launch_ft_format_tables)
int launch_ft_bwd_gbin_umma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *g_ft, void *workspace,
                            float *gbin, cudaStream_t st, const void *table_tiles) {
    const int n_mt = ceil_div(s.B, kUM), n_nt = ceil_div(s.PP, kUmmaGbinN);
    unsigned char *at = static_cast<unsigned char *>(workspace);
    unsigned char *bt = at + align_up((size_t)n_mt * kUM * s.L1 * 6, 256);
    long long n = 1LL * n_mt * kUM * (s.L1 / 8);
    // wide tables: the kernel's epilogue warps split g_ft themselves (64 k-steps give them time; at L1 = 64 the four
    // k-steps do not, and the 5 us formatter stays)
    const bool a_inline = get_option(kOptInlineA) && s.L1 >= 256 && (reinterpret_cast<uintptr_t>(g_ft) & 15) == 0;
    if (!a_inline) {
        umma_format_rows_kernel<false, kUM><<<(int)((n + 255) / 256), 256, 0, st>>>(s, g_ft, s.B, n_mt, at);
        NNUE_CHECK_LAUNCH("umma_format_rows_kernel");
    }
    if (table_tiles) {
        bt = const_cast<unsigned char *>(static_cast<const unsigned char *>(table_tiles)) + align_up(umma_kt_bytes((size_t)s.PP, s), 256);
    } else {
        n = 1LL * n_nt * kUmmaGbinN * (s.L1 / 8);
        umma_format_rows_kernel<true, kUmmaGbinN><<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, s.PP, n_nt, bt);
        NNUE_CHECK_LAUNCH("umma_format_rows_kernel");
    }
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_gbin_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGbSmem));
    ft_gbin_umma_kernel<<<dim3(n_nt, n_mt), kGbThreads, kGbSmem, st>>>(s, bits_s, at, bt, gbin, a_inline ? g_ft : nullptr);
    NNUE_CHECK_LAUNCH("ft_gbin_umma_kernel");
    return NNUE_OK;
}

}  // namespace nnue
