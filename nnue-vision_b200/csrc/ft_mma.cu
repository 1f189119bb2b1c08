// ft_mma.cu -- the three feature-transformer contractions on the tensor cores, at fp32 accuracy.
//
// ncu on the CUDA-core kernels (profiles/r1_ncu_v5*) and tools/ubench show why: at config D the
// "sparse" transformer is a 43 %-dense 0/1 contraction against a 205 KB table -- a real dense
// contraction at the benchmark batch.  On the FP32 pipe it costs 1.0 G FMA per kernel and is capped
// by the broadcast rate of shared memory (an LDS.128 read by all lanes sustains 0.46 / clk / SM);
// warp-level bf16 MMA runs at 1010 FMA / clk / SM on B200, eight times the FP32 pipe.
//
// Accuracy: the 1e-5 parity bar rules out plain bf16 / tf32 operands, but not the tensor core:
//   * the bitmask operand is exactly representable (0.0 / 1.0 in bf16);
//   * an fp32 value splits EXACTLY into three bf16 terms x = x1 + x2 + x3 (8 + 8 + 8 mantissa bits);
//   * bf16 x bf16 products are exact in fp32, accumulation is fp32.
// So out = bits . (W1 + W2 + W3) is three MMAs whose products are exact, and <g, W> uses the six
// term pairs (i, j), i + j <= 4 (the dropped pairs are below 2^-27 relative).  Only the
// summation order differs from an fp32 FMA chain.  K per accumulator is kept short (one table
// pass, or one K-chunk of samples folded later in fp32), so accumulation error stays ~1e-7.
//
// Operands are pre-formatted into MMA fragment order by small kernels, column-group major, so that
// the slice a CTA needs is one contiguous block: it is staged in shared memory with bulk TMA copies
// once per CTA and a lane then reads its B fragments as conflict-free LDS.128.  The bitmask is
// expanded to bf16 A fragments in registers (a few integer instructions per register, amortised
// over 12-24 MMAs).  (A first version loaded fragments straight from L2: ncu showed 14-16 stall
// cycles per issue on the long scoreboard and 13-19 % issue utilisation, profiles/r1_ncu_v6*.)
//
//   mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32, lane = 4 * g + t (g = lane / 4, t = lane % 4):
//     A (16 x 16): a0 = (row g,     k 2t, 2t+1)   a1 = (row g + 8, k 2t, 2t+1)
//                  a2 = (row g,     k 2t+8, +9)   a3 = (row g + 8, k 2t+8, +9)
//     B (16 x 8):  b0 = (k 2t, 2t+1, col g)       b1 = (k 2t+8, 2t+9, col g)
//     C (16 x 8):  c0, c1 = (row g, col 2t, 2t+1) c2, c3 = (row g + 8, col 2t, 2t+1)
#include <cuda_bf16.h>

#include "common.cuh"
#include "plan.cuh"

namespace nnue {

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x = h[0] + h[1] + h[2] exactly (round-to-nearest bf16 at each step; the residuals are exact in fp32)
__device__ __forceinline__ void split3(float x, uint32_t (&h)[3]) {
    float r = x;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const __nv_bfloat16 b = __float2bfloat16_rn(r);
        h[s] = (uint32_t)__bfloat16_as_ushort(b);
        r -= __bfloat162float(b);
    }
}
// two bits -> two bf16 values (0.0 / 1.0): bit 0 in the low half (the lower k index)
__device__ __forceinline__ uint32_t bits2_to_bf16x2(uint32_t x) {
    return ((x & 1u) * 0x3F80u) | ((x & 2u) * 0x1FC00000u);
}

// row of the table hit by padded position pp (bit pp % 32 of word pp / 32), or -1 for padding cells
__device__ __forceinline__ int table_row_of(const nnue_shape &s, int pp) {
    const int w = pp >> 5, c = w / s.CW, cell = (w % s.CW) * 32 + (pp & 31);
    if (c >= s.C || cell >= s.Gh * s.Gw) return -1;
    return min(c * s.Gh * s.Gw + cell, s.F - 1);  // clamp of nnue.py:701
}

// ---- operand formatting ----------------------------------------------------------------------------------
// B operand with k running over ROWS of a row-major fp32 matrix src[rows][L1] and n over its columns
// (forward: rows = padded positions of the table; weight gradient: rows = samples of g_ft).
// out[(((grp * n_kb + kb) * 3 + s) * kMmaGP + q) * 32 + lane] = uint4 {b0, b1 of column block 2 nbp, b0, b1 of
// 2 nbp + 1} with nbp = grp * kMmaGP + q: everything one 32-column group needs is contiguous.
template <bool TABLE>
__global__ void mma_format_rows_kernel(const nnue_shape s, const float *__restrict__ src, int nrows, int n_kb,
                                       uint4 *__restrict__ out) {
    const int NBP = s.L1 / 16;
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * n_kb * NBP * 32) return;
    const int lane = (int)(i & 31), nbp = (int)((i >> 5) % NBP), kb = (int)((i >> 5) / NBP);
    const int g = lane >> 2, t = lane & 3;
    uint32_t packed[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {          // e: {b0 of nb0, b1 of nb0, b0 of nb1, b1 of nb1}
        const int col = (nbp * 2 + (e >> 1)) * 8 + g;
        const int k0 = kb * 16 + t * 2 + (e & 1) * 8;
        uint32_t lo[3], hi[3];
        float v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = k0 + q;
            int row = k;
            if (TABLE) row = table_row_of(s, k);
            else if (k >= nrows) row = -1;
            v[q] = row >= 0 ? __ldg(src + (size_t)row * s.L1 + col) : 0.0f;
        }
        split3(v[0], lo);
        split3(v[1], hi);
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) packed[sp][e] = lo[sp] | (hi[sp] << 16);
    }
#pragma unroll
    for (int sp = 0; sp < 3; ++sp)
        out[((((size_t)(nbp / kMmaGP) * n_kb + kb) * 3 + sp) * kMmaGP + nbp % kMmaGP) * 32 + lane] =
            make_uint4(packed[sp][0], packed[sp][1], packed[sp][2], packed[sp][3]);
}

// B operand with k running over the COLUMNS of the table and n over padded positions (value gradient):
// out[((nb * 3 + s) * (KB / 2) + kbp) * 32 + lane] = uint4 {b0, b1 of k block 2 kbp, b0, b1 of 2 kbp + 1}.
__global__ void mma_format_cols_kernel(const nnue_shape s, const float *__restrict__ w, uint4 *__restrict__ out) {
    const int KBP = s.L1 / 32;
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * (s.PP / 8) * KBP * 32) return;
    const int lane = (int)(i & 31), kbp = (int)((i >> 5) % KBP), nb = (int)((i >> 5) / KBP);
    const int g = lane >> 2, t = lane & 3;
    const int row = table_row_of(s, nb * 8 + g);
    uint32_t packed[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int k0 = (kbp * 2 + (e >> 1)) * 16 + t * 2 + (e & 1) * 8;
        uint32_t lo[3], hi[3];
        split3(row >= 0 ? __ldg(w + (size_t)row * s.L1 + k0) : 0.0f, lo);
        split3(row >= 0 ? __ldg(w + (size_t)row * s.L1 + k0 + 1) : 0.0f, hi);
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) packed[sp][e] = lo[sp] | (hi[sp] << 16);
    }
#pragma unroll
    for (int sp = 0; sp < 3; ++sp)
        out[((size_t)(nb * 3 + sp) * KBP + kbp) * 32 + lane] = make_uint4(packed[sp][0], packed[sp][1], packed[sp][2], packed[sp][3]);
}

// Stage `bytes` of fragments into shared memory behind a 128-byte mbarrier header; all threads call it.
__device__ __forceinline__ const uint4 *stage_fragments(unsigned char *smem_raw, const void *src, uint32_t bytes) {
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    tma_stage(smem_raw + 128, src, bytes, bar, 0);
    return reinterpret_cast<const uint4 *>(smem_raw + 128);
}

// bits of two rows -> the four A registers of one m16 tile for k half `sh`
__device__ __forceinline__ void bits_to_afrag(uint32_t (&a)[4], uint32_t w_lo, uint32_t w_hi, int sh) {
    a[0] = bits2_to_bf16x2(w_lo >> sh);
    a[1] = bits2_to_bf16x2(w_hi >> sh);
    a[2] = bits2_to_bf16x2(w_lo >> (sh + 8));
    a[3] = bits2_to_bf16x2(w_hi >> (sh + 8));
}

// ---- forward: out[b] = bias + bits[b] . Wc ------------------------------------------------------------------
// A persistent CTA owns one group of 32 columns: its slice of the split table (PP/16 k-steps x 3 terms x
// 1 KB = 192 KB at config D) sits in shared memory; each warp takes 16-sample tiles (one m16 tile),
// K runs over the padded positions one bitmask word (two k16 steps) at a time.  Every A fragment built
// from the bitmask feeds 12 MMAs (the integer pipe that expands bits to bf16 is the scarce resource:
// ncu on the 16-column version showed math-pipe throttle, not tensor-pipe, as the top stall).
__global__ void __launch_bounds__(kMmaFwdThreads, 1)
ft_fwd_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ wfrag,
                  const float *__restrict__ bias, float *__restrict__ out) {
    constexpr int GP = kMmaGP, NB = 2 * GP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int groups = s.L1 / (16 * GP), n_kb = s.PP / 16;
    const int grp = blockIdx.x % groups, cta = blockIdx.x / groups, ctas = gridDim.x / groups;
    const uint4 *sf = stage_fragments(smem_raw, wfrag + (size_t)grp * n_kb * 3 * GP * 32, (uint32_t)n_kb * 3 * GP * 512u) + lane;
    const int col0 = grp * 16 * GP + t * 2;
    // 16-sample tiles (one m16): with 32-sample tiles only half of the warps of the machine would have work
    const int ntiles = ceil_div(s.B, 16), wpc = kMmaFwdThreads / 32;
    for (int tile = cta * wpc + warp; tile < ntiles; tile += ctas * wpc) {
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        float acc[NB][4];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.0f;
        const uint32_t *p0 = bits_s + (size_t)min(r0, s.B - 1) * s.NW, *p1 = bits_s + (size_t)min(r1, s.B - 1) * s.NW;
        uint32_t c0 = r0 < s.B ? __ldg(p0) : 0u, c1 = r1 < s.B ? __ldg(p1) : 0u;
        for (int w = 0; w < s.NW; ++w) {
            // the next word's bits fly while this word's MMAs issue
            const uint32_t n0 = (w + 1 < s.NW && r0 < s.B) ? __ldg(p0 + w + 1) : 0u;
            const uint32_t n1 = (w + 1 < s.NW && r1 < s.B) ? __ldg(p1 + w + 1) : 0u;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t a[4];
                bits_to_afrag(a, c0, c1, half * 16 + t * 2);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                    for (int q = 0; q < GP; ++q) {
                        const uint4 f = sf[(((2 * w + half) * 3 + sp) * GP + q) * 32];
                        mma_bf16(acc[2 * q], a, f.x, f.y);
                        mma_bf16(acc[2 * q + 1], a, f.z, f.w);
                    }
            }
            c0 = n0;
            c1 = n1;
        }
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const float2 bv = __ldg(reinterpret_cast<const float2 *>(bias + col0 + nb * 8));
            if (r0 < s.B)
                *reinterpret_cast<float2 *>(out + (size_t)r0 * s.L1 + col0 + nb * 8) = make_float2(bv.x + acc[nb][0], bv.y + acc[nb][1]);
            if (r1 < s.B)
                *reinterpret_cast<float2 *>(out + (size_t)r1 * s.L1 + col0 + nb * 8) = make_float2(bv.x + acc[nb][2], bv.y + acc[nb][3]);
        }
    }
}

// ---- weight gradient: dW[p] = sum_b bits[b, p] g_ft[b] --------------------------------------------------------
// CTA = (block of kMmaDwWarps bitmask words, 32-column group, K-chunk of samples): the chunk's slice of the
// split g_ft (chunk_blocks x 2 k-steps x 3 terms x 1 KB) is staged in shared memory once; a warp owns one
// word (32 positions = two m16 tiles) and transposes each 32 x 32 bit block of (sample, position) in
// registers so that k runs over samples.  Output: partial[chunk][p][L1] for the fold kernels of ft.cu
// (they also resolve the aliasing onto row F-1).  The warp of word 0 also multiplies an all-ones tile:
// its first row is the bias gradient.
__global__ void __launch_bounds__(kMmaDwWarps * 32, 2)
ft_bwd_dw_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ gfrag,
                     float *__restrict__ partial, float *__restrict__ bias_partial, int chunk_blocks) {
    constexpr int GP = kMmaGP, NB = 2 * GP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int n_kb = s.BW * 2;
    const int w = blockIdx.x * kMmaDwWarps + warp;  // my bitmask word
    const int grp = blockIdx.y, chunk = blockIdx.z;
    const bool active = w < s.NW;
    const bool do_bias = w == 0;
    const int sb_begin = chunk * chunk_blocks, sb_end = min(s.BW, sb_begin + chunk_blocks);  // 32-sample blocks
    const uint4 *sf = stage_fragments(smem_raw, gfrag + ((size_t)grp * n_kb + 2 * sb_begin) * 3 * GP * 32,
                                      (uint32_t)(sb_end - sb_begin) * 2 * 3 * GP * 512u) + lane;
    float acc[2][NB][4], accb[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) acc[mt][nb][0] = acc[mt][nb][1] = acc[mt][nb][2] = acc[mt][nb][3] = 0.0f;
        accb[nb][0] = accb[nb][1] = accb[nb][2] = accb[nb][3] = 0.0f;
    }
    const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};
    auto fetch = [&](int sb) -> uint32_t {
        const int b = sb * 32 + lane;
        return (active && sb < sb_end && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + w) : 0u;
    };
    uint32_t x = fetch(sb_begin);
    for (int sb = sb_begin; sb < sb_end; ++sb) {
        const uint32_t xn = fetch(sb + 1);
        const uint32_t tr = warp_bit_transpose(x, lane);  // lane L: the 32 samples of position w*32 + L
        uint32_t word[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) word[mt][h] = __shfl_sync(kFull, tr, mt * 16 + g + 8 * h);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) bits_to_afrag(a[mt], word[mt][0], word[mt][1], half * 16 + t * 2);
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int q = 0; q < GP; ++q) {
                    const uint4 f = sf[(((2 * (sb - sb_begin) + half) * 3 + sp) * GP + q) * 32];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_bf16(acc[mt][2 * q], a[mt], f.x, f.y);
                        mma_bf16(acc[mt][2 * q + 1], a[mt], f.z, f.w);
                    }
                    if (do_bias) {  // warp-uniform
                        mma_bf16(accb[2 * q], ones, f.x, f.y);
                        mma_bf16(accb[2 * q + 1], ones, f.z, f.w);
                    }
                }
        }
        x = xn;
    }
    const int col0 = grp * 16 * GP + t * 2;
    if (active) {
        const int cells = s.Gh * s.Gw, c = w / s.CW, cell0 = (w % s.CW) * 32;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cell = cell0 + mt * 16 + g + 8 * h;
                if (cell >= cells) continue;
                float *row = partial + ((size_t)chunk * s.P + (size_t)c * cells + cell) * s.L1 + col0;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                    *reinterpret_cast<float2 *>(row + nb * 8) = make_float2(acc[mt][nb][2 * h], acc[mt][nb][2 * h + 1]);
            }
    }
    if (do_bias && g == 0) {
        float *row = bias_partial + (size_t)chunk * s.L1 + col0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) *reinterpret_cast<float2 *>(row + nb * 8) = make_float2(accb[nb][0], accb[nb][1]);
    }
}

// ---- value gradient: gbin[b, pp] = bit(b, pp) ? <Wc[pp], g_ft[b]> : 0 ------------------------------------------
// A persistent CTA owns 1 / kGbinSplit of the padded positions: its slice of the split table (column
// layout) sits in shared memory.  A warp takes 16-sample tiles: the samples' g_ft rows are split into
// bf16 A fragments once per tile (K = L1) and stay in registers while the warp walks the CTA's
// positions eight at a time.
template <int KB>
__global__ void __launch_bounds__(kMmaGbinThreads, 1)
ft_bwd_gbin_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ wfrag,
                       const float *__restrict__ g_ft, float *__restrict__ gbin) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int part = blockIdx.x % kGbinSplit, cta = blockIdx.x / kGbinSplit, ctas = gridDim.x / kGbinSplit;
    // my share of the padded positions, in units of one bitmask word (4 column blocks of 8)
    const int words_per = ceil_div(s.NW, kGbinSplit);
    const int nb_begin = min(s.PP / 8, part * words_per * 4), nb_end = min(s.PP / 8, nb_begin + words_per * 4);
    const uint4 *sf = stage_fragments(smem_raw, wfrag + (size_t)nb_begin * 3 * (KB / 2) * 32,
                                      (uint32_t)(nb_end - nb_begin) * 3 * (KB / 2) * 512u) + lane;
    const int ntiles = ceil_div(s.B, 16), wpc = kMmaGbinThreads / 32;
    for (int tile = cta * wpc + warp; tile < ntiles; tile += ctas * wpc) {
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        uint32_t a[KB][3][4];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // a0: (row g, k 2t) a1: (row g+8, k 2t) a2: (row g, k 2t+8) a3: (row g+8, k 2t+8)
                const int r = (e & 1) ? r1 : r0, k0 = kb * 16 + t * 2 + (e >> 1) * 8;
                float2 v = make_float2(0.f, 0.f);
                if (r < s.B) v = __ldg(reinterpret_cast<const float2 *>(g_ft + (size_t)r * s.L1 + k0));
                uint32_t lo[3], hi[3];
                split3(v.x, lo);
                split3(v.y, hi);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) a[kb][sp][e] = lo[sp] | (hi[sp] << 16);
            }
        // One bitmask word (four column blocks of 8 positions) per pass.  Consecutive MMAs share their A
        // operand (the same g_ft fragment against the four blocks' table fragments): without operand reuse
        // HMMA runs at half rate on the register-file bandwidth (ncu: math-pipe throttle at 50 % tensor-pipe
        // utilisation with one block per pass).
        for (int nb = nb_begin; nb < nb_end; nb += 4) {
            const int wi = nb >> 2;
            const uint32_t word0 = r0 < s.B ? __ldg(bits_s + (size_t)r0 * s.NW + wi) : 0u;
            const uint32_t word1 = r1 < s.B ? __ldg(bits_s + (size_t)r1 * s.NW + wi) : 0u;
            float acc[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0f;
#pragma unroll
            for (int kbp = 0; kbp < KB / 2; ++kbp) {
                uint4 f[4][3];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int sw = 0; sw < 3; ++sw) f[u][sw] = sf[(((nb + u - nb_begin) * 3 + sw) * (KB / 2) + kbp) * 32];
                // term pairs (split of g, split of W) with i + j <= 2 (0-based): the rest is below 2^-27 relative
#pragma unroll
                for (int sa = 0; sa < 3; ++sa)
#pragma unroll
                    for (int sw = 0; sa + sw < 3; ++sw) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) mma_bf16(acc[u], a[2 * kbp][sa], f[u][sw].x, f[u][sw].y);
#pragma unroll
                        for (int u = 0; u < 4; ++u) mma_bf16(acc[u], a[2 * kbp + 1][sa], f[u][sw].z, f[u][sw].w);
                    }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int off = u * 8 + t * 2;
                float *o = gbin + (size_t)(nb + u) * 8 + t * 2;
                if (r0 < s.B) {
                    const uint32_t m = word0 >> off;
                    *reinterpret_cast<float2 *>(o + (size_t)r0 * s.PP) = make_float2((m & 1u) ? acc[u][0] : 0.0f, (m & 2u) ? acc[u][1] : 0.0f);
                }
                if (r1 < s.B) {
                    const uint32_t m = word1 >> off;
                    *reinterpret_cast<float2 *>(o + (size_t)r1 * s.PP) = make_float2((m & 1u) ? acc[u][2] : 0.0f, (m & 2u) ? acc[u][3] : 0.0f);
                }
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
static int format_rows(const nnue_shape &s, bool table, const float *src, int nrows, int n_kb, uint4 *out,
                       cudaStream_t st) {
    const long long n = 1LL * n_kb * (s.L1 / 16) * 32;
    if (table) mma_format_rows_kernel<true><<<(int)((n + 255) / 256), 256, 0, st>>>(s, src, nrows, n_kb, out);
    else mma_format_rows_kernel<false><<<(int)((n + 255) / 256), 256, 0, st>>>(s, src, nrows, n_kb, out);
    NNUE_CHECK_LAUNCH("mma_format_rows_kernel");
    return NNUE_OK;
}

int launch_ft_fwd_mma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *bias, float *out,
                      void *workspace, cudaStream_t st) {
    uint4 *wfrag = static_cast<uint4 *>(workspace);
    const int rc = format_rows(s, true, w, s.PP, s.PP / 16, wfrag, st);
    if (rc != NNUE_OK) return rc;
    const int groups = s.L1 / (16 * kMmaGP);
    const int ntiles = ceil_div(s.B, 16), wpc = kMmaFwdThreads / 32;
    int ctas = kNumSMs / groups;  // persistent CTAs per column group
    if (ctas > ceil_div(ntiles, wpc)) ctas = ceil_div(ntiles, wpc);
    const size_t smem = mma_fwd_smem(s);
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ft_fwd_mma_kernel<<<ctas * groups, kMmaFwdThreads, smem, st>>>(s, bits_s, wfrag, bias, out);
    NNUE_CHECK_LAUNCH("ft_fwd_mma_kernel");
    return NNUE_OK;
}

// partial[chunk][P][L1] and bias_partial[chunk][L1]; returns the number of chunks through *n_chunks
int launch_ft_bwd_dw_mma(const nnue_shape &s, const uint32_t *bits_s, const float *g_ft, uint4 *gfrag, float *partial,
                         float *bias_partial, int *n_chunks, cudaStream_t st) {
    const MmaPlan mp = plan_ft_mma(s);
    int rc = format_rows(s, false, g_ft, s.B, s.BW * 2, gfrag, st);
    if (rc != NNUE_OK) return rc;
    dim3 grid(ceil_div(s.NW, kMmaDwWarps), s.L1 / (16 * kMmaGP), mp.n_chunks);
    const size_t smem = 128 + (size_t)mp.chunk_blocks * 2 * 3 * kMmaGP * 512;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(ft_bwd_dw_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ft_bwd_dw_mma_kernel<<<grid, kMmaDwWarps * 32, smem, st>>>(s, bits_s, gfrag, partial, bias_partial, mp.chunk_blocks);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_mma_kernel");
    *n_chunks = mp.n_chunks;
    return NNUE_OK;
}

int launch_ft_bwd_gbin_mma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *g_ft, uint4 *wfrag,
                           float *gbin, cudaStream_t st) {
    const long long n = 1LL * (s.PP / 8) * (s.L1 / 32) * 32;
    mma_format_cols_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, wfrag);
    NNUE_CHECK_LAUNCH("mma_format_cols_kernel");
    const int ntiles = ceil_div(s.B, 16), wpc = kMmaGbinThreads / 32;
    int ctas = kNumSMs / kGbinSplit;  // persistent CTAs per position slice
    if (ctas > ceil_div(ntiles, wpc)) ctas = ceil_div(ntiles, wpc);
    const size_t smem = mma_gbin_smem(s);
    auto k = s.L1 == 64 ? ft_bwd_gbin_mma_kernel<4> : ft_bwd_gbin_mma_kernel<2>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<ctas * kGbinSplit, kMmaGbinThreads, smem, st>>>(s, bits_s, wfrag, g_ft, gbin);
    NNUE_CHECK_LAUNCH("ft_bwd_gbin_mma_kernel");
    return NNUE_OK;
}

}  // namespace nnue
