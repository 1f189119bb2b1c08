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
// Operands are pre-formatted into MMA fragment order by small kernels (a lane then loads its
// B fragments with one coalesced 16-byte load); the bitmask is expanded to bf16 A fragments in
// registers (two integer instructions per register, amortised over 24 MMAs).
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
// out[((kb * 3 + s) * (NB / 2) + nbp) * 32 + lane] = uint4 {b0, b1 of column block 2 nbp, b0, b1 of 2 nbp + 1}.
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
        out[((size_t)(kb * 3 + sp) * NBP + nbp) * 32 + lane] = make_uint4(packed[sp][0], packed[sp][1], packed[sp][2], packed[sp][3]);
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

// ---- forward: out[b] = bias + bits[b] . Wc ------------------------------------------------------------------
// A warp owns 32 samples (two m16 tiles) and NB of the L1 / 8 column blocks (the warps of a CTA share
// the samples and split the columns, so there are enough warps to fill the machine); K runs over the
// padded positions, one bitmask word (two k16 steps) at a time.
template <int NB>
__global__ void __launch_bounds__(kMmaThreads)
ft_fwd_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ wfrag,
                  const float *__restrict__ bias, float *__restrict__ out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NBP_ALL = s.L1 / 16;                       // column-block pairs of the whole row
    const int col_groups = NBP_ALL / (NB / 2);           // warps that share one sample tile
    const int wid = blockIdx.x * (kMmaThreads / 32) + (threadIdx.x >> 5);
    const int b_base = (wid / col_groups) * 32;
    const int nbp0 = (wid % col_groups) * (NB / 2);      // my first column-block pair
    if (b_base >= s.B) return;
    float acc[2][NB][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[mt][nb][0] = acc[mt][nb][1] = acc[mt][nb][2] = acc[mt][nb][3] = 0.0f;
    int rows[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) rows[mt][h] = b_base + mt * 16 + g + 8 * h;

    // operands of one bitmask word (two k16 steps): my rows' words and my B fragments; the next word's
    // operands are fetched before this word's MMAs are issued, so the L1 / L2 latency hides under them
    struct Operands {
        uint32_t word[2][2];
        uint4 f[2][3][NB / 2];
    };
    auto fetch = [&](Operands &o, int w) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                o.word[mt][h] = rows[mt][h] < s.B ? __ldg(bits_s + (size_t)rows[mt][h] * s.NW + w) : 0u;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int nbp = 0; nbp < NB / 2; ++nbp)
                    o.f[half][sp][nbp] = __ldg(wfrag + ((size_t)((2 * w + half) * 3 + sp) * NBP_ALL + nbp0 + nbp) * 32 + lane);
    };
    Operands cur, nxt;
    fetch(cur, 0);
    for (int w = 0; w < s.NW; ++w) {
        if (w + 1 < s.NW) fetch(nxt, w + 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int sh = half * 16 + t * 2;
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                a[mt][0] = bits2_to_bf16x2(cur.word[mt][0] >> sh);
                a[mt][1] = bits2_to_bf16x2(cur.word[mt][1] >> sh);
                a[mt][2] = bits2_to_bf16x2(cur.word[mt][0] >> (sh + 8));
                a[mt][3] = bits2_to_bf16x2(cur.word[mt][1] >> (sh + 8));
            }
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int nbp = 0; nbp < NB / 2; ++nbp) {
                    const uint4 f = cur.f[half][sp][nbp];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_bf16(acc[mt][2 * nbp], a[mt], f.x, f.y);
                        mma_bf16(acc[mt][2 * nbp + 1], a[mt], f.z, f.w);
                    }
                }
        }
        cur = nxt;
    }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const int col = (nbp0 * 2 + nb) * 8 + t * 2;
        const float2 bv = __ldg(reinterpret_cast<const float2 *>(bias + col));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (rows[mt][h] < s.B)
                    *reinterpret_cast<float2 *>(out + (size_t)rows[mt][h] * s.L1 + col) =
                        make_float2(bv.x + acc[mt][nb][2 * h], bv.y + acc[mt][nb][2 * h + 1]);
    }
}

// ---- weight gradient: dW[p] = sum_b bits[b, p] g_ft[b] --------------------------------------------------------
// A warp owns one bitmask word (32 positions = two m16 tiles) and one K-chunk of samples; the 32 x 32
// bit block of (sample, position) is transposed in registers so that k runs over samples.  Output:
// partial[chunk][p][L1] for the fold kernels of ft.cu (they also resolve the aliasing onto row F-1).
// Warp 0 of the first word block also multiplies an all-ones tile: its first row is the bias gradient.
template <int NB>
__global__ void __launch_bounds__(kMmaThreads)
ft_bwd_dw_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ gfrag,
                     float *__restrict__ partial, float *__restrict__ bias_partial, int chunk_blocks) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NBP_ALL = s.L1 / 16, col_groups = NBP_ALL / (NB / 2);
    const int wid = blockIdx.x * (kMmaThreads / 32) + (threadIdx.x >> 5);
    const int w = wid / col_groups;                      // my bitmask word
    const int nbp0 = (wid % col_groups) * (NB / 2);      // my first column-block pair
    const int chunk = blockIdx.y;
    const bool active = w < s.NW;
    const bool do_bias = w == 0;
    float acc[2][NB][4], accb[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) acc[mt][nb][0] = acc[mt][nb][1] = acc[mt][nb][2] = acc[mt][nb][3] = 0.0f;
        accb[nb][0] = accb[nb][1] = accb[nb][2] = accb[nb][3] = 0.0f;
    }
    const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};
    const int sb_begin = chunk * chunk_blocks, sb_end = min(s.BW, sb_begin + chunk_blocks);  // 32-sample blocks
    struct Operands {
        uint32_t x;            // my sample's word (lane = sample within the block)
        uint4 f[2][3][NB / 2];
    };
    auto fetch = [&](Operands &o, int sb) {
        const int b = sb * 32 + lane;
        o.x = (active && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + w) : 0u;
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int nbp = 0; nbp < NB / 2; ++nbp)
                    o.f[half][sp][nbp] = __ldg(gfrag + ((size_t)((2 * sb + half) * 3 + sp) * NBP_ALL + nbp0 + nbp) * 32 + lane);
    };
    Operands cur, nxt;
    if (sb_begin < sb_end) fetch(cur, sb_begin);
    for (int sb = sb_begin; sb < sb_end; ++sb) {
        if (sb + 1 < sb_end) fetch(nxt, sb + 1);
        const uint32_t tr = warp_bit_transpose(cur.x, lane);  // lane L: the 32 samples of position w*32 + L
        uint32_t word[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) word[mt][h] = __shfl_sync(kFull, tr, mt * 16 + g + 8 * h);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int sh = half * 16 + t * 2;
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                a[mt][0] = bits2_to_bf16x2(word[mt][0] >> sh);
                a[mt][1] = bits2_to_bf16x2(word[mt][1] >> sh);
                a[mt][2] = bits2_to_bf16x2(word[mt][0] >> (sh + 8));
                a[mt][3] = bits2_to_bf16x2(word[mt][1] >> (sh + 8));
            }
#pragma unroll
            for (int sp = 0; sp < 3; ++sp)
#pragma unroll
                for (int nbp = 0; nbp < NB / 2; ++nbp) {
                    const uint4 f = cur.f[half][sp][nbp];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_bf16(acc[mt][2 * nbp], a[mt], f.x, f.y);
                        mma_bf16(acc[mt][2 * nbp + 1], a[mt], f.z, f.w);
                    }
                    if (do_bias) {  // warp-uniform
                        mma_bf16(accb[2 * nbp], ones, f.x, f.y);
                        mma_bf16(accb[2 * nbp + 1], ones, f.z, f.w);
                    }
                }
        }
        cur = nxt;
    }
    if (active) {
        const int cells = s.Gh * s.Gw, c = w / s.CW, cell0 = (w % s.CW) * 32;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cell = cell0 + mt * 16 + g + 8 * h;
                if (cell >= cells) continue;
                float *row = partial + ((size_t)chunk * s.P + (size_t)c * cells + cell) * s.L1;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                    *reinterpret_cast<float2 *>(row + (nbp0 * 2 + nb) * 8 + t * 2) =
                        make_float2(acc[mt][nb][2 * h], acc[mt][nb][2 * h + 1]);
            }
    }
    if (do_bias && g == 0)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
            *reinterpret_cast<float2 *>(bias_partial + (size_t)chunk * s.L1 + (nbp0 * 2 + nb) * 8 + t * 2) =
                make_float2(accb[nb][0], accb[nb][1]);
}

// ---- value gradient: gbin[b, pp] = bit(b, pp) ? <Wc[pp], g_ft[b]> : 0 ------------------------------------------
// A warp owns 32 samples and 1 / kGbinSplit of the padded positions; the samples' g_ft rows are split
// into bf16 A fragments once (K = L1) and stay in registers while the warp walks its positions eight
// at a time.
constexpr int kGbinSplit = 8;
template <int KB>
__global__ void __launch_bounds__(kMmaThreads)
ft_bwd_gbin_mma_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const uint4 *__restrict__ wfrag,
                       const float *__restrict__ g_ft, float *__restrict__ gbin) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wid = blockIdx.x * (kMmaThreads / 32) + (threadIdx.x >> 5);
    const int b_base = (wid / kGbinSplit) * 32;
    if (b_base >= s.B) return;
    // my share of the padded positions, in units of one bitmask word (4 column blocks of 8)
    const int words_per = ceil_div(s.NW, kGbinSplit);
    const int nb_begin = (wid % kGbinSplit) * words_per * 4, nb_end = min(s.PP / 8, nb_begin + words_per * 4);
    int rows[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) rows[mt][h] = b_base + mt * 16 + g + 8 * h;
    uint32_t a[2][KB][3][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // a0: (row g, k 2t) a1: (row g+8, k 2t) a2: (row g, k 2t+8) a3: (row g+8, k 2t+8)
                const int r = rows[mt][e & 1], k0 = kb * 16 + t * 2 + (e >> 1) * 8;
                float2 v = make_float2(0.f, 0.f);
                if (r < s.B) v = __ldg(reinterpret_cast<const float2 *>(g_ft + (size_t)r * s.L1 + k0));
                uint32_t lo[3], hi[3];
                split3(v.x, lo);
                split3(v.y, hi);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) a[mt][kb][sp][e] = lo[sp] | (hi[sp] << 16);
            }
    uint32_t word[2][2] = {{0u, 0u}, {0u, 0u}};
    uint4 fnext[3][KB / 2];
#pragma unroll
    for (int sp = 0; sp < 3; ++sp)
#pragma unroll
        for (int kbp = 0; kbp < KB / 2; ++kbp)
            fnext[sp][kbp] = nb_begin < nb_end ? __ldg(wfrag + ((size_t)(nb_begin * 3 + sp) * (KB / 2) + kbp) * 32 + lane)
                                               : make_uint4(0u, 0u, 0u, 0u);
    for (int nb = nb_begin; nb < nb_end; ++nb) {
        if ((nb & 3) == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    word[mt][h] = rows[mt][h] < s.B ? __ldg(bits_s + (size_t)rows[mt][h] * s.NW + (nb >> 2)) : 0u;
        }
        float acc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.0f;
        uint4 f[3][KB / 2];
#pragma unroll
        for (int sp = 0; sp < 3; ++sp)
#pragma unroll
            for (int kbp = 0; kbp < KB / 2; ++kbp) {
                f[sp][kbp] = fnext[sp][kbp];
                if (nb + 1 < nb_end)  // the next block's fragments fly while this block's MMAs issue
                    fnext[sp][kbp] = __ldg(wfrag + ((size_t)((nb + 1) * 3 + sp) * (KB / 2) + kbp) * 32 + lane);
            }
        // term pairs (split of g, split of W) with i + j <= 2 (0-based): the rest is below 2^-27 relative
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int sa = 0; sa < 3; ++sa)
#pragma unroll
                for (int sw = 0; sw + sa < 3; ++sw) {
                    const uint4 ff = f[sw][kb >> 1];
                    const uint32_t b0 = (kb & 1) ? ff.z : ff.x, b1 = (kb & 1) ? ff.w : ff.y;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_bf16(acc[mt], a[mt][kb][sa], b0, b1);
                }
        const int off = (nb & 3) * 8 + t * 2;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (rows[mt][h] < s.B) {
                    const uint32_t m = word[mt][h] >> off;
                    *reinterpret_cast<float2 *>(gbin + (size_t)rows[mt][h] * s.PP + nb * 8 + t * 2) =
                        make_float2((m & 1u) ? acc[mt][2 * h] : 0.0f, (m & 2u) ? acc[mt][2 * h + 1] : 0.0f);
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
    // 2 column blocks (16 columns) per warp: L1 / 16 warps share a 32-sample tile
    const int grid = ceil_div(ceil_div(s.B, 32) * (s.L1 / 16), kMmaThreads / 32);
    ft_fwd_mma_kernel<2><<<grid, kMmaThreads, 0, st>>>(s, bits_s, wfrag, bias, out);
    NNUE_CHECK_LAUNCH("ft_fwd_mma_kernel");
    return NNUE_OK;
}

// partial[chunk][P][L1] and bias_partial[chunk][L1]; returns the number of chunks through *n_chunks
int launch_ft_bwd_dw_mma(const nnue_shape &s, const uint32_t *bits_s, const float *g_ft, uint4 *gfrag, float *partial,
                         float *bias_partial, int *n_chunks, cudaStream_t st) {
    const MmaPlan mp = plan_ft_mma(s);
    int rc = format_rows(s, false, g_ft, s.B, s.BW * 2, gfrag, st);
    if (rc != NNUE_OK) return rc;
    // 2 column blocks (16 columns) per warp: L1 / 16 warps share one bitmask word
    dim3 grid(ceil_div(s.NW * (s.L1 / 16), kMmaThreads / 32), mp.n_chunks);
    ft_bwd_dw_mma_kernel<2><<<grid, kMmaThreads, 0, st>>>(s, bits_s, gfrag, partial, bias_partial, mp.chunk_blocks);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_mma_kernel");
    *n_chunks = mp.n_chunks;
    return NNUE_OK;
}

int launch_ft_bwd_gbin_mma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *g_ft, uint4 *wfrag,
                           float *gbin, cudaStream_t st) {
    const long long n = 1LL * (s.PP / 8) * (s.L1 / 32) * 32;
    mma_format_cols_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s, w, wfrag);
    NNUE_CHECK_LAUNCH("mma_format_cols_kernel");
    const int grid = ceil_div(ceil_div(s.B, 32) * kGbinSplit, kMmaThreads / 32);
    if (s.L1 == 64) ft_bwd_gbin_mma_kernel<4><<<grid, kMmaThreads, 0, st>>>(s, bits_s, wfrag, g_ft, gbin);
    else ft_bwd_gbin_mma_kernel<2><<<grid, kMmaThreads, 0, st>>>(s, bits_s, wfrag, g_ft, gbin);
    NNUE_CHECK_LAUNCH("ft_bwd_gbin_mma_kernel");
    return NNUE_OK;
}

}  // namespace nnue
