// ft.cu -- the sparse feature transformer: forward gather-accumulate, weight gradient as a
// segment reduction over the transposed bitmask, and the value gradient (row . grad dot).
//
// Column geometry (plan.cuh ColPlan): a table row is read as float4s by LPR lanes (one
// coalesced 16*LPR-byte segment); with LPR < 32 a warp splits into NG = 32/LPR lane groups that
// take different active features and are combined with __shfl_xor at the end.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kFtThreads = 256;

// The set bit of rank g (0-based) among the NG lowest set bits of `word`, or -1 when the word
// has fewer; the NG lowest set bits are then stripped.  __ffs(0) == 0 gives the -1 for free.
template <int NG>
__device__ __forceinline__ int take_bit(unsigned &word, int g) {
    int k = -1;
#pragma unroll
    for (int t = 0; t < NG; ++t) {
        const int kt = __ffs(word) - 1;
        if (t == g) k = kt;
        word &= word - 1;
    }
    return k;
}

// Sum a float4 across the NG lane groups of a warp (lanes l, l+LPR, l+2*LPR, ...).
template <int LPR>
__device__ __forceinline__ float4 group_sum(float4 v) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(kFull, v.x, o);
        v.y += __shfl_xor_sync(kFull, v.y, o);
        v.z += __shfl_xor_sync(kFull, v.z, o);
        v.w += __shfl_xor_sync(kFull, v.w, o);
    }
    return v;
}

// ---- forward on the bitmask ----------------------------------------------------------------
// unit = (sample, column chunk); one warp per unit.  STAGED: the whole table sits in shared
// memory, brought in by bulk TMA copies once per (persistent) CTA.
template <int LPR, bool STAGED>
__global__ void __launch_bounds__(kFtThreads)
ft_fwd_bits_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                   const float *__restrict__ bias, float *__restrict__ out, int nchunks) {
    constexpr int NG = 32 / LPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float *table = w;
    if (STAGED) {
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
        float *stab = reinterpret_cast<float *>(smem_raw + 128);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        tma_stage(stab, w, (uint32_t)((size_t)s.F * s.L1 * sizeof(float)), bar, 0);
        table = stab;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int g = lane / LPR, li = lane % LPR;
    const int cells = s.Gh * s.Gw;
    const long long units = 1LL * s.B * nchunks;
    for (long long u = 1LL * blockIdx.x * wpb + warp; u < units; u += 1LL * gridDim.x * wpb) {
        const int b = (int)(u / nchunks), ch = (int)(u % nchunks);
        const int col = ch * (4 * LPR) + li * 4;
        const float *tcol = table + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w0 = 0; w0 < s.NW; w0 += 32) {
            const unsigned mine = (w0 + lane < s.NW) ? __ldg(bits_s + (size_t)b * s.NW + w0 + lane) : 0u;
            unsigned nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                unsigned word = __shfl_sync(kFull, mine, jw);
                const int widx = w0 + jw;
                const int base = (widx / s.CW) * cells + (widx % s.CW) * 32;
                while (word) {
                    const int k = take_bit<NG>(word, g);
                    if (k >= 0) {
                        const int row = min(base + k, s.F - 1);  // clamp of nnue.py:701
                        const float4 v = STAGED ? *reinterpret_cast<const float4 *>(tcol + (size_t)row * s.L1)
                                                : __ldg(reinterpret_cast<const float4 *>(tcol + (size_t)row * s.L1));
                        acc = f4_add(acc, v);
                    }
                }
            }
        }
        acc = group_sum<LPR>(acc);
        if (g == 0) {
            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + col));
            *reinterpret_cast<float4 *>(out + (size_t)b * s.L1 + col) = f4_add(bv, acc);
        }
    }
}

// any L1: one warp per sample, lanes stride the columns, 32 columns per pass
__global__ void __launch_bounds__(kFtThreads)
ft_fwd_bits_generic_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                           const float *__restrict__ bias, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= s.B) return;
    const int cells = s.Gh * s.Gw;
    for (int c0 = 0; c0 < s.L1; c0 += 32) {
        const int col = c0 + lane;
        float acc = 0.0f;
        for (int widx = 0; widx < s.NW; ++widx) {
            unsigned word = __ldg(bits_s + (size_t)b * s.NW + widx);
            const int base = (widx / s.CW) * cells + (widx % s.CW) * 32;
            while (word) {
                const int k = __ffs(word) - 1;
                word &= word - 1;
                const int row = min(base + k, s.F - 1);
                if (col < s.L1) acc += __ldg(w + (size_t)row * s.L1 + col);
            }
        }
        if (col < s.L1) out[(size_t)b * s.L1 + col] = bias[col] + acc;
    }
}

// ---- forward / backward on explicit (idx, val) lists: model.input(idx, val) ------------------
__global__ void __launch_bounds__(kFtThreads)
ft_fwd_indexed_kernel(int B, int K, int F, int L1, const int64_t *__restrict__ idx, const float *__restrict__ val,
                      const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    for (int c0 = 0; c0 < L1; c0 += 32) {
        const int col = c0 + lane;
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) {
            const long long f = idx[(size_t)b * K + k];
            if (f < 0) continue;
            const int row = f > F - 1 ? F - 1 : (int)f;
            const float v = val[(size_t)b * K + k];
            if (col < L1) acc = fmaf(__ldg(w + (size_t)row * L1 + col), v, acc);
        }
        if (col < L1) out[(size_t)b * L1 + col] = bias[col] + acc;
    }
}

// g_val[b,k] = <W[row], g_out[b]> for idx >= 0, else 0
__global__ void __launch_bounds__(kFtThreads)
ft_bwd_indexed_dval_kernel(int B, int K, int F, int L1, const int64_t *__restrict__ idx,
                           const float *__restrict__ w, const float *__restrict__ g_out,
                           float *__restrict__ g_val) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    for (int k = 0; k < K; ++k) {
        const long long f = idx[(size_t)b * K + k];
        float d = 0.0f;
        if (f >= 0) {
            const int row = f > F - 1 ? F - 1 : (int)f;
            for (int col = lane; col < L1; col += 32)
                d = fmaf(__ldg(w + (size_t)row * L1 + col), __ldg(g_out + (size_t)b * L1 + col), d);
            d = warp_sum(d);
        }
        if (lane == 0) g_val[(size_t)b * K + k] = d;
    }
}

// Sorted segment reduction: triples (row, sample, value) arrive sorted by row; one warp owns a
// table row, finds its segment by binary search and sums value * g_out[sample] in segment order.
__global__ void __launch_bounds__(kFtThreads)
ft_bwd_indexed_dw_kernel(int F, int L1, int n_pairs, const int32_t *__restrict__ row, const int32_t *__restrict__ sample,
                         const float *__restrict__ pval, const float *__restrict__ g_out, float *__restrict__ g_w) {
    const int lane = threadIdx.x & 31;
    const int r = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= F) return;
    int lo = 0, hi = n_pairs;  // first index with row[i] >= r
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(row + mid) < r) lo = mid + 1; else hi = mid;
    }
    const int start = lo;
    hi = n_pairs;  // first index with row[i] > r
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(row + mid) <= r) lo = mid + 1; else hi = mid;
    }
    const int end = lo;
    for (int c0 = 0; c0 < L1; c0 += 32) {
        const int col = c0 + lane;
        float acc = 0.0f;
        for (int i = start; i < end; ++i)
            if (col < L1) acc = fmaf(__ldg(pval + i), __ldg(g_out + (size_t)__ldg(sample + i) * L1 + col), acc);
        if (col < L1) g_w[(size_t)r * L1 + col] = acc;
    }
}

// ---- column sums (bias gradient): partial[chunk][col] over kColsumRows rows, then a fold ----
__global__ void colsum_partial_kernel(int rows, int cols, const float *__restrict__ x, float *__restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const int r0 = blockIdx.y * kColsumRows, r1 = min(rows, r0 + kColsumRows);
    float acc = 0.0f;
    for (int r = r0; r < r1; ++r) acc += __ldg(x + (size_t)r * cols + col);
    partial[(size_t)blockIdx.y * cols + col] = acc;
}
// one thread per column over all rows (the indexed interface only sees small batches)
__global__ void colsum_full_kernel(int rows, int cols, const float *__restrict__ x, float *__restrict__ out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    float acc = 0.0f;
    for (int r = 0; r < rows; ++r) acc += __ldg(x + (size_t)r * cols + col);
    out[col] = acc;
}
__global__ void fold_partials_kernel(int n, int nparts, const float *__restrict__ partial, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.0f;
    for (int k = 0; k < nparts; ++k) acc += partial[(size_t)k * n + i];
    out[i] = acc;
}

// ---- weight gradient: segment reduction over the transposed bitmask --------------------------
// CTA = (position chunk of kDwPPC padded positions, column chunk, tile group).  For each sample
// tile of its group the CTA stages g_ft[tile rows][column chunk] in shared memory with bulk TMA
// copies; each warp then walks the set bits (= the sorted sample list) of its kDwPPW positions
// and adds the staged rows into register accumulators.  Output: partial[group][p][L1] (or g_w
// directly when there is a single group and no clamp aliasing).
template <int LPR>
__global__ void __launch_bounds__(kDwWarps * 32)
ft_bwd_dw_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_t, const float *__restrict__ g_ft,
                 float *__restrict__ dst, int TS, int tpg, int ntiles, int direct) {
    constexpr int NG = 32 / LPR;
    constexpr int CC = 4 * LPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *tile = reinterpret_cast<float *>(smem_raw + 64);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPR, li = lane % LPR;
    const int ch = blockIdx.y, group = blockIdx.z;
    const int col0 = ch * CC;
    const int cells = s.Gh * s.Gw;
    const int wpt = TS / 32;  // bitmask words per tile

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    float4 acc[kDwPPW];
#pragma unroll
    for (int q = 0; q < kDwPPW; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pp0 = blockIdx.x * kDwPPC + warp * kDwPPW;

    uint32_t parity = 0;
    const int t_begin = group * tpg, t_end = min(ntiles, t_begin + tpg);
    for (int t = t_begin; t < t_end; ++t) {
        const int b0 = t * TS;
        const int rows = min(TS, s.B - b0);
        // stage g_ft[b0 .. b0+rows)[col0 .. col0+CC) -> tile[rows][CC]
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)rows * CC * 4);
            if (CC == s.L1) {
                constexpr uint32_t kChunk = 32768;
                const uint32_t bytes = (uint32_t)rows * CC * 4;
                for (uint32_t off = 0; off < bytes; off += kChunk)
                    tma_bulk_g2s(reinterpret_cast<char *>(tile) + off,
                                 reinterpret_cast<const char *>(g_ft + (size_t)b0 * s.L1) + off,
                                 bytes - off < kChunk ? bytes - off : kChunk, bar);
            } else {
                for (int r = 0; r < rows; ++r)
                    tma_bulk_g2s(tile + (size_t)r * CC, g_ft + (size_t)(b0 + r) * s.L1 + col0, CC * 4, bar);
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        const float *tcol = tile + li * 4;
#pragma unroll
        for (int q = 0; q < kDwPPW; ++q) {
            const int pp = pp0 + q;
            if (pp >= s.PP || (pp & 31) + ((pp >> 5) % s.CW) * 32 >= cells) continue;  // warp-uniform
            const uint32_t *wrow = bits_t + (size_t)pp * s.BW + (size_t)t * wpt;
            const int nw = min(wpt, s.BW - t * wpt);
            const unsigned mine = lane < nw ? __ldg(wrow + lane) : 0u;  // wpt <= 8 words
            unsigned nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                unsigned word = __shfl_sync(kFull, mine, jw);
                while (word) {
                    const int k = take_bit<NG>(word, g);
                    if (k >= 0) acc[q] = f4_add(acc[q], *reinterpret_cast<const float4 *>(tcol + (size_t)(jw * 32 + k) * CC));
                }
            }
        }
        __syncthreads();  // everyone done with the tile before the next bulk copy overwrites it
    }
#pragma unroll
    for (int q = 0; q < kDwPPW; ++q) {
        const int pp = pp0 + q;
        if (pp >= s.PP) continue;
        const int cell = (pp & 31) + ((pp >> 5) % s.CW) * 32;
        if (cell >= cells) continue;
        const float4 v = group_sum<LPR>(acc[q]);
        if (g == 0) {
            const int p = ((pp >> 5) / s.CW) * cells + cell;  // flat CHW position
            float *o = direct ? dst + (size_t)p * s.L1 : dst + ((size_t)group * s.P + p) * s.L1;
            *reinterpret_cast<float4 *>(o + col0 + li * 4) = v;
        }
    }
}

// g_w[r] = sum over groups and over positions p with min(p, F-1) == r of partial[group][p]
__global__ void ft_bwd_dw_fold_kernel(const nnue_shape s, int ngroups, const float *__restrict__ partial,
                                      float *__restrict__ g_w) {
    const int v4 = s.L1 / 4;
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * s.F * v4) return;
    const int r = (int)(i / v4), c4 = (int)(i % v4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p_begin = r, p_end = (r == s.F - 1) ? s.P : min(r + 1, s.P);
    for (int p = p_begin; p < p_end; ++p)
        for (int gI = 0; gI < ngroups; ++gI)
            acc = f4_add(acc, __ldg(reinterpret_cast<const float4 *>(partial + ((size_t)gI * s.P + p) * s.L1) + c4));
    reinterpret_cast<float4 *>(g_w + (size_t)r * s.L1)[c4] = acc;
}

// any L1: one warp per table row, walks every position that maps to the row and every sample bit
__global__ void __launch_bounds__(kFtThreads)
ft_bwd_dw_generic_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_t, const float *__restrict__ g_ft,
                         float *__restrict__ g_w) {
    const int lane = threadIdx.x & 31;
    const int r = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= s.F) return;
    const int cells = s.Gh * s.Gw;
    const int p_end = (r == s.F - 1) ? s.P : min(r + 1, s.P);
    for (int c0 = 0; c0 < s.L1; c0 += 32) {
        const int col = c0 + lane;
        float acc = 0.0f;
        for (int p = r; p < p_end; ++p) {
            const int c = p / cells, cell = p % cells;
            const size_t pp = ((size_t)c * s.CW + cell / 32) * 32 + cell % 32;
            for (int bw = 0; bw < s.BW; ++bw) {
                unsigned word = __ldg(bits_t + pp * s.BW + bw);
                while (word) {
                    const int k = __ffs(word) - 1;
                    word &= word - 1;
                    if (col < s.L1) acc += __ldg(g_ft + (size_t)(bw * 32 + k) * s.L1 + col);
                }
            }
        }
        if (col < s.L1) g_w[(size_t)r * s.L1 + col] = acc;
    }
}

// ---- value gradient at active positions --------------------------------------------------------
// one warp per sample; each lane group takes an active position and all its lanes share the dot.
template <int LPR>
__global__ void __launch_bounds__(kFtThreads)
ft_bwd_dval_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                   const float *__restrict__ g_ft, float *__restrict__ dval, int nchunks) {
    constexpr int NG = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int g = lane / LPR, li = lane % LPR;
    const int cells = s.Gh * s.Gw;
    for (int b = blockIdx.x * wpb + warp; b < s.B; b += gridDim.x * wpb) {
        const float *grow = g_ft + (size_t)b * s.L1 + li * 4;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nchunks == 1) g0 = __ldg(reinterpret_cast<const float4 *>(grow));
        for (int w0 = 0; w0 < s.NW; w0 += 32) {
            const unsigned mine = (w0 + lane < s.NW) ? __ldg(bits_s + (size_t)b * s.NW + w0 + lane) : 0u;
            unsigned nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                unsigned word = __shfl_sync(kFull, mine, jw);
                const int widx = w0 + jw;
                const int base = (widx / s.CW) * cells + (widx % s.CW) * 32;
                while (word) {  // warp-uniform trip count
                    const int k = take_bit<NG>(word, g);
                    float d = 0.0f;
                    if (k >= 0) {
                        const int row = min(base + k, s.F - 1);
                        const float *wrow = w + (size_t)row * s.L1 + li * 4;
                        if (nchunks == 1) {
                            const float4 v = __ldg(reinterpret_cast<const float4 *>(wrow));
                            d = fmaf(v.x, g0.x, fmaf(v.y, g0.y, fmaf(v.z, g0.z, v.w * g0.w)));
                        } else {
                            for (int ch = 0; ch < nchunks; ++ch) {
                                const float4 v = __ldg(reinterpret_cast<const float4 *>(wrow + ch * 128));
                                const float4 gv = __ldg(reinterpret_cast<const float4 *>(grow + ch * 128));
                                d = fmaf(v.x, gv.x, fmaf(v.y, gv.y, fmaf(v.z, gv.z, fmaf(v.w, gv.w, d))));
                            }
                        }
                    }
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
                    if (k >= 0 && li == 0) dval[(size_t)b * s.PP + (size_t)widx * 32 + k] = d;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kFtThreads)
ft_bwd_dval_generic_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                           const float *__restrict__ g_ft, float *__restrict__ dval) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= s.B) return;
    const int cells = s.Gh * s.Gw;
    for (int widx = 0; widx < s.NW; ++widx) {
        unsigned word = __ldg(bits_s + (size_t)b * s.NW + widx);
        const int base = (widx / s.CW) * cells + (widx % s.CW) * 32;
        while (word) {
            const int k = __ffs(word) - 1;
            word &= word - 1;
            const int row = min(base + k, s.F - 1);
            float d = 0.0f;
            for (int col = lane; col < s.L1; col += 32)
                d = fmaf(__ldg(w + (size_t)row * s.L1 + col), __ldg(g_ft + (size_t)b * s.L1 + col), d);
            d = warp_sum(d);
            if (lane == 0) dval[(size_t)b * s.PP + (size_t)widx * 32 + k] = d;
        }
    }
}

static int max_dyn_smem() {
    static int v = -1;
    if (v < 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            v = 48 * 1024;
    }
    return v;
}

template <int LPR>
static int launch_ft_fwd(const nnue_shape &s, const uint32_t *bits, const float *w, const float *b, float *out,
                         int nchunks, cudaStream_t st) {
    const size_t table_bytes = (size_t)s.F * s.L1 * 4;
    const size_t staged_smem = table_bytes + 128;
    const int wpb = kFtThreads / 32;
    const long long units = 1LL * s.B * nchunks;
    // stage the table when it fits one CTA's shared memory and there is enough work to amortise it
    const int mode = get_option(kOptFtFwdStaging);  // 0 never, 1 auto, 2 whenever it fits
    const bool fits = nchunks == 1 && staged_smem <= (size_t)max_dyn_smem() && table_bytes % 16 == 0 &&
                      table_bytes < (1u << 20);
    const bool staged = fits && (mode == 2 || (mode == 1 && units >= 8LL * kNumSMs));
    if (staged) {
        auto k = ft_fwd_bits_kernel<LPR, true>;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
        k<<<kNumSMs, kFtThreads, staged_smem, st>>>(s, bits, w, b, out, nchunks);
    } else {
        long long grid = (units + wpb - 1) / wpb;
        if (grid > 64LL * kNumSMs) grid = 64LL * kNumSMs;
        ft_fwd_bits_kernel<LPR, false><<<(int)grid, kFtThreads, 0, st>>>(s, bits, w, b, out, nchunks);
    }
    NNUE_CHECK_LAUNCH("ft_fwd_bits_kernel");
    return NNUE_OK;
}

template <int LPR>
static int launch_ft_bwd_dw(const nnue_shape &s, const DwPlan &d, const uint32_t *bits_t, const float *g_ft,
                            float *dst, cudaStream_t st) {
    auto k = ft_bwd_dw_kernel<LPR>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem));
    dim3 grid(d.pchunks, d.col.nchunks, d.ngroups);
    k<<<grid, kDwWarps * 32, d.smem, st>>>(s, bits_t, g_ft, dst, d.TS, d.tpg, d.ntiles, d.direct ? 1 : 0);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_kernel");
    return NNUE_OK;
}

template <int LPR>
static int launch_ft_bwd_dval(const nnue_shape &s, const uint32_t *bits, const float *w, const float *g_ft,
                              float *dval, int nchunks, cudaStream_t st) {
    const int wpb = kFtThreads / 32;
    long long grid = (s.B + wpb - 1) / wpb;
    if (grid > 64LL * kNumSMs) grid = 64LL * kNumSMs;
    ft_bwd_dval_kernel<LPR><<<(int)grid, kFtThreads, 0, st>>>(s, bits, w, g_ft, dval, nchunks);
    NNUE_CHECK_LAUNCH("ft_bwd_dval_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

#define NNUE_DISPATCH_LPR(lpr, CALL)              \
    switch (lpr) {                                \
        case 4: { constexpr int LPR = 4; CALL; } break;   \
        case 8: { constexpr int LPR = 8; CALL; } break;   \
        case 16: { constexpr int LPR = 16; CALL; } break; \
        default: { constexpr int LPR = 32; CALL; } break; \
    }

extern "C" {

int nnue_ft_fwd(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *ft_b_d,
                float *ft_out_d, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !ft_b_d || !ft_out_d) return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const ColPlan cp = col_plan(s->L1);
    if (!cp.LPR) {
        ft_fwd_bits_generic_kernel<<<ceil_div(s->B, kFtThreads / 32), kFtThreads, 0, st>>>(*s, bits_s_d, ft_w_d, ft_b_d,
                                                                                          ft_out_d);
        NNUE_CHECK_LAUNCH("ft_fwd_bits_generic_kernel");
        return NNUE_OK;
    }
    int rc = NNUE_OK;
    NNUE_DISPATCH_LPR(cp.LPR, rc = launch_ft_fwd<LPR>(*s, bits_s_d, ft_w_d, ft_b_d, ft_out_d, cp.nchunks, st));
    return rc;
}

int nnue_ft_fwd_indexed(int B, int K, int F, int L1, const int64_t *idx_d, const float *val_d, const float *ft_w_d,
                        const float *ft_b_d, float *ft_out_d, void *stream) {
    if (B < 1 || K < 1 || F < 1 || L1 < 1 || !idx_d || !val_d || !ft_w_d || !ft_b_d || !ft_out_d)
        return NNUE_ERR_INVALID_ARG;
    ft_fwd_indexed_kernel<<<ceil_div(B, kFtThreads / 32), kFtThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        B, K, F, L1, idx_d, val_d, ft_w_d, ft_b_d, ft_out_d);
    NNUE_CHECK_LAUNCH("ft_fwd_indexed_kernel");
    return NNUE_OK;
}

int nnue_ft_bwd_indexed(int B, int K, int F, int L1, const int64_t *idx_d, const float *ft_w_d, const float *g_out_d,
                        int n_pairs, const int32_t *row_d, const int32_t *sample_d, const float *pval_d, float *g_w_d,
                        float *g_b_d, float *g_val_d, void *stream) {
    if (B < 1 || K < 1 || F < 1 || L1 < 1 || !idx_d || !ft_w_d || !g_out_d || !g_w_d || !g_b_d || n_pairs < 0 ||
        (n_pairs > 0 && (!row_d || !sample_d || !pval_d)))
        return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wpb = kFtThreads / 32;
    ft_bwd_indexed_dw_kernel<<<ceil_div(F, wpb), kFtThreads, 0, st>>>(F, L1, n_pairs, row_d, sample_d, pval_d, g_out_d,
                                                                      g_w_d);
    NNUE_CHECK_LAUNCH("ft_bwd_indexed_dw_kernel");
    colsum_full_kernel<<<ceil_div(L1, 128), 128, 0, st>>>(B, L1, g_out_d, g_b_d);
    NNUE_CHECK_LAUNCH("colsum_full_kernel");
    if (g_val_d) {
        ft_bwd_indexed_dval_kernel<<<ceil_div(B, wpb), kFtThreads, 0, st>>>(B, K, F, L1, idx_d, ft_w_d, g_out_d,
                                                                            g_val_d);
        NNUE_CHECK_LAUNCH("ft_bwd_indexed_dval_kernel");
    }
    return NNUE_OK;
}

int nnue_ft_bwd_dw(const nnue_shape *s, const uint32_t *bits_t_d, const float *g_ft_d, float *g_w_d, float *g_b_d,
                   void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !bits_t_d || !g_ft_d || !g_w_d || !g_b_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_ft_bwd_dw(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // bias gradient = column sums of g_ft
    float *cs_partial = static_cast<float *>(workspace_d);
    const int nrow_chunks = ceil_div(s->B, kColsumRows);
    colsum_partial_kernel<<<dim3(ceil_div(s->L1, 128), nrow_chunks), 128, 0, st>>>(s->B, s->L1, g_ft_d, cs_partial);
    NNUE_CHECK_LAUNCH("colsum_partial_kernel");
    fold_partials_kernel<<<ceil_div(s->L1, 128), 128, 0, st>>>(s->L1, nrow_chunks, cs_partial, g_b_d);
    NNUE_CHECK_LAUNCH("fold_partials_kernel");

    const DwPlan d = plan_ft_bwd_dw(*s);
    if (!d.col.LPR || d.smem > (size_t)max_dyn_smem()) {
        ft_bwd_dw_generic_kernel<<<ceil_div(s->F, kFtThreads / 32), kFtThreads, 0, st>>>(*s, bits_t_d, g_ft_d, g_w_d);
        NNUE_CHECK_LAUNCH("ft_bwd_dw_generic_kernel");
        return NNUE_OK;
    }
    float *partial = reinterpret_cast<float *>(static_cast<char *>(workspace_d) +
                                               align_up((size_t)nrow_chunks * s->L1 * 4, 256));
    float *dst = d.direct ? g_w_d : partial;
    if (d.direct && s->P < s->F)
        NNUE_CUDA_TRY(cudaMemsetAsync(g_w_d + (size_t)s->P * s->L1, 0, (size_t)(s->F - s->P) * s->L1 * 4, st));
    int rc = NNUE_OK;
    NNUE_DISPATCH_LPR(d.col.LPR, rc = launch_ft_bwd_dw<LPR>(*s, d, bits_t_d, g_ft_d, dst, st));
    if (rc != NNUE_OK) return rc;
    if (!d.direct) {
        const long long n = 1LL * s->F * (s->L1 / 4);
        ft_bwd_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(*s, d.ngroups, partial, g_w_d);
        NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_kernel");
    }
    return NNUE_OK;
}

int nnue_ft_bwd_dval(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *g_ft_d,
                     float *dval_d, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !g_ft_d || !dval_d) return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const ColPlan cp = col_plan(s->L1);
    if (!cp.LPR) {
        ft_bwd_dval_generic_kernel<<<ceil_div(s->B, kFtThreads / 32), kFtThreads, 0, st>>>(*s, bits_s_d, ft_w_d, g_ft_d,
                                                                                          dval_d);
        NNUE_CHECK_LAUNCH("ft_bwd_dval_generic_kernel");
        return NNUE_OK;
    }
    int rc = NNUE_OK;
    NNUE_DISPATCH_LPR(cp.LPR, rc = launch_ft_bwd_dval<LPR>(*s, bits_s_d, ft_w_d, g_ft_d, dval_d, cp.nchunks, st));
    return rc;
}

}  // extern "C"
