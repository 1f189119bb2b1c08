// ft.cu -- the sparse feature transformer: forward gather-accumulate, weight gradient as a
// segment reduction over the transposed bitmask, and the value gradient (row . grad dot).
//
// Column geometry (plan.cuh ColPlan): a table row is read as float4s by LPR lanes (one
// coalesced 16*LPR-byte segment); with LPR < 32 a warp splits into NG = 32/LPR lane groups that
// take different active features and are combined with __shfl_xor at the end.
//
// Active features are walked straight off the bitmask words (BitWalk): no index lists are
// materialised.  When the table fits one CTA's shared memory it is staged there once per
// persistent CTA with bulk TMA copies (cp.async.bulk -> UBLKCP) and rows are read with LDS.128.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kFtThreads = 256;
constexpr int kFtStagedThreads = 1024;  // one persistent CTA per SM owns the shared-memory table
constexpr int kDvalStagedThreads = 512;  // the value-gradient kernel keeps 4 dots in flight: 16 warps x 128 registers

// Splits the set bits of a warp-uniform 32-bit word over NG lane groups without any cross-lane
// traffic.  The word is cut into NS = max(1, NG/2) segments; inside a segment the "up" group takes
// the bits from the low end and its partner from the high end until they meet, so the two always
// share the work evenly.  Bits are taken most-significant-first with one FLO (the up group walks the
// bit-reversed segment): 4 integer instructions per bit.
template <int NG>
struct BitWalk {
    static constexpr int NS = NG >= 2 ? NG / 2 : 1;
    static constexpr int SW = 32 / NS;
    int seg_shift, flip;
    bool up;
    __device__ __forceinline__ explicit BitWalk(int g) {
        seg_shift = NG >= 2 ? (g >> 1) * SW : 0;
        up = NG == 1 || !(g & 1);
        flip = up ? 0 : 31;
    }
    // my group's view of the word: top-aligned for the up walker, bottom-aligned for the down walker
    __device__ __forceinline__ unsigned view(unsigned word) const {
        const unsigned sub = NS == 1 ? word : (word >> seg_shift) & ((1u << (SW & 31)) - 1u);
        return up ? __brev(sub) : sub;
    }
    // pop the next bit of a view; returns its position in the original word
    __device__ __forceinline__ int take(unsigned &x) const {
        const int c = __clz(x);
        x &= ~(0x80000000u >> c);
        return (c ^ flip) + seg_shift;
    }
    // bits of `word` this group takes, and (warp-uniform) the most any group takes
    __device__ __forceinline__ int mine(unsigned word) const {
        const unsigned sub = NS == 1 ? word : (word >> seg_shift) & ((1u << (SW & 31)) - 1u);
        const int c = __popc(sub);
        return NG == 1 ? c : (up ? (c + 1) >> 1 : c >> 1);
    }
    static __device__ __forceinline__ int trips(unsigned word) {
        int t = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const unsigned seg = NS == 1 ? word : (word >> (s * SW)) & ((1u << (SW & 31)) - 1u);
            const int c = __popc(seg);
            t = max(t, NG >= 2 ? (c + 1) >> 1 : c);
        }
        return t;
    }
};

// Calls f(k, slot) for every bit k of the warp-uniform `word` that belongs to this lane's group; slot
// alternates 0/1 so callers can keep two independent accumulators.  For NG <= 2 (L1 >= 64) all
// iterations but at most one are executed by the whole warp with no validity test.
template <int NG, typename F>
__device__ __forceinline__ void for_my_bits(unsigned word, const BitWalk<NG> &walk, F &&f) {
    unsigned x = walk.view(word);
    if (NG <= 2) {
        const int c = __popc(word);
        const int full = NG == 2 ? c >> 1 : c;  // bits every group takes
        int i = 0;
        for (; i + 1 < full; i += 2) {
            const int k0 = walk.take(x), k1 = walk.take(x);
            f(k0, 0);
            f(k1, 1);
        }
        if (i < full) f(walk.take(x), 0);
        if (NG == 2 && (c & 1) && walk.up) f(walk.take(x), 1);  // odd count: the up group takes the middle bit
    } else {
        const int n = walk.mine(word), trips = BitWalk<NG>::trips(word);
        for (int i = 0; i < trips; i += 2) {
            const int k0 = walk.take(x), k1 = walk.take(x);
            if (i < n) f(k0, 0);
            if (i + 1 < n) f(k1, 1);
        }
    }
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

// Keeps a value in a register: stops the compiler from re-deriving shared-window addresses (S2UR +
// ULEA) inside the hot loops.
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// Row source: the global table (read-only path) or its shared-memory copy.
template <bool STAGED>
struct RowSrc {
    const float *g;   // global: table + column offset
    uint32_t s;       // shared: byte address of table + column offset
    int row_elems;    // L1
    __device__ __forceinline__ float4 load(int row) const {
        if (STAGED) return lds_f4(s + (uint32_t)(row * row_elems) * 4u);
        return __ldg(reinterpret_cast<const float4 *>(g + (size_t)row * row_elems));
    }
};

// CHW position of bit 0 of bitmask word `widx` (word c*CW + j covers cells 32j..32j+31 of channel c)
__device__ __forceinline__ int word_base(const nnue_shape &s, int widx, int cells) {
    return (widx / s.CW) * cells + (widx % s.CW) * 32;
}

// ---- forward on the bitmask ----------------------------------------------------------------
// unit = (sample, column chunk); one warp per unit.
template <int LPR, bool STAGED>
__global__ void __launch_bounds__(STAGED ? kFtStagedThreads : kFtThreads)
ft_fwd_bits_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                   const float *__restrict__ bias, float *__restrict__ out, int nchunks) {
    constexpr int NG = 32 / LPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (STAGED) {
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        tma_stage(smem_raw + 128, w, (uint32_t)((size_t)s.F * s.L1 * sizeof(float)), bar, 0);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int g = lane / LPR, li = lane % LPR;
    const int cells = s.Gh * s.Gw;
    const int last_row = s.F - 1;
    BitWalk<NG> walk(g);
    const uint32_t tab = pin_u32(smem_u32(smem_raw + 128));
    const int units = s.B * nchunks;
    for (int u = blockIdx.x * wpb + warp; u < units; u += gridDim.x * wpb) {
        const int b = nchunks == 1 ? u : u / nchunks, ch = nchunks == 1 ? 0 : u % nchunks;
        const int col = ch * (4 * LPR) + li * 4;
        RowSrc<STAGED> src{w + col, tab + (uint32_t)col * 4u, s.L1};
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        for (int w0 = 0; w0 < s.NW; w0 += 32) {
            const bool have = w0 + lane < s.NW;
            const unsigned mine = have ? __ldg(bits_s + (size_t)b * s.NW + w0 + lane) : 0u;
            const int mybase = have ? word_base(s, w0 + lane, cells) : 0;
            unsigned nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                const unsigned word = __shfl_sync(kFull, mine, jw);
                const int base = __shfl_sync(kFull, mybase, jw);
                for_my_bits<NG>(word, walk, [&](int k, int slot) {
                    const float4 v = src.load(min(base + k, last_row));  // clamp of nnue.py:701
                    if (slot) acc1 = f4_add(acc1, v);
                    else acc0 = f4_add(acc0, v);
                });
            }
        }
        acc0 = group_sum<LPR>(f4_add(acc0, acc1));
        if (g == 0) {
            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + col));
            *reinterpret_cast<float4 *>(out + (size_t)b * s.L1 + col) = f4_add(bv, acc0);
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
            const int base = word_base(s, widx, cells);
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
__global__ void __launch_bounds__(kDwWarps * 32, 2)
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
    const int wpt = TS / 32;  // bitmask words per tile (<= 8)

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    float4 acc[kDwPPW];
#pragma unroll
    for (int q = 0; q < kDwPPW; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pp0 = blockIdx.x * kDwPPC + warp * kDwPPW;
    BitWalk<NG> walk(g);
    const uint32_t tcol = pin_u32(smem_u32(tile) + (uint32_t)li * 16u);

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
        // fetch this tile's bitmask words for my positions while the copy is in flight
        const int nw = min(wpt, s.BW - t * wpt);
        unsigned mine[kDwPPW];
#pragma unroll
        for (int q = 0; q < kDwPPW; ++q) {
            const int pp = pp0 + q;
            const bool live = pp < s.PP && (pp & 31) + ((pp >> 5) % s.CW) * 32 < cells && lane < nw;
            mine[q] = live ? __ldg(bits_t + (size_t)pp * s.BW + (size_t)t * wpt + lane) : 0u;
        }
        mbar_wait(bar, parity);
        parity ^= 1;
#pragma unroll
        for (int q = 0; q < kDwPPW; ++q) {
            unsigned nonzero = __ballot_sync(kFull, mine[q] != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                const unsigned word = __shfl_sync(kFull, mine[q], jw);
                const uint32_t wbase = tcol + (uint32_t)(jw * 32) * (CC * 4);
                float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
                for_my_bits<NG>(word, walk, [&](int k, int slot) {
                    const float4 v = lds_f4(wbase + (uint32_t)k * (CC * 4));
                    if (slot) a1 = f4_add(a1, v);
                    else acc[q] = f4_add(acc[q], v);
                });
                acc[q] = f4_add(acc[q], a1);
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

// ---- dense, register-stationary backward kernels (small tables) ----------------------------------
// With ~43 % of the positions active an index-driven walk spends more
// instructions on bit scanning and cross-lane reductions than a masked dense pass over every
// (sample, position) pair costs, as long as the per-position operand never leaves the register
// file.  Both kernels below give a warp one bitmask word (32 positions, one per lane) for the
// whole kernel and stream the g_ft rows of its sample tiles through shared memory (one bulk TMA
// copy per 32-sample tile, full/empty mbarrier ring, producer = lane 0 of warp 0); a g_ft row is
// read as broadcast LDS.128.  The 32 bitmask words of a tile are loaded one per lane and
// transposed in-register (warp_bit_transpose), so a lane tests "is my position active in sample
// r" with a shift, and the inner loops are branch-free.

// Stage ring shared by the two kernels: tile i of this CTA's stream is tile q + i * nq.
struct TileRing {
    uint64_t *full, *empty;
    float *stages;
    int stage_floats, nq, q, n_mine;
    int p_st; uint32_t p_ph;  // producer cursor: stage of tile i + ahead, parity of the release it waits for
    int st; uint32_t ph;      // consumer cursor
};
constexpr int kTileAhead = kTileStages - 1;  // tiles in flight
template <int L1>
__device__ __forceinline__ void ring_produce(const TileRing &rg, const nnue_shape &s, const float *g_ft, int ii, int st,
                                             uint32_t ph) {
    if (ii >= kTileStages) mbar_wait(&rg.empty[st], ph);
    const int b0 = (rg.q + ii * rg.nq) * kTileTS, rows = min(kTileTS, s.B - b0);
    mbar_arrive_expect_tx(&rg.full[st], (uint32_t)rows * L1 * 4u);
    tma_bulk_g2s(rg.stages + (size_t)st * rg.stage_floats, g_ft + (size_t)b0 * L1, (uint32_t)rows * L1 * 4u,
                 &rg.full[st]);
}
template <int L1>
__device__ __forceinline__ TileRing ring_init(unsigned char *smem_raw, const nnue_shape &s, const float *g_ft, int ntiles,
                                              int nq, int q, int consumers) {
    TileRing rg;
    rg.full = reinterpret_cast<uint64_t *>(smem_raw);
    rg.empty = rg.full + kTileStages;
    rg.stages = reinterpret_cast<float *>(smem_raw + 128);
    rg.stage_floats = kTileTS * L1;
    rg.nq = nq; rg.q = q;
    rg.n_mine = ntiles > q ? (ntiles - q + nq - 1) / nq : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTileStages; ++i) {
            mbar_init(&rg.full[i], 1);
            mbar_init(&rg.empty[i], consumers);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int ii = 0; ii < kTileAhead && ii < rg.n_mine; ++ii) ring_produce<L1>(rg, s, g_ft, ii, ii, 0);
    // tile kTileAhead goes to stage kTileAhead (first use, no wait); the first real wait is on parity 0
    rg.p_st = kTileAhead % kTileStages; rg.p_ph = 1u;
    rg.st = 0; rg.ph = 0;
    return rg;
}
// top of consumer iteration i: the producer lane refills the ring, then everyone waits for tile i
template <int L1>
__device__ __forceinline__ void ring_acquire(TileRing &rg, const nnue_shape &s, const float *g_ft, int i) {
    const int ii = i + kTileAhead;
    if (threadIdx.x == 0 && ii < rg.n_mine) ring_produce<L1>(rg, s, g_ft, ii, rg.p_st, rg.p_ph);
    if (++rg.p_st == kTileStages) { rg.p_st = 0; rg.p_ph ^= 1u; }
    __syncwarp();
    mbar_wait(&rg.full[rg.st], rg.ph);
}
__device__ __forceinline__ void ring_release(TileRing &rg, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&rg.empty[rg.st]);
    if (++rg.st == kTileStages) { rg.st = 0; rg.ph ^= 1u; }
}

// Weight gradient, row-owner form: a lane keeps its table row's gradient (L1 floats) in registers and
// adds g_ft[b] for every sample whose bit is set, walking the samples in order -- the per-position
// sample list is consumed sorted, nothing is atomic.  Output partial[stream][p][L1] for the fold
// kernels below.  The warps of role 0 also sum one float4 column slice of the staged tiles each: the
// bias gradient (column sums of g_ft).
template <int L1>
__global__ void __launch_bounds__(kOwnWarps * 32, 1)
ft_bwd_dw_owner_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ g_ft,
                       float *__restrict__ partial, float *__restrict__ bias_partial, const OwnPlan pl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    TileRing rg = ring_init<L1>(smem_raw, s, g_ft, pl.ntiles, pl.nq, q, kOwnWarps);

    // ---- row owners ----
    const int cells = s.Gh * s.Gw;
    const int widx = role * kOwnWarps + warp;
    const bool active = widx < s.NW;
    const int c = active ? widx / s.CW : 0, j = active ? widx % s.CW : 0;
    const int cell = j * 32 + lane;
    float acc[L1];
#pragma unroll
    for (int k = 0; k < L1; ++k) acc[k] = 0.0f;
    const bool do_bias = role == 0 && warp < L1 / 4;  // role 0 always has at least L1/4 <= kOwnWarps warps
    float4 accb = make_float4(0.f, 0.f, 0.f, 0.f);
    auto tile_word = [&](int i) -> unsigned {  // this lane's row of tile i: sample b0 + lane
        const int b = (q + i * pl.nq) * kTileTS + lane;
        return (active && i < rg.n_mine && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + widx) : 0u;
    };
    unsigned next_word = tile_word(0);
    for (int i = 0; i < rg.n_mine; ++i) {
        const unsigned mybits = warp_bit_transpose(next_word, lane);  // bit r: my position is active in sample b0 + r
        ring_acquire<L1>(rg, s, g_ft, i);
        next_word = tile_word(i + 1);
        const float4 *sg = reinterpret_cast<const float4 *>(rg.stages + (size_t)rg.st * rg.stage_floats);
        const int rows = min(kTileTS, s.B - (q + i * pl.nq) * kTileTS);
        if (do_bias)
            for (int r = 0; r < rows; ++r) accb = f4_add(accb, sg[r * (L1 / 4) + warp]);
        if (active) {
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const float on = (float)((mybits >> r) & 1u);
#pragma unroll
                for (int v = 0; v < L1 / 4; ++v) {
                    const float4 g = sg[r * (L1 / 4) + v];
                    acc[4 * v + 0] = fmaf(on, g.x, acc[4 * v + 0]);
                    acc[4 * v + 1] = fmaf(on, g.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on, g.z, acc[4 * v + 2]);
                    acc[4 * v + 3] = fmaf(on, g.w, acc[4 * v + 3]);
                }
            }
        }
        ring_release(rg, lane);
    }
    if (do_bias && lane == 0) reinterpret_cast<float4 *>(bias_partial + (size_t)q * L1)[warp] = accb;
    if (active && cell < cells) {
        const int p = c * cells + cell;
        float4 *o = reinterpret_cast<float4 *>(partial + ((size_t)q * s.P + p) * L1);
#pragma unroll
        for (int v = 0; v < L1 / 4; ++v) o[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    }
}

// Value gradient, dense form: a warp keeps the table rows of kDvCH channels of one cell word in
// registers and computes dval[b, p] = <W[min(p, F-1)], g_ft[b]> for every sample of its stream,
// masked by the bit (inactive positions get 0): dval [B][PP], one coalesced 128-byte store per
// (sample, word).  Feeds conv_bwd_kernel (input_bwd.cu), which needs no bitmask any more.
template <int L1>
__global__ void __launch_bounds__(kDvWarps * 32, 1)
ft_bwd_dval_dense_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ ft_w,
                         const float *__restrict__ g_ft, float *__restrict__ dval, const DvPlan pl) {
    constexpr int CH = kDvCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    TileRing rg = ring_init<L1>(smem_raw, s, g_ft, pl.ntiles, pl.nq, q, kDvWarps);

    const int cells = s.Gh * s.Gw;
    const int CP = ceil_div(s.C, CH);
    const int unit = role * kDvWarps + warp;  // (channel group, cell word)
    const bool active = unit < CP * s.CW;
    const int cg = active ? unit / s.CW : 0, j = active ? unit % s.CW : 0;
    const int cell = j * 32 + lane;
    float wr[CH][L1];
    bool chan_ok[CH];
    int widx[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int c = cg * CH + k;
        chan_ok[k] = active && c < s.C;
        widx[k] = chan_ok[k] ? c * s.CW + j : 0;
        const int row = min(c * cells + cell, s.F - 1);  // clamp of nnue.py:701
        const bool ok = chan_ok[k] && cell < cells;
#pragma unroll
        for (int v = 0; v < L1 / 4; ++v) {
            const float4 t = ok ? __ldg(reinterpret_cast<const float4 *>(ft_w + (size_t)row * L1) + v)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            wr[k][4 * v + 0] = t.x; wr[k][4 * v + 1] = t.y; wr[k][4 * v + 2] = t.z; wr[k][4 * v + 3] = t.w;
        }
    }
    auto tile_word = [&](int i, int k) -> unsigned {
        const int b = (q + i * pl.nq) * kTileTS + lane;
        return (chan_ok[k] && i < rg.n_mine && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + widx[k]) : 0u;
    };
    unsigned next_word[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) next_word[k] = tile_word(0, k);
    for (int i = 0; i < rg.n_mine; ++i) {
        unsigned mybits[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) mybits[k] = warp_bit_transpose(next_word[k], lane);
        ring_acquire<L1>(rg, s, g_ft, i);
#pragma unroll
        for (int k = 0; k < CH; ++k) next_word[k] = tile_word(i + 1, k);
        if (active) {
            const float4 *sg = reinterpret_cast<const float4 *>(rg.stages + (size_t)rg.st * rg.stage_floats);
            const int b0 = (q + i * pl.nq) * kTileTS, rows = min(kTileTS, s.B - b0);
#pragma unroll 2
            for (int r = 0; r < rows; ++r) {
                float d[CH][4];
#pragma unroll
                for (int k = 0; k < CH; ++k) d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0.0f;
#pragma unroll
                for (int v = 0; v < L1 / 4; ++v) {
                    const float4 g = sg[r * (L1 / 4) + v];
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        d[k][0] = fmaf(wr[k][4 * v + 0], g.x, d[k][0]);
                        d[k][1] = fmaf(wr[k][4 * v + 1], g.y, d[k][1]);
                        d[k][2] = fmaf(wr[k][4 * v + 2], g.z, d[k][2]);
                        d[k][3] = fmaf(wr[k][4 * v + 3], g.w, d[k][3]);
                    }
                }
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const float v = ((mybits[k] >> r) & 1u) ? (d[k][0] + d[k][1]) + (d[k][2] + d[k][3]) : 0.0f;
                    if (chan_ok[k]) dval[(size_t)(b0 + r) * s.PP + (size_t)widx[k] * 32 + lane] = v;
                }
            }
        }
        ring_release(rg, lane);
    }
}

// Both feature-transformer gradients in one pass over g_ft (the training path for small tables): a
// lane keeps its table row AND that row's gradient in registers (2 x L1 floats), so every broadcast
// g_ft float4 feeds eight FMAs -- four for the value gradient <W[row], g_ft[b]>, four for the
// row-owner accumulation -- and shared-memory bandwidth stops being the limit.  Outputs: g_bin
// [B][PP] (masked value gradient), partial[stream][p][L1] and the bias-gradient partials.
template <int L1>
__global__ void __launch_bounds__(kFbWarps * 32, 1)
ft_bwd_both_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ ft_w,
                   const float *__restrict__ g_ft, float *__restrict__ gbin, float *__restrict__ partial,
                   float *__restrict__ bias_partial, const FbPlan pl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    TileRing rg = ring_init<L1>(smem_raw, s, g_ft, pl.ntiles, pl.nq, q, kFbWarps);

    const int cells = s.Gh * s.Gw;
    const int widx = role * kFbWarps + warp;
    const bool active = widx < s.NW;
    const int c = active ? widx / s.CW : 0, j = active ? widx % s.CW : 0;
    const int cell = j * 32 + lane;
    const bool valid = active && cell < cells;
    const int p = c * cells + cell;
    float wr[L1], acc[L1];
    {
        const int row = min(p, s.F - 1);  // clamp of nnue.py:701
#pragma unroll
        for (int v = 0; v < L1 / 4; ++v) {
            const float4 t = valid ? __ldg(reinterpret_cast<const float4 *>(ft_w + (size_t)row * L1) + v)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            wr[4 * v + 0] = t.x; wr[4 * v + 1] = t.y; wr[4 * v + 2] = t.z; wr[4 * v + 3] = t.w;
            acc[4 * v + 0] = acc[4 * v + 1] = acc[4 * v + 2] = acc[4 * v + 3] = 0.0f;
        }
    }
    // bias gradient: the warps of role 0 each sum kFbBiasSlices float4 column slices of the staged tiles
    constexpr int kSlices = (L1 / 4 + kFbWarps - 1) / kFbWarps;
    float4 accb[kSlices];
#pragma unroll
    for (int k = 0; k < kSlices; ++k) accb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool do_bias = role == 0;

    auto tile_word = [&](int i) -> unsigned {
        const int b = (q + i * pl.nq) * kTileTS + lane;
        return (active && i < rg.n_mine && b < s.B) ? __ldg(bits_s + (size_t)b * s.NW + widx) : 0u;
    };
    unsigned next_word = tile_word(0);
    for (int i = 0; i < rg.n_mine; ++i) {
        const unsigned mybits = warp_bit_transpose(next_word, lane);
        ring_acquire<L1>(rg, s, g_ft, i);
        next_word = tile_word(i + 1);
        const float4 *sg = reinterpret_cast<const float4 *>(rg.stages + (size_t)rg.st * rg.stage_floats);
        const int b0 = (q + i * pl.nq) * kTileTS, rows = min(kTileTS, s.B - b0);
        if (do_bias)
            for (int r = 0; r < rows; ++r)
#pragma unroll
                for (int k = 0; k < kSlices; ++k)
                    if (warp + k * kFbWarps < L1 / 4) accb[k] = f4_add(accb[k], sg[r * (L1 / 4) + warp + k * kFbWarps]);
        if (active) {
            float *grow = gbin + (size_t)b0 * s.PP + (size_t)widx * 32 + lane;
            // Four samples per pass.  The FP32 pipe issues one FFMA per cycle only when at most one new
            // register per bank is read, so the loops are ordered for the operand-reuse cache: a table-row
            // element (wr) is reused by the four samples' dots, a sample's 0/1 flag (on) by four columns.
            int r = 0;
#pragma unroll 1
            for (; r + 4 <= rows; r += 4) {
                const unsigned m4 = mybits >> r;
                const float on0 = (float)(m4 & 1u), on1 = (float)((m4 >> 1) & 1u), on2 = (float)((m4 >> 2) & 1u),
                            on3 = (float)((m4 >> 3) & 1u);
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
                const float4 *s0 = sg + r * (L1 / 4);
#pragma unroll
                for (int v = 0; v < L1 / 4; ++v) {
                    const float4 g0 = s0[v], g1 = s0[L1 / 4 + v], g2 = s0[2 * (L1 / 4) + v], g3 = s0[3 * (L1 / 4) + v];
                    d0 = fmaf(wr[4 * v + 0], g0.x, d0); d1 = fmaf(wr[4 * v + 0], g1.x, d1);
                    d2 = fmaf(wr[4 * v + 0], g2.x, d2); d3 = fmaf(wr[4 * v + 0], g3.x, d3);
                    d0 = fmaf(wr[4 * v + 1], g0.y, d0); d1 = fmaf(wr[4 * v + 1], g1.y, d1);
                    d2 = fmaf(wr[4 * v + 1], g2.y, d2); d3 = fmaf(wr[4 * v + 1], g3.y, d3);
                    d0 = fmaf(wr[4 * v + 2], g0.z, d0); d1 = fmaf(wr[4 * v + 2], g1.z, d1);
                    d2 = fmaf(wr[4 * v + 2], g2.z, d2); d3 = fmaf(wr[4 * v + 2], g3.z, d3);
                    d0 = fmaf(wr[4 * v + 3], g0.w, d0); d1 = fmaf(wr[4 * v + 3], g1.w, d1);
                    d2 = fmaf(wr[4 * v + 3], g2.w, d2); d3 = fmaf(wr[4 * v + 3], g3.w, d3);
                    acc[4 * v + 0] = fmaf(on0, g0.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(on0, g0.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on0, g0.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on0, g0.w, acc[4 * v + 3]);
                    acc[4 * v + 0] = fmaf(on1, g1.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(on1, g1.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on1, g1.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on1, g1.w, acc[4 * v + 3]);
                    acc[4 * v + 0] = fmaf(on2, g2.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(on2, g2.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on2, g2.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on2, g2.w, acc[4 * v + 3]);
                    acc[4 * v + 0] = fmaf(on3, g3.x, acc[4 * v + 0]); acc[4 * v + 1] = fmaf(on3, g3.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on3, g3.z, acc[4 * v + 2]); acc[4 * v + 3] = fmaf(on3, g3.w, acc[4 * v + 3]);
                }
                float *gr = grow + (size_t)r * s.PP;
                gr[0] = d0 * on0;
                gr[(size_t)s.PP] = d1 * on1;
                gr[2 * (size_t)s.PP] = d2 * on2;
                gr[3 * (size_t)s.PP] = d3 * on3;
            }
#pragma unroll 1
            for (; r < rows; ++r) {  // tail of the batch's last tile
                const bool bit = (mybits >> r) & 1u;
                const float on = bit ? 1.0f : 0.0f;
                float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
#pragma unroll
                for (int v = 0; v < L1 / 4; ++v) {
                    const float4 g = sg[r * (L1 / 4) + v];
                    d0 = fmaf(wr[4 * v + 0], g.x, d0);
                    d1 = fmaf(wr[4 * v + 1], g.y, d1);
                    d2 = fmaf(wr[4 * v + 2], g.z, d2);
                    d3 = fmaf(wr[4 * v + 3], g.w, d3);
                    acc[4 * v + 0] = fmaf(on, g.x, acc[4 * v + 0]);
                    acc[4 * v + 1] = fmaf(on, g.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(on, g.z, acc[4 * v + 2]);
                    acc[4 * v + 3] = fmaf(on, g.w, acc[4 * v + 3]);
                }
                grow[(size_t)r * s.PP] = bit ? (d0 + d1) + (d2 + d3) : 0.0f;
            }
        }
        ring_release(rg, lane);
    }
    if (do_bias && lane == 0)
#pragma unroll
        for (int k = 0; k < kSlices; ++k)
            if (warp + k * kFbWarps < L1 / 4)
                reinterpret_cast<float4 *>(bias_partial + (size_t)q * L1)[warp + k * kFbWarps] = accb[k];
    if (valid) {
        float4 *o = reinterpret_cast<float4 *>(partial + ((size_t)q * s.P + p) * L1);
#pragma unroll
        for (int v = 0; v < L1 / 4; ++v) o[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    }
}

// Fold stage 1: every position p sums its partials over the tile groups.  Positions below F-1 are
// table rows and go straight to g_w; positions >= F-1 all alias onto the last row (the clamp of
// nnue.py:701) and are parked in `alias` [P-(F-1)][L1] for stage 2; rows in [P, F-1) get zeros.
__global__ void ft_bwd_dw_fold_kernel(const nnue_shape s, int ngroups, const float *__restrict__ partial,
                                      float *__restrict__ g_w, float *__restrict__ alias) {
    const int v4 = s.L1 / 4;
    const int nrows = max(s.P, s.F - 1);
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1LL * nrows * v4) return;
    const int p = (int)(i / v4), c4 = (int)(i % v4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < s.P)
        for (int gI = 0; gI < ngroups; ++gI)
            acc = f4_add(acc, __ldg(reinterpret_cast<const float4 *>(partial + ((size_t)gI * s.P + p) * s.L1) + c4));
    if (p < s.F - 1) reinterpret_cast<float4 *>(g_w + (size_t)p * s.L1)[c4] = acc;
    else reinterpret_cast<float4 *>(alias + (size_t)(p - (s.F - 1)) * s.L1)[c4] = acc;
}
// Fold stage 2: g_w[F-1] = sum of the parked rows, 8 slices per column combined in a fixed order.
__global__ void __launch_bounds__(256)
ft_bwd_dw_fold_last_kernel(const nnue_shape s, const float *__restrict__ alias, float *__restrict__ g_w) {
    __shared__ float4 red[256];
    const int v4 = s.L1 / 4;
    const int nalias = max(0, s.P - (s.F - 1));
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c4 = blockIdx.x * 32 + col;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < v4)
        for (int i = sl; i < nalias; i += 8)
            acc = f4_add(acc, __ldg(reinterpret_cast<const float4 *>(alias + (size_t)i * s.L1) + c4));
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sl == 0 && c4 < v4) {
        for (int k = 1; k < 8; ++k) acc = f4_add(acc, red[k * 32 + col]);
        reinterpret_cast<float4 *>(g_w + (size_t)(s.F - 1) * s.L1)[c4] = acc;
    }
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

// ---- value gradient at active positions (+ threshold gradient) ---------------------------------
// dval[b, p] = <W[min(p, F-1)], g_ft[b]>.  One warp per sample; each lane group takes an active
// position and its LPR lanes share the dot.  Four positions per group are in flight at once and
// their partial dots are combined with a transposed shuffle reduction (5 shuffles for 4 dots
// instead of 4 x log2(LPR)); the lanes holding a finished dot store it.
// The same lanes fold in the straight-through threshold gradient (nnue.py:36-52):
//   g_thr[c] = -sum over active (b, p in channel c) of dval * k * sig * (1 - sig),  sig = sigmoid(k (x - thr[c]))
// with x the stored pre-threshold activation; per-CTA partials go to thr_partial[blockIdx.x][C].
constexpr float kSteSharpnessFt = 10.0f;  // nnue.py:41

template <int LPR, bool STAGED>
__global__ void __launch_bounds__(STAGED ? kDvalStagedThreads : kFtThreads, 1)
ft_bwd_dval_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                   const float *__restrict__ g_ft, const float *__restrict__ xpad, const float *__restrict__ thr,
                   float *__restrict__ dval, float *__restrict__ thr_partial, int nchunks) {
    constexpr int NG = 32 / LPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [0,128) mbarrier | per-warp threshold-gradient sums [wpb][C] | table (STAGED only)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float *sdth = reinterpret_cast<float *>(smem_raw + 128);
    const uint32_t tab_off = 128 + (uint32_t)align_up((size_t)wpb * s.C * 4, 128);
    for (int i = threadIdx.x; i < wpb * s.C; i += blockDim.x) sdth[i] = 0.0f;
    if (STAGED) {
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        tma_stage(smem_raw + tab_off, w, (uint32_t)((size_t)s.F * s.L1 * sizeof(float)), bar, 0);
    } else {
        __syncthreads();
    }
    const int g = lane / LPR, li = lane % LPR;
    const int cells = s.Gh * s.Gw;
    const int last_row = s.F - 1;
    const bool hi = (li & (LPR / 2)) != 0, hi2 = (li & (LPR / 4)) != 0;
    const int my_slot = (hi ? 2 : 0) + (hi2 ? 1 : 0);  // which of the 4 in-flight dots this lane finishes
    const bool storer = (li & (LPR / 4 - 1)) == 0;
    BitWalk<NG> walk(g);
    RowSrc<STAGED> src{w + li * 4, pin_u32(smem_u32(smem_raw + tab_off) + (uint32_t)li * 16u), s.L1};
    float dth = 0.0f;   // this lane's share of g_thr[cur_c]
    int cur_c = 0;
    float *my_sdth = sdth + warp * s.C;
    for (int b = blockIdx.x * wpb + warp; b < s.B; b += gridDim.x * wpb) {
        const float *grow = g_ft + (size_t)b * s.L1 + li * 4;
        const float4 g0 = nchunks == 1 ? __ldg(reinterpret_cast<const float4 *>(grow)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float *drow = dval + (size_t)b * s.PP;
        const float *xrow = xpad + (size_t)b * s.PP;
        for (int w0 = 0; w0 < s.NW; w0 += 32) {
            const bool have = w0 + lane < s.NW;
            const unsigned mine = have ? __ldg(bits_s + (size_t)b * s.NW + w0 + lane) : 0u;
            const int myc = have ? (w0 + lane) / s.CW : 0;
            unsigned nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                const unsigned word = __shfl_sync(kFull, mine, jw);
                const int c = __shfl_sync(kFull, myc, jw);
                const int widx = w0 + jw;
                const int base = c * cells + (widx - c * s.CW) * 32;
                if (c != cur_c) {  // warp-uniform: close the running channel sum
                    const float v = warp_sum(dth);
                    if (lane == 0) my_sdth[cur_c] += v;
                    dth = 0.0f;
                    cur_c = c;
                }
                const float thr_c = __ldg(thr + c);
                unsigned x = walk.view(word);
                const int n = walk.mine(word), trips = BitWalk<NG>::trips(word);
                for (int i = 0; i < trips; i += 4) {
                    int k[4];
                    float d[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        k[u] = walk.take(x);
                        d[u] = 0.0f;
                        if (i + u < n) {
                            const int row = min(base + k[u], last_row);
                            if (nchunks == 1) {
                                const float4 v = src.load(row);
                                d[u] = fmaf(v.x, g0.x, fmaf(v.y, g0.y, fmaf(v.z, g0.z, v.w * g0.w)));
                            } else {
                                const float *wrow = w + (size_t)row * s.L1 + li * 4;
                                for (int ch = 0; ch < nchunks; ++ch) {
                                    const float4 v = __ldg(reinterpret_cast<const float4 *>(wrow + ch * 128));
                                    const float4 gv = __ldg(reinterpret_cast<const float4 *>(grow + ch * 128));
                                    d[u] = fmaf(v.x, gv.x, fmaf(v.y, gv.y, fmaf(v.z, gv.z, fmaf(v.w, gv.w, d[u]))));
                                }
                            }
                        }
                    }
                    // transposed reduction: 4 dots x LPR lanes -> each lane ends with one full dot
                    float k0 = hi ? d[2] : d[0], k1 = hi ? d[3] : d[1];
                    k0 += __shfl_xor_sync(kFull, hi ? d[0] : d[2], LPR / 2);
                    k1 += __shfl_xor_sync(kFull, hi ? d[1] : d[3], LPR / 2);
                    float r = hi2 ? k1 : k0;
                    r += __shfl_xor_sync(kFull, hi2 ? k0 : k1, LPR / 4);
#pragma unroll
                    for (int o = LPR / 8; o > 0; o >>= 1) r += __shfl_xor_sync(kFull, r, o);
                    const int kk = my_slot == 0 ? k[0] : my_slot == 1 ? k[1] : my_slot == 2 ? k[2] : k[3];
                    if (storer && i + my_slot < n) {
                        const int pp = widx * 32 + kk;
                        drow[pp] = r;
                        const float z = kSteSharpnessFt * (__ldg(xrow + pp) - thr_c);
                        const float sg = __fdividef(1.0f, 1.0f + __expf(-z));
                        dth = fmaf(-r, kSteSharpnessFt * sg * (1.0f - sg), dth);
                    }
                }
            }
        }
    }
    {
        const float v = warp_sum(dth);
        if (lane == 0) my_sdth[cur_c] += v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
        float v = 0.0f;
        for (int wv = 0; wv < wpb; ++wv) v += sdth[wv * s.C + c];
        thr_partial[(size_t)blockIdx.x * s.C + c] = v;
    }
}

// any L1: one warp per sample and one dot at a time; same outputs as the vector kernel
__global__ void __launch_bounds__(kFtThreads)
ft_bwd_dval_generic_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                           const float *__restrict__ g_ft, const float *__restrict__ xpad,
                           const float *__restrict__ thr, float *__restrict__ dval, float *__restrict__ thr_partial) {
    extern __shared__ float sdth_gen[];  // [wpb][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int i = threadIdx.x; i < wpb * s.C; i += blockDim.x) sdth_gen[i] = 0.0f;
    __syncthreads();
    const int cells = s.Gh * s.Gw;
    for (int b = blockIdx.x * wpb + warp; b < s.B; b += gridDim.x * wpb) {
        for (int widx = 0; widx < s.NW; ++widx) {
            unsigned word = __ldg(bits_s + (size_t)b * s.NW + widx);
            const int c = widx / s.CW;
            const int base = word_base(s, widx, cells);
            while (word) {
                const int k = __ffs(word) - 1;
                word &= word - 1;
                const int row = min(base + k, s.F - 1);
                float d = 0.0f;
                for (int col = lane; col < s.L1; col += 32)
                    d = fmaf(__ldg(w + (size_t)row * s.L1 + col), __ldg(g_ft + (size_t)b * s.L1 + col), d);
                d = warp_sum(d);
                if (lane == 0) {
                    const size_t o = (size_t)b * s.PP + (size_t)widx * 32 + k;
                    dval[o] = d;
                    const float z = kSteSharpnessFt * (__ldg(xpad + o) - __ldg(thr + c));
                    const float sg = __fdividef(1.0f, 1.0f + __expf(-z));
                    sdth_gen[warp * s.C + c] = fmaf(-d, kSteSharpnessFt * sg * (1.0f - sg), sdth_gen[warp * s.C + c]);
                }
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < s.C; c += blockDim.x) {
        float v = 0.0f;
        for (int wv = 0; wv < wpb; ++wv) v += sdth_gen[wv * s.C + c];
        thr_partial[(size_t)blockIdx.x * s.C + c] = v;
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

// Stage the table when it fits one CTA's shared memory and there is enough work to amortise it.
static bool use_staging(const nnue_shape &s, int nchunks, long long units) {
    const size_t table_bytes = (size_t)s.F * s.L1 * 4;
    const int mode = get_option(kOptFtFwdStaging);  // 0 never, 1 auto, 2 whenever it fits
    const bool fits = nchunks == 1 && table_bytes + 128 <= (size_t)max_dyn_smem() && table_bytes % 16 == 0 &&
                      table_bytes < (1u << 20);
    return fits && (mode == 2 || (mode == 1 && units >= 8LL * kNumSMs));
}

template <int LPR>
static int launch_ft_fwd(const nnue_shape &s, const uint32_t *bits, const float *w, const float *b, float *out,
                         int nchunks, cudaStream_t st) {
    const size_t staged_smem = (size_t)s.F * s.L1 * 4 + 128;
    const long long units = 1LL * s.B * nchunks;
    if (use_staging(s, nchunks, units)) {
        auto k = ft_fwd_bits_kernel<LPR, true>;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_smem));
        k<<<kNumSMs, kFtStagedThreads, staged_smem, st>>>(s, bits, w, b, out, nchunks);
    } else {
        const int wpb = kFtThreads / 32;
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
                              const float *xpad, const float *thr, float *dval, float *thr_partial, int grid,
                              bool staged, int nchunks, cudaStream_t st) {
    if (staged) {
        const size_t smem = 128 + align_up((size_t)(kDvalStagedThreads / 32) * s.C * 4, 128) + (size_t)s.F * s.L1 * 4;
        auto k = ft_bwd_dval_kernel<LPR, true>;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kDvalStagedThreads, smem, st>>>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, nchunks);
    } else {
        const size_t smem = 128 + align_up((size_t)(kFtThreads / 32) * s.C * 4, 128);
        ft_bwd_dval_kernel<LPR, false><<<grid, kFtThreads, smem, st>>>(s, bits, w, g_ft, xpad, thr, dval, thr_partial,
                                                                       nchunks);
    }
    NNUE_CHECK_LAUNCH("ft_bwd_dval_kernel");
    return NNUE_OK;
}

// g_thr[c] = -sum_{b, cell} g_bin[b, c, cell] * k * sig * (1 - sig), sig = sigmoid(k (x - thr[c]))   (nnue.py:36-52)
// over dense [B][PP] rows (g_bin is zero at inactive positions).  CTA = (channel, chunk of samples); per-thread
// sums, a fixed-order block reduction, partial[chunk][C] for fold_partials_kernel.
__global__ void __launch_bounds__(256)
thr_grad_kernel(const nnue_shape s, const float *__restrict__ gbin, const float *__restrict__ xpad,
                const float *__restrict__ thr, float *__restrict__ partial, int rows_per_chunk) {
    __shared__ float red[256];
    const int c = blockIdx.x, chunk = blockIdx.y;
    const int b0 = chunk * rows_per_chunk, b1 = min(s.B, b0 + rows_per_chunk);
    const int cw32 = s.CW * 32, cells = s.Gh * s.Gw;
    const float thr_c = __ldg(thr + c);
    float acc = 0.0f;
    const long long n = 1LL * max(0, b1 - b0) * cw32;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const int b = b0 + (int)(i / cw32), cell = (int)(i % cw32);
        if (cell >= cells) continue;
        const size_t at = (size_t)b * s.PP + (size_t)c * cw32 + cell;
        const float g = __ldg(gbin + at);
        if (g != 0.0f) {
            const float z = kSteSharpnessFt * (__ldg(xpad + at) - thr_c);
            const float sgm = __fdividef(1.0f, 1.0f + __expf(-z));
            acc = fmaf(-g, kSteSharpnessFt * sgm * (1.0f - sgm), acc);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)chunk * s.C + c] = red[0];
}

int launch_ft_bwd_dval_dense(const nnue_shape &s, const uint32_t *bits_s, const float *ft_w, const float *g_ft,
                             float *dval, cudaStream_t st) {
    const DvPlan pl = plan_ft_bwd_dval_dense(s);
    if (!pl.ok) return NNUE_ERR_UNSUPPORTED;
    auto k = s.L1 == 64 ? ft_bwd_dval_dense_kernel<64> : ft_bwd_dval_dense_kernel<32>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k<<<pl.grid, kDvWarps * 32, pl.smem, st>>>(s, bits_s, ft_w, g_ft, dval, pl);
    NNUE_CHECK_LAUNCH("ft_bwd_dval_dense_kernel");
    return NNUE_OK;
}


// ---- (row, sample, value) triples of the indexed interface sorted by row: a chunked stable counting sort ------------
// The weight gradient of FeatureTransformer.forward(idx, val) (nnue.py:702-708: autograd's index_put accumulates in
// whatever order its atomics land) is a segment reduction over the pairs sorted by table row, in (sample, slot) order
// inside a row -- deterministic.  Keys are min(idx, F - 1) (the clamp of nnue.py:701), idx < 0 sorts behind everything as
// key F with value 0, so the pair count is always B * K and nothing is read back to the host.
//   pass A  a warp per chunk of pairs counts its keys (match.any aggregates equal keys of a 32-pair group)
//   pass B  per key, the counts become exclusive prefixes over the chunks; one CTA scans the per-key totals
//   pass C  the warps walk their chunks again and place every pair at base[key] + prefix[chunk][key] + rank in group
constexpr int kSortChunkPairs = 2048, kSortMaxChunks = 256;
__host__ __device__ inline int sort_chunks(long long n) {
    long long c = (n + kSortChunkPairs - 1) / kSortChunkPairs;
    return (int)(c < 1 ? 1 : (c > kSortMaxChunks ? kSortMaxChunks : c));
}
__device__ __forceinline__ int sort_key(long long v, int F) { return v < 0 ? F : (int)(v < F ? v : F - 1); }

__global__ void __launch_bounds__(128)
ft_sort_count_kernel(long long n, int F, int C, const int64_t *__restrict__ idx, unsigned *__restrict__ hist) {
    const int lane = threadIdx.x & 31, c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= C) return;
    const int NK = F + 1;
    unsigned *h = hist + (size_t)c * NK;
    for (int f = lane; f < NK; f += 32) h[f] = 0u;
    __syncwarp();
    const long long per = (n + C - 1) / C, i0 = (long long)c * per, i1 = i0 + per < n ? i0 + per : n;
    for (long long i = i0; i < i1; i += 32) {
        const bool on = i + lane < i1;
        const int key = on ? sort_key(__ldg(idx + i + lane), F) : -1 - lane;  // idle lanes get unique keys
        const unsigned peers = __match_any_sync(kFull, key);
        if (on && lane == __ffs(peers) - 1) h[key] += __popc(peers);
        __syncwarp();
    }
}
// hist[c][f] -> exclusive prefix over c; tot[f] = total of key f
__global__ void ft_sort_prefix_kernel(int NK, int C, unsigned *__restrict__ hist, unsigned *__restrict__ tot) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= NK) return;
    unsigned run = 0u;
    for (int c = 0; c < C; ++c) {
        const unsigned v = hist[(size_t)c * NK + f];
        hist[(size_t)c * NK + f] = run;
        run += v;
    }
    tot[f] = run;
}
// exclusive scan of tot[0 .. NK) into base, one CTA
__global__ void __launch_bounds__(1024)
ft_sort_scan_kernel(int NK, const unsigned *__restrict__ tot, unsigned *__restrict__ base) {
    __shared__ unsigned part[1024];
    const int per = (NK + 1023) / 1024, f0 = threadIdx.x * per, f1 = min(NK, f0 + per);
    unsigned sum = 0u;
    for (int f = f0; f < f1; ++f) sum += tot[f];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
        const unsigned v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - sum;
    for (int f = f0; f < f1; ++f) {
        base[f] = run;
        run += tot[f];
    }
}
__global__ void __launch_bounds__(128)
ft_sort_place_kernel(long long n, int K, int F, int C, const int64_t *__restrict__ idx, const float *__restrict__ val,
                     unsigned *__restrict__ hist, const unsigned *__restrict__ base, int32_t *__restrict__ row_out,
                     int32_t *__restrict__ sample_out, float *__restrict__ pval_out) {
    const int lane = threadIdx.x & 31, c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= C) return;
    const int NK = F + 1;
    unsigned *h = hist + (size_t)c * NK;  // running cursor of this chunk per key (starts at the prefix over earlier chunks)
    const long long per = (n + C - 1) / C, i0 = (long long)c * per, i1 = i0 + per < n ? i0 + per : n;
    for (long long i = i0; i < i1; i += 32) {
        const bool on = i + lane < i1;
        const long long raw = on ? __ldg(idx + i + lane) : 0;
        const int key = on ? sort_key(raw, F) : -1 - lane;
        const unsigned peers = __match_any_sync(kFull, key);
        const unsigned cur = on ? h[key] : 0u;
        __syncwarp();
        if (on && lane == __ffs(peers) - 1) h[key] = cur + __popc(peers);
        __syncwarp();
        if (on) {
            const unsigned pos = __ldg(base + key) + cur + __popc(peers & ((1u << lane) - 1u));
            row_out[pos] = key;
            sample_out[pos] = (int32_t)((i + lane) / K);
            pval_out[pos] = raw < 0 ? 0.0f : __ldg(val + i + lane);
        }
    }
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
                float *ft_out_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !ft_b_d || !ft_out_d) return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // small tables with scratch available: tensor-core contraction (ft_mma.cu); otherwise the row gather
    if (ft_umma_ok(*s) && workspace_d && workspace_bytes >= umma_kt_bytes((size_t)s->PP, *s))  // tcgen05 / TMEM form
        return launch_ft_fwd_umma(*s, bits_s_d, ft_w_d, ft_b_d, ft_out_d, workspace_d, st);
    if (plan_ft_mma(*s).ok && workspace_d && workspace_bytes >= ws_ft_fwd(*s))
        return launch_ft_fwd_mma(*s, bits_s_d, ft_w_d, ft_b_d, ft_out_d, workspace_d, st);
    const ColPlan cp = col_plan(s->L1);
    // tables that do not fit one CTA's shared memory: rows staged through a shared-memory ring by bulk TMA (ft_gather.cu)
    if (ft_gather_ok(*s) && !use_staging(*s, cp.nchunks, 1LL * s->B * cp.nchunks))
        return launch_ft_gather_fwd(*s, bits_s_d, ft_w_d, ft_b_d, ft_out_d, workspace_d, workspace_bytes, st);
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

// ---- pre-formatted table tiles (tcgen05 shapes): the formatting depends only on the weights and can run beside the
//      extraction on another stream ----
size_t nnue_ft_tables_bytes(const nnue_shape *s) { return s ? ft_tables_bytes(*s) : 0; }

int nnue_ft_format_tables(const nnue_shape *s, const float *ft_w_d, void *tables_d, void *stream) {
    if (!s || !ft_w_d || !tables_d) return NNUE_ERR_INVALID_ARG;
    if (!ft_umma_ok(*s)) return NNUE_ERR_UNSUPPORTED;
    return launch_ft_format_tables(*s, ft_w_d, tables_d, 3, static_cast<cudaStream_t>(stream));
}

int nnue_ft_fwd_tables(const nnue_shape *s, const uint32_t *bits_s_d, const void *tables_d, const float *ft_b_d,
                       float *ft_out_d, void *stream) {
    if (!s || !bits_s_d || !tables_d || !ft_b_d || !ft_out_d) return NNUE_ERR_INVALID_ARG;
    if (!ft_umma_ok(*s)) return NNUE_ERR_UNSUPPORTED;
    return launch_ft_fwd_umma_tiles(*s, bits_s_d, tables_d, ft_b_d, ft_out_d, static_cast<cudaStream_t>(stream));
}

int nnue_ft_bwd_gbin_tables(const nnue_shape *s, const uint32_t *bits_s_d, const void *tables_d, const float *g_ft_d,
                            float *gbin_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !bits_s_d || !tables_d || !g_ft_d || !gbin_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    if (!ft_umma_ok(*s) || !plan_input_bwd(*s).fused) return NNUE_ERR_UNSUPPORTED;
    if (workspace_bytes < ws_ft_gbin_umma(*s)) return NNUE_ERR_WORKSPACE;
    return launch_ft_bwd_gbin_umma(*s, bits_s_d, nullptr, g_ft_d, workspace_d, gbin_d, static_cast<cudaStream_t>(stream), tables_d);
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

size_t nnue_ft_sort_pairs_workspace_bytes(int B, int K, int F) {
    if (B < 1 || K < 1 || F < 1) return 0;
    return ((size_t)sort_chunks(1LL * B * K) * (F + 1) + 2 * (size_t)(F + 1)) * 4 + 256;
}

int nnue_ft_sort_pairs(int B, int K, int F, const int64_t *idx_d, const float *val_d, int32_t *row_out_d, int32_t *sample_out_d,
                       float *pval_out_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (B < 1 || K < 1 || F < 1 || !idx_d || !val_d || !row_out_d || !sample_out_d || !pval_out_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < nnue_ft_sort_pairs_workspace_bytes(B, K, F)) return NNUE_ERR_WORKSPACE;
    if (1LL * B * K > 2147483647LL) return NNUE_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = 1LL * B * K;
    const int C = sort_chunks(n), NK = F + 1;
    unsigned *hist = static_cast<unsigned *>(workspace_d), *tot = hist + (size_t)C * NK, *base = tot + NK;
    ft_sort_count_kernel<<<ceil_div(C, 4), 128, 0, st>>>(n, F, C, idx_d, hist);
    NNUE_CHECK_LAUNCH("ft_sort_count_kernel");
    ft_sort_prefix_kernel<<<ceil_div(NK, 256), 256, 0, st>>>(NK, C, hist, tot);
    NNUE_CHECK_LAUNCH("ft_sort_prefix_kernel");
    ft_sort_scan_kernel<<<1, 1024, 0, st>>>(NK, tot, base);
    NNUE_CHECK_LAUNCH("ft_sort_scan_kernel");
    ft_sort_place_kernel<<<ceil_div(C, 4), 128, 0, st>>>(n, K, F, C, idx_d, val_d, hist, base, row_out_d, sample_out_d, pval_out_d);
    NNUE_CHECK_LAUNCH("ft_sort_place_kernel");
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

// tensor-core weight/bias gradient; workspace layout: wfrag (value gradient) | gfrag | bias partials | partials | alias
static int ft_bwd_dw_mma_path(const nnue_shape *s, const uint32_t *bits_s_d, const float *g_ft_d, float *g_w_d, float *g_b_d,
                              void *workspace_d, size_t workspace_bytes, cudaStream_t st) {
    if (workspace_bytes < ws_ft_bwd_mma(*s)) return NNUE_ERR_WORKSPACE;
    const MmaPlan mp = plan_ft_mma(*s);
    char *ws = static_cast<char *>(workspace_d) + ws_bwd_front_mma(*s);
    uint4 *gfrag = reinterpret_cast<uint4 *>(ws); ws += align_up(mma_gfrag_bytes(*s), 256);
    float *bias_partial = reinterpret_cast<float *>(ws); ws += align_up((size_t)mp.n_chunks * s->L1 * 4, 256);
    float *partial = reinterpret_cast<float *>(ws);
    float *alias = partial + (size_t)mp.n_chunks * s->P * s->L1;
    int n_chunks = 0;
    const int rc = launch_ft_bwd_dw_mma(*s, bits_s_d, g_ft_d, gfrag, partial, bias_partial, &n_chunks, st);
    if (rc != NNUE_OK) return rc;
    fold_partials_kernel<<<ceil_div(s->L1, 128), 128, 0, st>>>(s->L1, n_chunks, bias_partial, g_b_d);
    NNUE_CHECK_LAUNCH("fold_partials_kernel");
    const long long n = 1LL * (s->P > s->F - 1 ? s->P : s->F - 1) * (s->L1 / 4);
    ft_bwd_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(*s, n_chunks, partial, g_w_d, alias);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_kernel");
    ft_bwd_dw_fold_last_kernel<<<ceil_div(s->L1 / 4, 32), 256, 0, st>>>(*s, alias, g_w_d);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_last_kernel");
    return NNUE_OK;
}

// tcgen05 weight/bias gradient; it owns the second region of the workspace (after the value gradient's operands)
static int ft_bwd_dw_umma_path(const nnue_shape *s, const uint32_t *bits_s_d, const float *g_ft_d, float *g_w_d, float *g_b_d,
                               void *workspace_d, size_t workspace_bytes, cudaStream_t st) {
    if (workspace_bytes < ws_ft_bwd_umma(*s)) return NNUE_ERR_WORKSPACE;
    return launch_ft_bwd_dw_umma(*s, bits_s_d, g_ft_d, static_cast<char *>(workspace_d) + ws_bwd_front_umma(*s), g_w_d, g_b_d, st);
}

int nnue_ft_uses_umma(const nnue_shape *s) { return s && ft_umma_ok(*s) ? 1 : 0; }

int nnue_ft_uses_mma(const nnue_shape *s) { return s && (ft_umma_ok(*s) || plan_ft_mma(*s).ok) ? 1 : 0; }

int nnue_ft_bwd_is_fused(const nnue_shape *s) {
    return s && (ft_umma_ok(*s) || plan_ft_mma(*s).ok || plan_ft_bwd_both(*s).ok) && plan_input_bwd(*s).fused ? 1 : 0;
}

int nnue_ft_bwd(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *g_ft_d, float *g_w_d,
                float *g_b_d, float *gbin_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !g_ft_d || !g_w_d || !g_b_d || !gbin_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    if (ft_umma_ok(*s)) {  // tcgen05 path
        const int rc = ft_bwd_dw_umma_path(s, bits_s_d, g_ft_d, g_w_d, g_b_d, workspace_d, workspace_bytes,
                                           static_cast<cudaStream_t>(stream));
        if (rc != NNUE_OK) return rc;
        return launch_ft_bwd_gbin_umma(*s, bits_s_d, ft_w_d, g_ft_d, workspace_d, gbin_d, static_cast<cudaStream_t>(stream));
    }
    if (plan_ft_mma(*s).ok) {  // tensor-core path
        const int rc = ft_bwd_dw_mma_path(s, bits_s_d, g_ft_d, g_w_d, g_b_d, workspace_d, workspace_bytes,
                                          static_cast<cudaStream_t>(stream));
        if (rc != NNUE_OK) return rc;
        return launch_ft_bwd_gbin_mma(*s, bits_s_d, ft_w_d, g_ft_d, static_cast<uint4 *>(workspace_d), gbin_d,
                                      static_cast<cudaStream_t>(stream));
    }
    const FbPlan fb = plan_ft_bwd_both(*s);
    if (!fb.ok) return NNUE_ERR_UNSUPPORTED;
    if (workspace_bytes < ws_ft_bwd_both(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *bias_partial = static_cast<float *>(workspace_d);
    float *partial = reinterpret_cast<float *>(static_cast<char *>(workspace_d) + align_up((size_t)fb.nq * s->L1 * 4, 256));
    float *alias = partial + (size_t)fb.nq * s->P * s->L1;
    auto k = s->L1 == 64 ? ft_bwd_both_kernel<64> : ft_bwd_both_kernel<32>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fb.smem));
    k<<<fb.grid, kFbWarps * 32, fb.smem, st>>>(*s, bits_s_d, ft_w_d, g_ft_d, gbin_d, partial, bias_partial, fb);
    NNUE_CHECK_LAUNCH("ft_bwd_both_kernel");
    fold_partials_kernel<<<ceil_div(s->L1, 128), 128, 0, st>>>(s->L1, fb.nq, bias_partial, g_b_d);
    NNUE_CHECK_LAUNCH("fold_partials_kernel");
    const long long n = 1LL * (s->P > s->F - 1 ? s->P : s->F - 1) * (s->L1 / 4);
    ft_bwd_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(*s, fb.nq, partial, g_w_d, alias);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_kernel");
    ft_bwd_dw_fold_last_kernel<<<ceil_div(s->L1 / 4, 32), 256, 0, st>>>(*s, alias, g_w_d);
    NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_last_kernel");
    return NNUE_OK;
}

int nnue_wants_transposed_bits(const nnue_shape *s) { return s && (ft_umma_ok(*s) || plan_ft_bwd_dw_owner(*s).ok) ? 0 : 1; }

int nnue_ft_bwd_dw(const nnue_shape *s, const uint32_t *bits_s_d, const uint32_t *bits_t_d, const float *g_ft_d,
                   float *g_w_d, float *g_b_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !g_ft_d || !g_w_d || !g_b_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ft_umma_ok(*s) && bits_s_d)
        return ft_bwd_dw_umma_path(s, bits_s_d, g_ft_d, g_w_d, g_b_d, workspace_d, workspace_bytes, st);
    if (plan_ft_mma(*s).ok && bits_s_d)
        return ft_bwd_dw_mma_path(s, bits_s_d, g_ft_d, g_w_d, g_b_d, workspace_d, workspace_bytes, st);
    if (workspace_bytes < ws_ft_bwd_dw(*s)) return NNUE_ERR_WORKSPACE;
    const OwnPlan own = plan_ft_bwd_dw_owner(*s);
    if (own.ok) {
        if (!bits_s_d) return NNUE_ERR_INVALID_ARG;
        float *bias_partial = static_cast<float *>(workspace_d);
        float *partial = reinterpret_cast<float *>(static_cast<char *>(workspace_d) +
                                                   align_up((size_t)own.nq * s->L1 * 4, 256));
        float *alias = partial + (size_t)own.nq * s->P * s->L1;
        if (s->L1 == 64) {
            auto k = ft_bwd_dw_owner_kernel<64>;
            NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)own.smem));
            k<<<own.grid, kOwnWarps * 32, own.smem, st>>>(*s, bits_s_d, g_ft_d, partial, bias_partial, own);
        } else {
            auto k = ft_bwd_dw_owner_kernel<32>;
            NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)own.smem));
            k<<<own.grid, kOwnWarps * 32, own.smem, st>>>(*s, bits_s_d, g_ft_d, partial, bias_partial, own);
        }
        NNUE_CHECK_LAUNCH("ft_bwd_dw_owner_kernel");
        fold_partials_kernel<<<ceil_div(s->L1, 128), 128, 0, st>>>(s->L1, own.nq, bias_partial, g_b_d);
        NNUE_CHECK_LAUNCH("fold_partials_kernel");
        const long long n = 1LL * (s->P > s->F - 1 ? s->P : s->F - 1) * (s->L1 / 4);
        ft_bwd_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(*s, own.nq, partial, g_w_d, alias);
        NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_kernel");
        ft_bwd_dw_fold_last_kernel<<<ceil_div(s->L1 / 4, 32), 256, 0, st>>>(*s, alias, g_w_d);
        NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_last_kernel");
        return NNUE_OK;
    }
    if (!bits_t_d) return NNUE_ERR_INVALID_ARG;
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
        float *alias = partial + (size_t)d.ngroups * s->P * s->L1;
        const long long n = 1LL * (s->P > s->F - 1 ? s->P : s->F - 1) * (s->L1 / 4);
        if (n > 0) {
            ft_bwd_dw_fold_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(*s, d.ngroups, partial, g_w_d, alias);
            NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_kernel");
        }
        ft_bwd_dw_fold_last_kernel<<<ceil_div(s->L1 / 4, 32), 256, 0, st>>>(*s, alias, g_w_d);
        NNUE_CHECK_LAUNCH("ft_bwd_dw_fold_last_kernel");
    }
    return NNUE_OK;
}

int nnue_ft_bwd_dval(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *g_ft_d,
                     const float *xpad_d, const float *thr_d, float *dval_d, float *g_thr_d, void *workspace_d,
                     size_t workspace_bytes, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !g_ft_d || !xpad_d || !thr_d || !dval_d || !g_thr_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_ft_bwd_dval(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ft_umma_ok(*s)) {  // tcgen05 contraction for the dense masked rows, then the threshold-gradient reduction
        const int rc = launch_ft_bwd_gbin_umma(*s, bits_s_d, ft_w_d, g_ft_d, workspace_d, dval_d, st);
        if (rc != NNUE_OK) return rc;
        float *part = reinterpret_cast<float *>(static_cast<char *>(workspace_d) + ws_ft_gbin_umma(*s));
        const int nch = thr_chunks(*s), rpc = ceil_div(s->B, nch);
        thr_grad_kernel<<<dim3(s->C, nch), 256, 0, st>>>(*s, dval_d, xpad_d, thr_d, part, rpc);
        NNUE_CHECK_LAUNCH("thr_grad_kernel");
        fold_partials_kernel<<<ceil_div(s->C, 128), 128, 0, st>>>(s->C, nch, part, g_thr_d);
        NNUE_CHECK_LAUNCH("fold_partials_kernel");
        return NNUE_OK;
    }
    float *thr_partial = static_cast<float *>(workspace_d);
    const ColPlan cp = col_plan(s->L1);
    if (ft_gather_dval_ok(*s) && !use_staging(*s, cp.nchunks, s->B)) {  // rows staged by bulk TMA, dots reduced with shuffles
        const int rc = launch_ft_gather_dval(*s, bits_s_d, ft_w_d, g_ft_d, xpad_d, thr_d, dval_d, thr_partial, st);
        if (rc != NNUE_OK) return rc;
        fold_partials_kernel<<<ceil_div(s->C, 128), 128, 0, st>>>(s->C, ft_gather_dval_grid(*s), thr_partial, g_thr_d);
        NNUE_CHECK_LAUNCH("fold_partials_kernel");
        return NNUE_OK;
    }
    const int grid = dval_grid(*s);
    if (!cp.LPR) {
        ft_bwd_dval_generic_kernel<<<grid, kFtThreads, (size_t)(kFtThreads / 32) * s->C * 4, st>>>(
            *s, bits_s_d, ft_w_d, g_ft_d, xpad_d, thr_d, dval_d, thr_partial);
        NNUE_CHECK_LAUNCH("ft_bwd_dval_generic_kernel");
    } else {
        const size_t staged_smem =
            128 + align_up((size_t)(kDvalStagedThreads / 32) * s->C * 4, 128) + (size_t)s->F * s->L1 * 4;
        const bool staged = use_staging(*s, cp.nchunks, s->B) && staged_smem <= (size_t)max_dyn_smem();
        int rc = NNUE_OK;
        NNUE_DISPATCH_LPR(cp.LPR, rc = launch_ft_bwd_dval<LPR>(*s, bits_s_d, ft_w_d, g_ft_d, xpad_d, thr_d, dval_d,
                                                              thr_partial, staged ? kNumSMs : grid, staged,
                                                              cp.nchunks, st));
        if (rc != NNUE_OK) return rc;
        if (staged) {  // the staged kernel always runs kNumSMs CTAs
            fold_partials_kernel<<<ceil_div(s->C, 128), 128, 0, st>>>(s->C, kNumSMs, thr_partial, g_thr_d);
            NNUE_CHECK_LAUNCH("fold_partials_kernel");
            return NNUE_OK;
        }
    }
    fold_partials_kernel<<<ceil_div(s->C, 128), 128, 0, st>>>(s->C, grid, thr_partial, g_thr_d);
    NNUE_CHECK_LAUNCH("fold_partials_kernel");
    return NNUE_OK;
}

}  // extern "C"
