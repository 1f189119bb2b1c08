// ft_gather.cu -- the index-driven feature transformer for tables that live in HBM / L2 (nnue.py:686-710):
// coalesced row gather staged through shared memory by bulk TMA copies, for the forward accumulate and for the value
// gradient (row . g_ft dot products with warp-shuffle reductions).
//
// A persistent CTA owns a column slab of CS floats and walks (sample, slab) units.  One producer warp reads the
// sample's bitmask, and for every group of SLOTS consecutive positions that has an active one it claims the next stage
// of a shared-memory ring and issues one cp.async.bulk (UBLKCP) per ACTIVE position: CS * 4 contiguous bytes of that
// table row land in consecutive slots of the stage, completion on the stage's mbarrier (expect_tx = rows * slot bytes).
// The consumer warps wait on the barrier, add the slots in ascending position order (forward) or dot them with the
// sample's g_ft slab held in registers (value gradient), and hand the stage back.  The ring keeps ST stages of up to
// SLOTS rows in flight per SM -- ~80 KB on average at the 0.43 density of the reference model, which is what 6.5 TB/s
// times ~1 us of HBM latency needs per SM -- without tying up a register per byte in flight, and it runs across unit
// boundaries (a zero-row "terminator" stage ends each unit), so the pipeline never drains.
//
// No index list is materialised: the active set is consumed straight off the bitmask words.  Determinism: every sum has
// a fixed order (positions ascending, four interleaved accumulators per column; lanes then warps for the dots).
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr float kGaSharp = 10.0f;  // STE sharpness k (nnue.py:41)

__device__ __forceinline__ float4 ga_lds_f4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// shared-memory header: full[ST] | empty[ST] | meta[ST] (uint4: group bits, CHW position of the group's bit 0, sample)
template <int ST>
struct GaRing {
    uint64_t *full, *empty;
    uint4 *meta;
    unsigned char *stages;
    __device__ __forceinline__ explicit GaRing(unsigned char *smem) {
        full = reinterpret_cast<uint64_t *>(smem);
        empty = full + ST;
        meta = reinterpret_cast<uint4 *>(empty + ST);
        stages = smem + 1024;
    }
};

// The producer warp: for every unit of this CTA walk the bitmask in groups of SLOTS positions and stage the active rows.
// Shared by both kernels.  UNIT_END: a terminator stage (bits = 0) closes each unit (the forward writes its sums then);
// `exit_stages` more terminators follow the last unit (the value gradient's warps each consume one and leave).
template <int CS, int SLOTS, int ST, bool UNIT_END = true>
__device__ __forceinline__ void ga_produce(const nnue_shape &s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                                           GaRing<ST> &rg, int n_slabs, long long units, int n_ranges = 1, int exit_stages = 0) {
    constexpr uint32_t kSlot = CS * 4, kStage = SLOTS * kSlot;
    constexpr int GPW = 32 / SLOTS;  // groups per bitmask word
    const int lane = threadIdx.x & 31;
    const int cells = s.Gh * s.Gw, last_row = s.F - 1;
    uint32_t seq = 0;
    int b = 0;
    auto claim = [&](uint32_t gbits, int base) {  // warp-uniform arguments
        const uint32_t st = seq % ST;
        if (seq >= ST) mbar_wait(&rg.empty[st], ((seq / ST) - 1) & 1);
        if (lane == 0) {
            rg.meta[st] = make_uint4(gbits, (uint32_t)base, (uint32_t)b, 0u);
            if (gbits) mbar_arrive_expect_tx(&rg.full[st], (uint32_t)__popc(gbits) * kSlot);
            else mbar_arrive(&rg.full[st]);
        }
        __syncwarp();
        ++seq;
        return st;
    };
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        // unit = (sample, word range, slab): small batches split a sample's bitmask words over several CTAs
        const int slab = (int)(u % n_slabs), range = (int)((u / n_slabs) % n_ranges);
        b = (int)(u / ((long long)n_slabs * n_ranges));
        const float *wcol = w + (size_t)slab * CS;
        const uint32_t *brow = bits_s + (size_t)b * s.NW;
        const int wpr = ceil_div(s.NW, n_ranges), w_lo = range * wpr, w_hi = min(s.NW, w_lo + wpr);
        for (int w0 = w_lo; w0 < w_hi; w0 += 32) {
            const uint32_t mine = w0 + lane < w_hi ? __ldg(brow + w0 + lane) : 0u;
            uint32_t nonzero = __ballot_sync(kFull, mine != 0u);
            while (nonzero) {
                const int jw = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                const uint32_t word = __shfl_sync(kFull, mine, jw);
                const int widx = w0 + jw;
                const int wbase = (widx / s.CW) * cells + (widx % s.CW) * 32;  // CHW position of bit 0 of the word
#pragma unroll
                for (int g = 0; g < GPW; ++g) {
                    const uint32_t gbits = SLOTS == 32 ? word : (word >> (g * SLOTS)) & ((1u << (SLOTS & 31)) - 1u);
                    if (!gbits) continue;  // warp-uniform
                    const uint32_t st = claim(gbits, wbase + g * SLOTS);
                    if (lane < SLOTS && ((gbits >> lane) & 1u)) {
                        const int slot = __popc(gbits & ((1u << lane) - 1u));
                        const int row = min(wbase + g * SLOTS + lane, last_row);  // clamp of nnue.py:701
                        tma_bulk_g2s(rg.stages + st * kStage + (uint32_t)slot * kSlot, wcol + (size_t)row * s.L1, kSlot, &rg.full[st]);
                    }
                }
            }
        }
        if (UNIT_END) claim(0u, 0);  // terminator: the unit is complete
    }
    for (int i = 0; i < exit_stages; ++i) claim(0u, 0);
}

template <int ST>
__device__ __forceinline__ void ga_ring_init(GaRing<ST> &rg, int consumer_warps) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) {
            mbar_init(&rg.full[i], 1);
            mbar_init(&rg.empty[i], consumer_warps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// ---- forward: out[b, slab] = bias + sum over active p of W[min(p, F-1), slab] ------------------------------------------
// consumers: CS / 4 threads, thread t owns columns 4t .. 4t+3 of the slab
template <int CS, int SLOTS, int ST>
__global__ void __launch_bounds__(32 + CS / 4, 1)
ft_gather_fwd_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                     const float *__restrict__ bias, float *__restrict__ out, int n_slabs, int n_ranges,
                     float *__restrict__ partial) {
    constexpr uint32_t kSlot = CS * 4, kStage = SLOTS * kSlot;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    GaRing<ST> rg(smem_raw);
    ga_ring_init<ST>(rg, CS / 128);
    const long long units = 1LL * s.B * n_slabs * n_ranges;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        ga_produce<CS, SLOTS, ST>(s, bits_s, w, rg, n_slabs, units, n_ranges);
        return;
    }
    const int t = threadIdx.x - 32;
    const uint32_t tbase = smem_u32(rg.stages) + (uint32_t)t * 16u;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    uint32_t seq = 0;
    long long u = blockIdx.x;
    while (u < units) {
        const uint32_t st = seq % ST;
        mbar_wait(&rg.full[st], (seq / ST) & 1);
        const uint32_t gbits = rg.meta[st].x;
        if (gbits) {
            const int c = __popc(gbits);
            const uint32_t base = tbase + st * kStage;
            int r = 0;
            for (; r + 4 <= c; r += 4) {
                const float4 v0 = ga_lds_f4(base + (uint32_t)r * kSlot), v1 = ga_lds_f4(base + (uint32_t)(r + 1) * kSlot);
                const float4 v2 = ga_lds_f4(base + (uint32_t)(r + 2) * kSlot), v3 = ga_lds_f4(base + (uint32_t)(r + 3) * kSlot);
                a0 = f4_add(a0, v0); a1 = f4_add(a1, v1); a2 = f4_add(a2, v2); a3 = f4_add(a3, v3);
            }
            if (r < c) a0 = f4_add(a0, ga_lds_f4(base + (uint32_t)r * kSlot));
            if (r + 1 < c) a1 = f4_add(a1, ga_lds_f4(base + (uint32_t)(r + 1) * kSlot));
            if (r + 2 < c) a2 = f4_add(a2, ga_lds_f4(base + (uint32_t)(r + 2) * kSlot));
        } else {
            const int col = (int)(u % n_slabs) * CS + t * 4;
            const float4 sum = f4_add(f4_add(a0, a1), f4_add(a2, a3));
            if (n_ranges == 1) {
                const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + col));
                *reinterpret_cast<float4 *>(out + (size_t)(u / n_slabs) * s.L1 + col) = f4_add(bv, sum);
            } else {  // partial[b][range][L1]: (u / n_slabs) = b * n_ranges + range
                *reinterpret_cast<float4 *>(partial + (size_t)(u / n_slabs) * s.L1 + col) = sum;
            }
            a0 = a1 = a2 = a3 = make_float4(0.f, 0.f, 0.f, 0.f);
            u += gridDim.x;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty[st]);
        ++seq;
    }
}

// ---- value gradient: dval[b, p] = <W[min(p, F-1)], g_ft[b]> at active p, fused with the threshold gradient ---------------
// The CTA takes whole rows (CS = L1).  Stages are dealt to the consumer warps round-robin (stage seq -> warp seq % NCW,
// ST = NCW), so NCW stages are being consumed at once and the latency of a stage's tail (the activation load, the
// sigmoid, the barrier round trip) is covered by the other warps.  For every row of its stage a warp reads 4 L1 bytes as
// LDS.128 against the sample's g_ft held in registers (L1 / 32 floats per lane, reloaded when the sample changes) and
// combines the 32 partial dots with five shuffles; lane r then finishes row r: dval store, straight-through threshold
// gradient (nnue.py:36-52) into a per-lane sum that is folded per channel.  Per-CTA partials [grid][C] -> fold_partials_kernel.
template <int L1T, int SLOTS, int NCW>
__global__ void __launch_bounds__(32 + 32 * NCW, 1)
ft_gather_dval_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ w,
                      const float *__restrict__ g_ft, const float *__restrict__ xpad, const float *__restrict__ thr,
                      float *__restrict__ dval, float *__restrict__ thr_partial, int n_ranges) {
    constexpr int ST = NCW;
    constexpr uint32_t kSlot = L1T * 4, kStage = SLOTS * kSlot;
    constexpr int V = L1T / 128;  // float4s per lane per row
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    GaRing<ST> rg(smem_raw);
    ga_ring_init<ST>(rg, 1);
    const long long units = 1LL * s.B * n_ranges;  // unit = (sample, word range): small batches spread a sample over the SMs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        ga_produce<L1T, SLOTS, ST, false>(s, bits_s, w, rg, 1, units, n_ranges, NCW);
        return;
    }
    const int k = warp - 1, cells = s.Gh * s.Gw;
    const uint32_t lbase = smem_u32(rg.stages) + (uint32_t)lane * 16u;
    float *my_thr = reinterpret_cast<float *>(rg.stages + (size_t)ST * kStage) + (size_t)k * s.C;  // [NCW][C] behind the ring
    for (int c = lane; c < s.C; c += 32) my_thr[c] = 0.0f;
    __syncwarp();
    float4 g[V];
    float thr_acc = 0.0f;  // per lane: rows this lane finished for channel `cur_c`
    int cur_c = -1, cur_b = -1;
    auto flush = [&]() {
        const float v = warp_sum(thr_acc);
        if (lane == 0 && cur_c >= 0) my_thr[cur_c] += v;
        thr_acc = 0.0f;
    };
    for (uint32_t seq = (uint32_t)k;; seq += NCW) {
        const uint32_t st = seq % ST;  // (= k: ST == NCW)
        mbar_wait(&rg.full[st], (seq / ST) & 1);
        const uint4 m = rg.meta[st];
        if (!m.x) break;  // exit stage
        const int b = (int)m.z;
        if (b != cur_b) {
#pragma unroll
            for (int v = 0; v < V; ++v) g[v] = __ldg(reinterpret_cast<const float4 *>(g_ft + (size_t)b * L1T) + v * 32 + lane);
            cur_b = b;
        }
        const int ch = (int)m.y / cells, cell0 = (int)m.y % cells;  // a bitmask word, hence a group, belongs to one channel
        if (ch != cur_c) {  // warp-uniform
            flush();
            cur_c = ch;
        }
        const int c = __popc(m.x);
        float mine = 0.0f;
        for (int r = 0; r < c; ++r) {
            const uint32_t base = lbase + st * kStage + (uint32_t)r * kSlot;
            float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 x = ga_lds_f4(base + (uint32_t)v * 512u);
                d0 = fmaf(x.x, g[v].x, d0); d1 = fmaf(x.y, g[v].y, d1);
                d0 = fmaf(x.z, g[v].z, d0); d1 = fmaf(x.w, g[v].w, d1);
            }
            const float d = warp_sum(d0 + d1);
            if (lane == r) mine = d;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty[st]);  // the rows have been read: the producer may refill the stage
        if (lane < c) {  // lane r finishes row r
            const int bit = __fns(m.x, 0, lane + 1);  // position of the r-th active bit of the group
            const size_t at = (size_t)b * s.PP + (size_t)ch * s.CW * 32 + cell0 + bit;
            dval[at] = mine;
            const float z = kGaSharp * (__ldg(xpad + at) - __ldg(thr + ch));
            const float sg = __fdividef(1.0f, 1.0f + __expf(-z));
            thr_acc = fmaf(-mine, kGaSharp * sg * (1.0f - sg), thr_acc);
        }
    }
    flush();
    // CTA partial: warps in fixed order
    asm volatile("bar.sync 1, %0;" ::"r"(32 * NCW) : "memory");
    const float *all = reinterpret_cast<const float *>(rg.stages + (size_t)ST * kStage);
    for (int c = threadIdx.x - 32; c < s.C; c += 32 * NCW) {
        float v = 0.0f;
        for (int q = 0; q < NCW; ++q) v += all[(size_t)q * s.C + c];
        thr_partial[(size_t)blockIdx.x * s.C + c] = v;
    }
}

// out[b] = bias + sum over word ranges of partial[b][range].  CTA = (sample, 8 float4 columns): 32 slices of ranges
// summed side by side (slice q takes ranges q, q + 32, ...), then combined in slice order -- a fixed order for a given shape.
__global__ void __launch_bounds__(256)
ft_gather_fold_kernel(int L1, int n_ranges, const float *__restrict__ partial, const float *__restrict__ bias,
                      float *__restrict__ out) {
    __shared__ float4 red[32][8];
    const int b = blockIdx.y, c4 = blockIdx.x * 8 + (threadIdx.x & 7), q = threadIdx.x >> 3;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < L1 / 4) {
        const float4 *src = reinterpret_cast<const float4 *>(partial + (size_t)b * n_ranges * L1) + c4;
#pragma unroll 4
        for (int r = q; r < n_ranges; r += 32) acc = f4_add(acc, __ldg(src + (size_t)r * (L1 / 4)));
    }
    red[q][threadIdx.x & 7] = acc;
    __syncthreads();
    if (q == 0 && c4 < L1 / 4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        for (int k = 0; k < 32; ++k) v = f4_add(v, red[k][threadIdx.x & 7]);
        reinterpret_cast<float4 *>(out + (size_t)b * L1)[c4] = v;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// word ranges per sample: when the batch alone cannot occupy the SMs a sample's bitmask words are split over several
// CTAs (at config I one sample gathers 115 MB of rows: all 148 SMs pull on it)
int ft_gather_ranges(const nnue_shape &s, int cs) {
    const long long base = 1LL * s.B * (s.L1 / cs);
    if (base >= kNumSMs) return 1;
    int r = (int)(1LL * get_option(kOptGatherUnits) * kNumSMs / base);  // (units per SM: 1 measured best at one sample -- 36 / 38 / 42 / 44 us for 1 / 2 / 3 / 4: the fold grows with the partials)
    const int max_r = ceil_div(s.NW, 2);  // at least 2 words (64 positions) per range
    if (r > max_r) r = max_r;
    return r < 1 ? 1 : r;
}

template <int CS, int SLOTS, int ST>
static int launch_gather_fwd_t(const nnue_shape &s, const uint32_t *bits, const float *w, const float *bias, float *out,
                               float *partial, cudaStream_t st) {
    constexpr size_t smem = 1024 + (size_t)ST * SLOTS * CS * 4;
    static_assert(smem <= kMaxSmemOptin, "ring too large");
    auto k = ft_gather_fwd_kernel<CS, SLOTS, ST>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_slabs = s.L1 / CS;
    const int n_ranges = partial ? ft_gather_ranges(s, CS) : 1;
    const long long units = 1LL * s.B * n_slabs * n_ranges;
    const int grid = (int)(units < kNumSMs ? units : kNumSMs);
    k<<<grid, 32 + CS / 4, smem, st>>>(s, bits, w, bias, out, n_slabs, n_ranges, partial);
    NNUE_CHECK_LAUNCH("ft_gather_fwd_kernel");
    if (n_ranges > 1) {
        ft_gather_fold_kernel<<<dim3(ceil_div(s.L1 / 4, 8), s.B), 256, 0, st>>>(s.L1, n_ranges, partial, bias, out);
        NNUE_CHECK_LAUNCH("ft_gather_fold_kernel");
    }
    return NNUE_OK;
}

bool ft_gather_ok(const nnue_shape &s) { return s.L1 % 128 == 0 && s.L1 >= 128 && get_option(kOptFtGather); }

int ft_gather_slab(const nnue_shape &s) {
    const int variant = get_option(kOptFtGatherVariant);  // slab width in columns: 0 = widest that divides L1
    int cs = variant ? variant : (s.L1 % 1024 == 0 ? 1024 : s.L1 % 512 == 0 ? 512 : s.L1 % 256 == 0 ? 256 : 128);
    if ((cs != 128 && cs != 256 && cs != 512 && cs != 1024) || s.L1 % cs) cs = 128;
    return cs;
}
// scratch for the word-range partials of small batches: [B][ranges][L1] floats (0 when the batch fills the SMs)
size_t ws_ft_gather_fwd(const nnue_shape &s) {
    if (!ft_gather_ok(s)) return 0;
    const int r = ft_gather_ranges(s, ft_gather_slab(s));
    return r > 1 ? (size_t)s.B * r * s.L1 * 4 : 0;
}

int launch_ft_gather_fwd(const nnue_shape &s, const uint32_t *bits, const float *w, const float *bias, float *out,
                         void *workspace, size_t workspace_bytes, cudaStream_t st) {
    float *partial = workspace && workspace_bytes >= ws_ft_gather_fwd(s) && ws_ft_gather_fwd(s) ? static_cast<float *>(workspace) : nullptr;
    switch (ft_gather_slab(s)) {
        case 1024: return launch_gather_fwd_t<1024, 8, 6>(s, bits, w, bias, out, partial, st);
        case 512: return launch_gather_fwd_t<512, 16, 6>(s, bits, w, bias, out, partial, st);
        case 256: return launch_gather_fwd_t<256, 32, 6>(s, bits, w, bias, out, partial, st);
        default: return launch_gather_fwd_t<128, 32, 12>(s, bits, w, bias, out, partial, st);
    }
}

template <int L1T, int SLOTS, int NCW>
static int launch_gather_dval_t(const nnue_shape &s, const uint32_t *bits, const float *w, const float *g_ft, const float *xpad,
                                const float *thr, float *dval, float *thr_partial, int grid, cudaStream_t st) {
    const size_t smem = 1024 + (size_t)NCW * SLOTS * L1T * 4 + (size_t)NCW * s.C * 4;
    if (smem > kMaxSmemOptin) return NNUE_ERR_UNSUPPORTED;
    auto k = ft_gather_dval_kernel<L1T, SLOTS, NCW>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 32 + 32 * NCW, smem, st>>>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, ft_gather_ranges(s, s.L1));
    NNUE_CHECK_LAUNCH("ft_gather_dval_kernel");
    return NNUE_OK;
}

bool ft_gather_dval_ok(const nnue_shape &s) {
    return ft_gather_ok(s) && (s.L1 == 128 || s.L1 == 256 || s.L1 == 512 || s.L1 == 1024) && s.C <= 512;
}
int ft_gather_dval_grid(const nnue_shape &s) {
    const long long units = 1LL * s.B * ft_gather_ranges(s, s.L1);
    return (int)(units < kNumSMs ? units : kNumSMs);
}

int launch_ft_gather_dval(const nnue_shape &s, const uint32_t *bits, const float *w, const float *g_ft, const float *xpad,
                          const float *thr, float *dval, float *thr_partial, cudaStream_t st) {
    const int grid = ft_gather_dval_grid(s);
    switch (s.L1) {
        // 16 KB stages, twelve of them, one consumer warp each
        case 1024: return launch_gather_dval_t<1024, 4, 12>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, grid, st);
        case 512: return launch_gather_dval_t<512, 8, 12>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, grid, st);
        case 256: return launch_gather_dval_t<256, 16, 12>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, grid, st);
        default: return launch_gather_dval_t<128, 32, 12>(s, bits, w, g_ft, xpad, thr, dval, thr_partial, grid, st);
    }
}

}  // namespace nnue
