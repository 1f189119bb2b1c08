// input_bwd_rows.cu -- conv weight gradient + straight-through threshold gradient for images that do not fit the
// whole-image staging of conv_bwd_kernel (ImageNet-shaped input, many channels):
//   g_conv_w[c][t] = sum over (b, cell) of g_bin[b, c, cell] * patch[b, cell][t]              (nnue.py:640 backward)
//   g_thr[c]       = -sum g_bin * k * sig * (1 - sig),  sig = sigmoid(k (x - thr[c])), k = 10   (nnue.py:36-52)
// with g_bin = dval at active positions, 0 elsewhere, and x the pre-threshold activations the forward stored.
//
// Work unit = (sample b, cell word j): 32 consecutive cells of the conv raster, one per lane.  For every raster row
// the word touches, ONE tensor-map TMA copy brings the three image rows x three planes the row's 3x3 taps read into a
// ring stage -- box = W columns x 3 rows x 3 planes, out-of-image rows ZERO-FILLED by the TMA unit, which is the conv's
// vertical padding; the horizontal padding taps (column -1 of the first cell of a raster row, column W where the last
// cell reaches it) are two per-lane flags.  With the row width a compile-time constant the tap loop has no address
// arithmetic (27 LDS with immediate offsets from one per-lane base).  With an odd conv
// stride the 32 lanes of a tap hit 32 different banks, where the direct global gather of extract_bwd_kernel spends one
// L1 wavefront per touched line and repeats it for every channel pair (ncu: 7 wavefronts per tap load; 3.0 ms of the
// 12.9 ms step at SURVEY config I).  A warp owns CH = 4 channels for the whole kernel and keeps their 4 x 27
// accumulators (+ the threshold gradient) in registers; a CTA's 8 warps cover 32 channels and share the staged rows.
// The unit's g_bin / activation segments (128 bytes per channel) and bitmask words ride in the same ring stage, copied
// by cp.async with completion on the stage's mbarrier, so ~10 units are in flight per SM and nothing waits on a global
// load.  Per-lane geometry (which staged row block, which column) comes from a table built once in shared memory, and
// the (sample, word) cursors advance incrementally: the first version spent 740 instructions per warp and unit, 15 % of
// them FFMA (ncu: profiles/r2_ncu_rows_v2_summary.txt).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr float kRowsSharp = 10.0f;  // nnue.py:41

// Ampere-style asynchronous copies (LDGSTS) whose completion is reported to an mbarrier: used for the many short
// segments (128 bytes of g_bin / activations per channel and unit) for which one bulk-TMA copy each would cost more
// issue slots than the data is worth
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the executing thread's arrival on `bar` happens when all its earlier cp.async have landed (no pending-count increment)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rows_tma_3d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// WBT: box (= image) width in floats when known at compile time (224), 0 = read it from the plan
template <int WBT>
__global__ void __launch_bounds__(kRowsWarps * 32, 1)
conv_bwd_rows_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, const float *__restrict__ dval,
                     const float *__restrict__ xpad, const float *__restrict__ thr, float *__restrict__ partial,
                     const RowsPlan pl, const __grid_constant__ CUtensorMap tmap) {
    constexpr int CH = kRowsCH, WARPS = kRowsWarps;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kInMaxStages;
    float *red = reinterpret_cast<float *>(smem_raw + kInHeader);                          // [WARPS][CH][28]
    uint32_t *rowtab = reinterpret_cast<uint32_t *>(red + WARPS * CH * 28);                // [CW]: first raster row | rows << 16
    uint32_t *geo = rowtab + s.CW;  // [CW][32]: float offset of a lane's centre-column tap (row 0, plane 0) | left edge << 30 | right edge << 31
    float *stages = reinterpret_cast<float *>(smem_raw + pl.stage_off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cb = blockIdx.x % pl.NCB, q = blockIdx.x / pl.NCB;  // channel block, unit stream
    const long long units = 1LL * s.B * s.CW;
    const long long n_mine = units > q ? (units - q + pl.nq - 1) / pl.nq : 0;
    const int WB = WBT ? WBT : pl.WB, TILE = (9 * WB + 31) / 32 * 32, cells = s.Gh * s.Gw;  // (boxes land 128-byte aligned)

    if (threadIdx.x == 0) {
        for (int i = 0; i < pl.ST; ++i) {
            mbar_init(&full[i], 1 + WARPS * 32);  // thread 0's expect_tx arrival + every thread's cp.async arrival
            mbar_init(&empty[i], WARPS);
        }
        mbar_fence_init();
    }
    for (int j = threadIdx.x; j < s.CW; j += blockDim.x) {
        const int first = (j * 32) / s.Gw, last = min(j * 32 + 31, cells - 1) / s.Gw;
        rowtab[j] = (uint32_t)first | ((uint32_t)(last - first + 1) << 16);
    }
    for (int i = threadIdx.x; i < s.CW * 32; i += blockDim.x) {
        const int cell = min(i, cells - 1), oy0 = ((i >> 5) * 32) / s.Gw;  // (cells past the raster carry no bits: any address will do)
        const int ox = cell % s.Gw;
        geo[i] = (uint32_t)((cell / s.Gw - oy0) * TILE + ox * s.stride) | (ox == 0 ? 1u << 30 : 0u) |
                 (ox * s.stride + 1 >= s.W ? 1u << 31 : 0u);
    }
    __syncthreads();

    // (sample, word) of unit ii of this stream, advanced without divisions
    const int db = pl.nq / s.CW, dj = pl.nq % s.CW;
    int pb = q / s.CW, pj = q % s.CW;  // producer cursor
    int cj = pj;                        // consumer cursor (the consumers never need the sample)
    auto advance = [&](int &b, int &j) {
        j += dj; b += db;
        if (j >= s.CW) { j -= s.CW; ++b; }
    };

    // Producer duty is shared by all threads: unit ii goes into stage ii % ST once every warp has released it.  Thread 0
    // arms the barrier with the bytes of the unit's row boxes; warp r's lane 0 issues the box of raster row r; every
    // thread copies 16 bytes of g_bin and of the activations (its channel = thread / 8), warp 0 also the 32 bitmask words.
    // (A 9-warp CTA with a dedicated producer warp would be allocated registers for 12 warps and cap the kernel at 168
    // registers -- below its 4 x 27 accumulators.)
    const int ahead = pl.ST - 2;  // units in flight; a refill waits for a stage released two iterations ago
    const int my_ch = threadIdx.x >> 3, my_seg = threadIdx.x & 7;
    const int my_c = cb * (WARPS * CH) + my_ch;
    const int bit_c = cb * (WARPS * CH) + (int)threadIdx.x;
    int p_st = 0;          // producer cursor: stage of the next unit to stage ...
    uint32_t p_ph = 1u;    // ... and the parity of the release it waits for (no wait during the first pass over the ring)
    bool p_wrapped = false;
    auto produce = [&]() {
        const int st = p_st;
        if (p_wrapped) mbar_wait(&empty[st], p_ph);
        if (++p_st == pl.ST) { p_st = 0; p_ph ^= 1u; p_wrapped = true; }
        float *stg = stages + (size_t)st * pl.stage_floats;
        const uint32_t rt = rowtab[pj];
        const int oy0 = (int)(rt & 0xFFFFu), nr = (int)(rt >> 16);
        if (threadIdx.x == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)(nr * 9 * WB) * 4u);  // (zero-filled bytes count)
        if (lane == 0 && warp < nr) rows_tma_3d(stg + warp * TILE, &tmap, 0, (oy0 + warp) * s.stride - 1, 3 * pb, &full[st]);
        if (threadIdx.x < 256 && my_c < s.C) {
            const size_t at = (size_t)pb * s.PP + (size_t)(my_c * s.CW + pj) * 32 + my_seg * 4;
            cp_async16(stg + pl.dx_off + my_ch * 32 + my_seg * 4, dval + at);
            cp_async16(stg + pl.dx_off + 1024 + my_ch * 32 + my_seg * 4, xpad + at);
        }
        if (threadIdx.x < 32 && bit_c < s.C)
            cp_async4(stg + pl.dx_off + 2048 + threadIdx.x, bits_s + (size_t)pb * s.NW + (size_t)bit_c * s.CW + pj);
        cp_async_arrive_noinc(&full[st]);
        advance(pb, pj);
    };
    for (long long ii = 0; ii < ahead && ii < n_mine; ++ii) produce();

    // ---- consumers ----
    const int c0 = cb * (WARPS * CH) + warp * CH;
    float acc[CH][27], dth[CH], thr_c[CH];
    bool chan_ok[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        chan_ok[k] = c0 + k < s.C;
        thr_c[k] = __ldg(thr + min(c0 + k, s.C - 1));
        dth[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < 27; ++t) acc[k][t] = 0.0f;
    }
    int st = 0;
    uint32_t ph = 0;
    for (long long i = 0; i < n_mine; ++i) {
        if (i + ahead < n_mine) produce();
        __syncwarp();
        const uint32_t gword = geo[cj * 32 + lane];
        int dummy = 0;
        advance(dummy, cj);
        mbar_wait(&full[st], ph);
        const float *stg = stages + (size_t)st * pl.stage_floats;
        const float *sdx = stg + pl.dx_off;
        float g[CH], x[CH];
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const uint32_t wk = chan_ok[k] ? __float_as_uint(sdx[2048 + warp * CH + k]) : 0u;
            const bool on = (wk >> lane) & 1u;
            g[k] = on ? sdx[(warp * CH + k) * 32 + lane] : 0.0f;
            x[k] = on ? sdx[1024 + (warp * CH + k) * 32 + lane] : 0.0f;
            any |= wk;
        }
        if (any) {  // warp-uniform
            const float *p = stg + (gword & 0x3FFFFFFFu);
            const bool left = (gword >> 30) & 1u, right = gword >> 31;
#pragma unroll
            for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float pt = p[(ic * 3 + kh) * WB + kw - 1];
                        if (kw == 0 && left) pt = 0.0f;   // horizontal padding (nnue.py:640, padding = 1)
                        if (kw == 2 && right) pt = 0.0f;
#pragma unroll
                        for (int k = 0; k < CH; ++k) acc[k][ic * 9 + kh * 3 + kw] = fmaf(g[k], pt, acc[k][ic * 9 + kh * 3 + kw]);
                    }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const float z = kRowsSharp * (x[k] - thr_c[k]);
                const float sgm = __fdividef(1.0f, 1.0f + __expf(-z));
                dth[k] = fmaf(-g[k], kRowsSharp * sgm * (1.0f - sgm), dth[k]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == pl.ST) { st = 0; ph ^= 1u; }
    }
    // warp reduction, then the CTA's partial [C][28] (27 taps + the threshold gradient); other channel blocks' rows are zero
#pragma unroll
    for (int k = 0; k < CH; ++k) {
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float v = warp_sum(acc[k][t]);
            if (lane == 0) red[(warp * CH + k) * 28 + t] = v;
        }
        const float v = warp_sum(dth[k]);
        if (lane == 0) red[(warp * CH + k) * 28 + 27] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s.C * 28; i += WARPS * 32) {
        const int c = i / 28, t = i % 28;
        const int lc = c - cb * (WARPS * CH);
        partial[(size_t)blockIdx.x * s.C * 28 + i] = (lc >= 0 && lc < WARPS * CH) ? red[lc * 28 + t] : 0.0f;
    }
}

// images [B][3][H][W] fp32 as a 3-D tensor (x, y, plane); box = W columns x 3 rows x 3 planes, zero fill for rows out of bounds
static bool make_rows_tmap(const float *images, const nnue_shape &s, int WB, CUtensorMap *tm) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn enc = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.B * 3};
    const cuuint64_t strides[2] = {(cuuint64_t)s.W * 4, (cuuint64_t)s.H * s.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)WB, 3, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(images), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_conv_bwd_rows(const nnue_shape &s, const RowsPlan &pl, const float *images, const uint32_t *bits_s, const float *dval,
                         const float *xpad, const float *thr, float *partial, cudaStream_t st) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (!make_rows_tmap(images, s, pl.WB, &tm)) return NNUE_ERR_UNSUPPORTED;
    auto k = pl.WB == 224 ? conv_bwd_rows_kernel<224> : conv_bwd_rows_kernel<0>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k<<<pl.grid, kRowsWarps * 32, pl.smem, st>>>(s, bits_s, dval, xpad, thr, partial, pl, tm);
    NNUE_CHECK_LAUNCH("conv_bwd_rows_kernel");
    return NNUE_OK;
}

}  // namespace nnue
