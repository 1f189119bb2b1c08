// input_bwd_fused.cu -- value gradient, conv-weight gradient and threshold gradient of the CIFAR-shaped step in ONE
// kernel: the value gradient g_bin[b, p] = bit(b, p) ? <W[min(p, F-1)], g_ft[b]> : 0 never leaves the chip.
//
// In round 1 ft_gbin_umma_kernel wrote the fp32 plane g_bin [B][PP] (67 MB at config D) and conv_bwd_kernel read it back
// 36 us later together with the images and the stored activations: 134 MB of HBM traffic and one kernel (+ its operand
// formatting) on the step's critical path for a quantity every consumer lane needs exactly once.  Here the persistent
// conv-gradient CTA computes it itself on the tensor cores, 32 samples at a time, and its consumer warps read their
// values straight out of tensor memory:
//
//   D_m [128 positions x 32 samples] = sum_k A_m [128 x 64] (table rows of channel m, three bf16 terms)
//                                            x B   [32 x 64]  (g_ft rows of the round's samples, three bf16 terms)
//
// for the C <= 8 channels m (one M tile per channel: CW = 4 cell words of 32 cells), six exact term-pair UMMAs
// (128 x 32 x 16) per 16-deep k step -- the same products in the same order as ft_gbin_umma_kernel.  TMEM holds two rounds
// (2 x C x 32 columns), so the tensor cores work one round ahead of the consumers.  The orientation is chosen so that a
// consumer lane finds its value where tcgen05.ld lets it look: consumer warp w = (channel group, cell word j) owns TMEM
// lanes 32 j .. 32 j + 31 (= w % 4), lane l = cell 32 j + l, and the sample is the COLUMN: one tcgen05.ld.32x32b.x1 per
// channel and sample, no shared-memory detour.
//
// Warp roles (12 warps; the consumer warp groups raise their register budget with setmaxnreg, the service group drops
// to 40): warps 0-7 consumers (4 channels x 27 accumulators each, as conv_bwd_kernel), warp 8 streams the samples'
// images (one swizzled tensor-map TMA copy each), activations and bitmask rows into a ring, warp 9 streams the operand
// tiles (the table's channel tiles through one 48 KB buffer, the rounds' g_ft tiles through two 12 KB buffers), warp 10
// owns tensor memory and issues the UMMAs.  Everything is handed over through mbarriers; tcgen05.commit releases the
// operand buffers and publishes a finished round.
//
// Shapes: 32 x 32 images, CW = 4 (97 .. 128 raster cells), C <= 8, L1 in {32, 64}, stored activations -- config D and
// its neighbours; other shapes keep the two-kernel path.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "plan.cuh"
#include "umma.cuh"

namespace nnue {

constexpr float kFuSharp = 10.0f;               // nnue.py:41
constexpr int kFuConsumers = 8, kFuWarps = 12;  // consumer warps, all warps
constexpr int kFuCH = 4;                        // channels per consumer warp
constexpr int kFuNS = 32;                       // samples per round (UMMA N)
constexpr int kFuSwzRows = 40, kFuSwzPlane = kFuSwzRows * 32;
constexpr uint32_t kFuImgBytes = 3 * kFuSwzPlane * 4;                    // 15360: box 32 x 40 x 3, zero fill included
constexpr uint32_t kFuStageBytes = 20480;                                // image | activations [1024] | bitmask [32] (1024-aligned)
constexpr uint32_t kFuXOff = kFuImgBytes, kFuBitsOff = kFuImgBytes + 4096;

__device__ __forceinline__ void fu_tma_3d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float fu_tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
    return __uint_as_float(r);
}

// g_ft [B][L1] fp32 -> the rounds' B tiles in the order the fused kernel consumes them: tile (q, r) holds the samples
// q + (32 r + i) nq, i < 32 (zero rows past the batch), as [L1 / 16][3 terms][32 x 16] bf16 in the canonical K-major
// layout.  A thread packs eight consecutive k of one row.
__global__ void fused_format_g_kernel(int B, int L1, int nq, int rounds, const float *__restrict__ g_ft, unsigned char *__restrict__ out) {
    const int kcs = L1 / 8, n_ks = L1 / 16;
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows = 1LL * nq * rounds * kFuNS;
    if (i >= rows * kcs) return;
    const long long r = i % rows;
    const int kc = (int)(i / rows);
    const int tile = (int)(r / kFuNS), rr = (int)(r % kFuNS);
    const int q = tile / rounds, rd = tile % rounds;
    const long long b = q + (long long)(kFuNS * rd + rr) * nq;
    float v[8];
    if (b < B) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(g_ft + (size_t)b * L1 + kc * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(g_ft + (size_t)b * L1 + kc * 8) + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    }
    uint4 o[3];
    split3x8(v, o);
    const int ks = kc >> 1;
    unsigned char *t0 = out + ((size_t)tile * n_ks + ks) * 3 * (kFuNS * 32) + (uint32_t)(kc & 1) * (kFuNS * 16) + (uint32_t)(rr >> 3) * 128 +
                        (uint32_t)(rr & 7) * 16;
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(t0 + (uint32_t)t * (kFuNS * 32)) = o[t];
}

struct FusedArgs {
    int nq, rounds, ST, n_ks;      // sample streams (= CTAs), rounds per stream, ring depth, k steps (L1 / 16)
    uint32_t a_bytes, b_bytes;     // one channel's table tiles, one round's g_ft tiles
    uint32_t a_off, b_off, stage_off;
};

__global__ void __launch_bounds__(kFuWarps * 32, 1)
conv_bwd_fused_kernel(const nnue_shape s, const float *__restrict__ xpad, const uint32_t *__restrict__ bits_s,
                      const unsigned char *__restrict__ wtiles, const unsigned char *__restrict__ gtiles,
                      const float *__restrict__ thr, float *__restrict__ partial, const FusedArgs fa,
                      const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);            // [16] ring: producer -> consumers
    uint64_t *empty = full + 16;                                        // [16] ring: consumers -> producer
    uint64_t *a_full = empty + 16, *a_empty = a_full + 1;               // table tile buffer
    uint64_t *b_full = a_empty + 1, *b_empty = b_full + 2;              // [2] g_ft tile buffers
    uint64_t *tm_full = b_empty + 2, *tm_empty = tm_full + 2;           // [2] accumulator rounds
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tm_empty + 2);
    float *red = reinterpret_cast<float *>(smem_raw + 512);             // [8][4][28]
    unsigned char *sa = smem_raw + fa.a_off, *sb = smem_raw + fa.b_off, *stages = smem_raw + fa.stage_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x;
    const int n_mine = s.B > q ? (s.B - q + fa.nq - 1) / fa.nq : 0;
    const int n_rounds = (n_mine + kFuNS - 1) / kFuNS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < fa.ST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kFuConsumers);
        }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
            mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], kFuConsumers);
        }
        mbar_fence_init();
    }
    if (warp == 10) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= kFuConsumers) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 8 && lane == 0) {
            // ---- ring producer: sample i of this stream -> stage i % ST ----
            for (int i = 0; i < n_mine; ++i) {
                const int st = i % fa.ST;
                if (i >= fa.ST) mbar_wait(&empty[st], ((i / fa.ST) - 1) & 1);
                const int b = q + i * fa.nq;
                unsigned char *stg = stages + (size_t)st * kFuStageBytes;
                mbar_arrive_expect_tx(&full[st], kFuImgBytes + 4096u + 128u);
                fu_tma_3d(stg, &tmap, 0, -1, 3 * b, &full[st]);
                tma_bulk_g2s(stg + kFuXOff, xpad + (size_t)b * s.PP, 4096u, &full[st]);
                tma_bulk_g2s(stg + kFuBitsOff, bits_s + (size_t)b * s.NW, 128u, &full[st]);
            }
        } else if (warp == 9 && lane == 0) {
            // ---- operand producer: per round the g_ft tiles, then the table's channel tiles one by one ----
            int a_use = 0;
            for (int r = 0; r < n_rounds; ++r) {
                const int buf = r & 1;
                if (r >= 2) mbar_wait(&b_empty[buf], ((r >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&b_full[buf], fa.b_bytes);
                tma_bulk_g2s(sb + (size_t)buf * fa.b_bytes, gtiles + ((size_t)q * fa.rounds + r) * fa.b_bytes, fa.b_bytes, &b_full[buf]);
                for (int m = 0; m < s.C; ++m, ++a_use) {
                    if (a_use >= 1) mbar_wait(a_empty, (a_use - 1) & 1);
                    mbar_arrive_expect_tx(a_full, fa.a_bytes);
                    const uint32_t half = fa.a_bytes / 2;  // (bulk copies of at most 32 KB)
                    tma_bulk_g2s(sa, wtiles + (size_t)m * fa.a_bytes, half, a_full);
                    tma_bulk_g2s(sa + half, wtiles + (size_t)m * fa.a_bytes + half, half, a_full);
                }
            }
        } else if (warp == 10) {
            if (lane == 0) {
                // ---- issuer: round r -> accumulator buffer r & 1, columns buf * 256 + 32 m ----
                constexpr uint32_t idesc = umma_idesc(128, kFuNS);
                int a_use = 0;
                for (int r = 0; r < n_rounds; ++r) {
                    const int buf = r & 1;
                    if (r >= 2) {
                        mbar_wait(&tm_empty[buf], ((r >> 1) - 1) & 1);
                        tcgen05_fence_after();
                    }
                    mbar_wait(&b_full[buf], (r >> 1) & 1);
                    const uint32_t b_base = smem_u32(sb + (size_t)buf * fa.b_bytes), a_base = smem_u32(sa);
                    for (int m = 0; m < s.C; ++m, ++a_use) {
                        mbar_wait(a_full, a_use & 1);
                        tcgen05_fence_after();
                        const uint32_t d = tmem + (uint32_t)(buf * 256 + m * kFuNS);
                        for (int ks = 0; ks < fa.n_ks; ++ks)
#pragma unroll
                            for (int tw = 0; tw < 3; ++tw)      // (term of W, term of g) with tw + tg <= 2
#pragma unroll
                                for (int tg = 0; tw + tg < 3; ++tg)
                                    umma_bf16(d, umma_smem_desc(a_base + (uint32_t)(ks * 3 + tw) * 4096u, 128 * 16, 128),
                                              umma_smem_desc(b_base + (uint32_t)(ks * 3 + tg) * 1024u, kFuNS * 16, 128), idesc,
                                              (ks | tw | tg) ? 1u : 0u);
                        umma_commit(a_empty);  // the table tile may be overwritten once these UMMAs have read it
                    }
                    umma_commit(&tm_full[buf]);
                    umma_commit(&b_empty[buf]);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ---- consumers: warp = (channel group cg, cell word j); lane = cell 32 j + lane ----
        const int cg = warp >> 2, j = warp & 3;
        const int c0 = cg * kFuCH;
        const int cells = s.Gh * s.Gw, cell = j * 32 + lane;
        const bool valid = cell < cells;
        const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;
        int off9[9];  // tap offsets inside a swizzled plane; out-of-image taps read box row 0 (image row -1: zero fill)
        {
            const int y0 = oy * s.stride - 1, x0 = ox * s.stride - 1;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int iy = y0 + kh, ix = x0 + kw;
                    const bool in = valid && (unsigned)iy < 32u && (unsigned)ix < 32u;
                    const int rr = iy + 1;
                    off9[kh * 3 + kw] = in ? rr * 32 + ((((ix >> 2) ^ (rr & 7)) << 2) | (ix & 3)) : 0;
                }
        }
        float acc[kFuCH][27], dth[kFuCH], thr_c[kFuCH];
        bool chan_ok[kFuCH];
#pragma unroll
        for (int k = 0; k < kFuCH; ++k) {
            chan_ok[k] = c0 + k < s.C;
            thr_c[k] = __ldg(thr + min(c0 + k, s.C - 1));
            dth[k] = 0.0f;
#pragma unroll
            for (int t = 0; t < 27; ++t) acc[k][t] = 0.0f;
        }
        const uint32_t lane_base = tmem + ((uint32_t)(j * 32) << 16);
        int st = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_mine; ++i) {
            const int r = i / kFuNS, col = i % kFuNS, buf = r & 1;
            if (col == 0) {
                mbar_wait(&tm_full[buf], (r >> 1) & 1);
                tcgen05_fence_after();
            }
            // my channels' values of this sample, straight from the accumulators (issued before the ring wait)
            float gv[kFuCH];
#pragma unroll
            for (int k = 0; k < kFuCH; ++k) gv[k] = fu_tmem_ld1(lane_base + (uint32_t)(buf * 256 + min(c0 + k, s.C - 1) * kFuNS + col));
            mbar_wait(&full[st], ph);
            const float *stg = reinterpret_cast<const float *>(stages + (size_t)st * kFuStageBytes);
            const uint32_t *sbits = reinterpret_cast<const uint32_t *>(stages + (size_t)st * kFuStageBytes + kFuBitsOff);
            tmem_ld_wait();
            float g[kFuCH], x[kFuCH];
#pragma unroll
            for (int k = 0; k < kFuCH; ++k) {
                const int c = min(c0 + k, s.C - 1);
                const bool on = chan_ok[k] && ((sbits[c * 4 + j] >> lane) & 1u);
                g[k] = on ? gv[k] : 0.0f;
                x[k] = on ? stg[kFuXOff / 4 + (c * 4 + j) * 32 + lane] : 0.0f;
            }
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9)
#pragma unroll
                for (int ic = 0; ic < 3; ++ic) {
                    const float pt = stg[ic * kFuSwzPlane + off9[t9]];
#pragma unroll
                    for (int k = 0; k < kFuCH; ++k) acc[k][ic * 9 + t9] = fmaf(g[k], pt, acc[k][ic * 9 + t9]);
                }
#pragma unroll
            for (int k = 0; k < kFuCH; ++k) {
                const float z = kFuSharp * (x[k] - thr_c[k]);
                const float sgm = __fdividef(1.0f, 1.0f + __expf(-z));
                dth[k] = fmaf(-g[k], kFuSharp * sgm * (1.0f - sgm), dth[k]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
            if (++st == fa.ST) { st = 0; ph ^= 1u; }
            if (col == kFuNS - 1 || i == n_mine - 1) {  // this warp is done with the round's accumulators
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tm_empty[buf]);
            }
        }
#pragma unroll
        for (int k = 0; k < kFuCH; ++k) {
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                const float v = warp_sum(acc[k][t]);
                if (lane == 0) red[(warp * kFuCH + k) * 28 + t] = v;
            }
            const float v = warp_sum(dth[k]);
            if (lane == 0) red[(warp * kFuCH + k) * 28 + 27] = v;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    // per-CTA partial [C][28] (27 taps + the threshold gradient): the four cell words of a channel in word order
    for (int i = threadIdx.x; i < s.C * 28; i += blockDim.x) {
        const int c = i / 28, t = i % 28;
        const int w0 = (c / kFuCH) * 4, k = c % kFuCH;
        float v = 0.0f;
        for (int jj = 0; jj < 4; ++jj) v += red[((w0 + jj) * kFuCH + k) * 28 + t];
        partial[(size_t)blockIdx.x * s.C * 28 + i] = v;
    }
    if (warp == 10) tmem_dealloc<512>(tmem);
}

// ---- host side ---------------------------------------------------------------------------------------------------
static bool fused_image_tmap(const float *images, int B, CUtensorMap *tm) {
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
    const cuuint64_t dims[3] = {32, 32, (cuuint64_t)B * 3};
    const cuuint64_t strides[2] = {128, 4096};
    const cuuint32_t box[3] = {32, (cuuint32_t)kFuSwzRows, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(images), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

FusedPlan plan_input_bwd_fused(const nnue_shape &s) {
    FusedPlan p{};
    if (!get_option(kOptInputFusedGbin) || !ft_umma_ok(s) || !plan_input_bwd(s).fused) return p;
    if (s.H != 32 || s.W != 32 || s.CW != 4 || s.C > 8 || s.NW != 4 * s.C || (s.L1 != 32 && s.L1 != 64)) return p;
    p.nq = s.B < kNumSMs ? s.B : kNumSMs;
    p.rounds = ceil_div(ceil_div(s.B, p.nq), kFuNS);
    p.n_ks = s.L1 / 16;
    p.a_bytes = (uint32_t)p.n_ks * 3u * 4096u;        // [L1 / 16][3][128 x 16] bf16
    p.b_bytes = (uint32_t)p.n_ks * 3u * 1024u;        // [L1 / 16][3][32 x 16] bf16
    p.a_off = 4096;
    p.b_off = p.a_off + p.a_bytes;
    p.stage_off = (uint32_t)align_up((size_t)p.b_off + 2 * p.b_bytes, 1024);
    int ST = (int)((kMaxSmemOptin - (size_t)p.stage_off) / kFuStageBytes);
    if (ST > 16) ST = 16;
    if (ST < 3) return p;
    p.ST = ST;
    p.smem = (size_t)p.stage_off + (size_t)ST * kFuStageBytes;
    p.ws_wtiles = align_up((size_t)s.C * p.a_bytes, 256);
    p.ws_gtiles = align_up((size_t)p.nq * p.rounds * p.b_bytes, 256);
    p.ws_bytes = p.ws_wtiles + p.ws_gtiles + align_up((size_t)p.nq * s.C * 28 * 4, 256);
    p.ok = true;
    return p;
}

size_t ws_input_bwd_fused(const nnue_shape &s) { return plan_input_bwd_fused(s).ws_bytes; }

// umma_format_rows_kernel<true, 128> of ft_umma.cu: the table as [PP / 128][L1 / 16][3][128 x 16] tiles
int launch_format_table_rows128(const nnue_shape &s, const float *w, unsigned char *out, cudaStream_t st);

int launch_input_bwd_fused(const nnue_shape &s, const FusedPlan &p, const float *images, const uint32_t *bits_s, const float *xpad,
                           const float *ft_w, const float *g_ft, const float *thr, void *workspace, float **partial_out,
                           cudaStream_t st) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (!fused_image_tmap(images, s.B, &tm)) return NNUE_ERR_UNSUPPORTED;
    unsigned char *wt = static_cast<unsigned char *>(workspace), *gt = wt + p.ws_wtiles;
    float *partial = reinterpret_cast<float *>(gt + p.ws_gtiles);
    int rc = launch_format_table_rows128(s, ft_w, wt, st);
    if (rc != NNUE_OK) return rc;
    const long long n = 1LL * p.nq * p.rounds * kFuNS * (s.L1 / 8);
    fused_format_g_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(s.B, s.L1, p.nq, p.rounds, g_ft, gt);
    NNUE_CHECK_LAUNCH("fused_format_g_kernel");
    FusedArgs fa{p.nq, p.rounds, p.ST, p.n_ks, p.a_bytes, p.b_bytes, p.a_off, p.b_off, p.stage_off};
    NNUE_CUDA_TRY(cudaFuncSetAttribute(conv_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    conv_bwd_fused_kernel<<<p.nq, kFuWarps * 32, p.smem, st>>>(s, xpad, bits_s, wt, gt, thr, partial, fa, tm);
    NNUE_CHECK_LAUNCH("conv_bwd_fused_kernel");
    *partial_out = partial;
    return NNUE_OK;
}

}  // namespace nnue
