// input_bwd.cu -- everything upstream of the feature transformer's input:
//   dval[b,p]  = <W[min(p,F-1)], g_ft[b]>                      (autograd edge of nnue.py:602-603, 705)
//   g_bin      = dval at active positions, 0 elsewhere          (straight-through, nnue.py:28-33)
//   g_thr[c]   = -sum g_bin * k * sig * (1 - sig)               (nnue.py:36-52, k = 10)
//   g_conv_w   = conv2d_weight(images, g_bin)
// Small tables (L1 <= 64, CIFAR-sized images) run two dense kernels: ft_bwd_dval_dense_kernel (ft.cu)
// writes g_bin, and conv_bwd_kernel below turns it into g_conv_w and g_thr in one pass over the
// images, recomputing the pre-threshold activations from the image taps it needs anyway (they are
// never stored by the forward).  Other shapes run the index-driven kernel pair on scratch.
//
// conv_bwd_kernel: a warp owns CH channels of one cell word (32 conv cells, one per lane) for the
// whole kernel and keeps their CH x 27 weight-gradient accumulators in registers.  The producer
// (lane 0 of warp 0) streams each sample's three image planes and its g_bin row into a ring of
// shared-memory stages with bulk TMA copies (full / empty mbarriers); consumers read image taps as
// LDS.32 and share them between their CH channels.  Out-of-image taps (padding = 1) are redirected
// to a zeroed pad word behind each plane, so the inner loop has no predicates.
#include <cuda.h>
#include <string.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr float kSteSharpness = 10.0f;  // nnue.py:41

// ---- swizzled staging of 32 x 32 images (SWZ) -------------------------------------------------------------------
// With a dense 32-word row pitch every image row starts in bank 0, so the lanes of a cell word (three raster rows
// of 11 cells at config D) hit each bank three times on every tap load (ncu: 56 % of the kernel's shared-memory
// wavefronts were conflicts).  The SWZ variant fetches a sample with ONE 3-D tensor-map TMA copy in the 128-byte
// swizzle mode: box = 32 columns x 40 rows x 3 planes starting at row -1, so that 16-byte chunk c of box row r lands
// at chunk c ^ (r & 7) -- rows three apart no longer share banks -- and the out-of-bounds rows (image rows -1 and
// 32 .. 38) arrive zero-filled, which also provides the padding taps.  Per-lane tap offsets absorb the swizzle.
constexpr int kSwzRows = 40;                         // box rows per plane (a multiple of the 8-row swizzle atom)
constexpr int kSwzPlane = kSwzRows * 32;             // floats per staged plane
constexpr uint32_t kSwzImgBytes = 3 * kSwzPlane * 4; // 15360 bytes per sample (12288 fetched, the rest zero fill)

__device__ __forceinline__ void tma_tensor3d_g2s(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// HWT: H*W when known at compile time (plane offsets become LDS immediates), 0 = read it from the shape
// XS: the forward stored the pre-threshold activations (xpad, staged next to g_bin) -- otherwise they are
//     recomputed from the taps and the conv weights
// P2: the conv-gradient accumulators as pairs of taps updated by the packed fp32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2:
//     two IEEE FMAs per instruction, so the results are bit-identical to the scalar form); needs XS
template <int CH, int WARPS, int HWT, bool XS, bool SWZ, bool P2 = false>
__global__ void __launch_bounds__(WARPS * 32, 1)
conv_bwd_kernel(const nnue_shape s, const float *__restrict__ images, const float *__restrict__ dval,
                const float *__restrict__ xpad, const float *__restrict__ conv_w, const float *__restrict__ thr,
                float *__restrict__ partial, const InPlan pl, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kInMaxStages;
    float *s_cw = reinterpret_cast<float *>(smem_raw + kInHeader);  // [C][28]
    float *s_thr = s_cw + s.C * 28;                                // [C]
    float *red = s_thr + align_up((size_t)s.C, 4);                 // [WARPS][CH][28]
    float *stages = reinterpret_cast<float *>(smem_raw + pl.stage_off);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    const int HW = HWT ? HWT : s.H * s.W, HWp = SWZ ? kSwzPlane : HW + 4;  // HWp: floats between staged planes

    if (threadIdx.x == 0) {
        for (int i = 0; i < pl.ST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], WARPS);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < s.C * 28; i += blockDim.x) {
        const int c = i / 28, t = i % 28;
        s_cw[i] = t < 27 ? conv_w[c * 27 + t] : 0.0f;
    }
    for (int i = threadIdx.x; i < s.C; i += blockDim.x) s_thr[i] = thr[i];
    // the 4 pad words behind every image plane stay zero for the whole kernel (TMA never writes them)
    if (!SWZ) for (int i = threadIdx.x; i < pl.ST * 3 * 4; i += blockDim.x) {
        const int st = i / 12, pln = (i % 12) / 4, k = i % 4;
        stages[(size_t)st * pl.stage_floats + pln * HWp + HW + k] = 0.0f;
    }
    __syncthreads();

    // Producer duty (lane 0 of warp 0): sample `ii` of this CTA's stream goes into stage ii % ST once every
    // warp has released that stage.  It refills kInLag samples behind the ring, so warp 0 only ever waits
    // for work the other warps finished two iterations ago.
    const int n_mine = s.B > q ? (s.B - q + pl.nq - 1) / pl.nq : 0;
    auto produce = [&](int ii, int st, uint32_t ph) {  // st = ii % ST, ph = parity of the release being waited for
        if (ii >= pl.ST) mbar_wait(&empty[st], ph);
        const int b = q + ii * pl.nq;
        float *stg = stages + (size_t)st * pl.stage_floats;
        if (SWZ) {
            mbar_arrive_expect_tx(&full[st], kSwzImgBytes + (uint32_t)((XS ? 2 : 1) * s.PP) * 4u);
            tma_tensor3d_g2s(stg, &tmap, 0, -1, 3 * b, &full[st]);
        } else {
            mbar_arrive_expect_tx(&full[st], (uint32_t)(3 * HW + (XS ? 2 : 1) * s.PP) * 4u);
            const float *img = images + (size_t)b * 3 * HW;
#pragma unroll
            for (int pln = 0; pln < 3; ++pln) tma_bulk_g2s(stg + pln * HWp, img + pln * HW, (uint32_t)HW * 4u, &full[st]);
        }
        tma_bulk_g2s(stg + 3 * HWp, dval + (size_t)b * s.PP, (uint32_t)s.PP * 4u, &full[st]);
        if (XS) tma_bulk_g2s(stg + 3 * HWp + s.PP, xpad + (size_t)b * s.PP, (uint32_t)s.PP * 4u, &full[st]);
    };
    const int ahead = pl.ST - (pl.ST > 2 ? kInLag : 1);  // samples in flight
    if (threadIdx.x == 0)
        for (int ii = 0; ii < ahead && ii < n_mine; ++ii) produce(ii, ii, 0);
    int p_st = ahead % pl.ST;  // stage / parity the producer uses next (sample i + ahead)
    uint32_t p_ph = 1u;        // flips to 0 when the cursor first wraps: ((ii / ST) - 1) & 1

    const int cells = s.Gh * s.Gw;
    const int CP = ceil_div(s.C, CH);                 // channel groups
    const int unit = role * WARPS + warp;             // (channel group, cell word)
    const bool active = unit < CP * s.CW;
    const int cg = active ? unit / s.CW : 0, j = active ? unit % s.CW : 0;
    const int c0 = cg * CH;
    const int cell = j * 32 + lane;
    const bool valid = active && cell < cells;
    const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;

    // tap offsets inside a plane; out-of-image taps point at the plane's zero pad
    int off9[9];
    {
        const int y0 = oy * s.stride - 1, x0 = ox * s.stride - 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iy = y0 + kh, ix = x0 + kw;
                const bool in = valid && (unsigned)iy < (unsigned)s.H && (unsigned)ix < (unsigned)s.W;
                if (SWZ) {  // box row iy + 1; out-of-image taps read box row 0 (image row -1: zero fill)
                    const int r = iy + 1;
                    off9[kh * 3 + kw] = in ? r * 32 + ((((ix >> 2) ^ (r & 7)) << 2) | (ix & 3)) : 0;
                } else {
                    off9[kh * 3 + kw] = in ? iy * s.W + ix : HW;
                }
            }
    }
    float acc[CH][27], dth[CH], thr_c[CH];
    float2 acc2[CH][14];  // P2: taps 2 u, 2 u + 1 (entry 13: tap 26 and nothing)
    const float *cwk[CH];
    int goff[CH];  // my g_bin element inside the staged dval row
    bool chan_ok[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        chan_ok[k] = active && c0 + k < s.C;
        const int c = min(c0 + k, s.C - 1);
        cwk[k] = s_cw + (size_t)c * 28;
        thr_c[k] = s_thr[c];
        goff[k] = 3 * HWp + (c * s.CW + j) * 32 + lane;
        dth[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < 27; ++t) acc[k][t] = 0.0f;
#pragma unroll
        for (int u = 0; u < 14; ++u) acc2[k][u] = make_float2(0.0f, 0.0f);
    }

    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_mine; ++i) {
        if (threadIdx.x == 0 && i + ahead < n_mine) produce(i + ahead, p_st, p_ph);
        if (++p_st == pl.ST) { p_st = 0; p_ph ^= 1u; }
        __syncwarp();
        mbar_wait(&full[st], ph);
        if (active) {
            const float *stg = stages + (size_t)st * pl.stage_floats;
            float gk[CH], x[CH][3];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                gk[k] = chan_ok[k] ? stg[goff[k]] : 0.0f;
                x[k][0] = (XS && chan_ok[k]) ? stg[goff[k] + s.PP] : 0.0f;
                x[k][1] = x[k][2] = 0.0f;
            }
            if (P2 && XS) {
                // the 27 taps two at a time: 13 packed FMAs + one scalar per channel instead of 27 scalar ones
                float2 g2[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) g2[k] = make_float2(gk[k], gk[k]);
#pragma unroll
                for (int u = 0; u < 13; ++u) {
                    const int ta = 2 * u, tb = 2 * u + 1;
                    const float2 pp = make_float2(stg[(ta / 9) * HWp + off9[ta % 9]], stg[(tb / 9) * HWp + off9[tb % 9]]);
#pragma unroll
                    for (int k = 0; k < CH; ++k) acc2[k][u] = __ffma2_rn(g2[k], pp, acc2[k][u]);
                }
                const float p26 = stg[2 * HWp + off9[8]];
#pragma unroll
                for (int k = 0; k < CH; ++k) acc2[k][13].x = fmaf(gk[k], p26, acc2[k][13].x);
            } else
            // one pass over the 27 taps: accumulate the conv weight gradient (and, without stored
            // activations, recompute the conv: one chain per input plane)
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9)
#pragma unroll
                for (int ic = 0; ic < 3; ++ic) {
                    const int t = ic * 9 + t9;
                    const float pt = stg[ic * HWp + off9[t9]];
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        if (!XS) x[k][ic] = fmaf(pt, cwk[k][t], x[k][ic]);
                        acc[k][t] = fmaf(gk[k], pt, acc[k][t]);
                    }
                }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const float z = kSteSharpness * ((x[k][0] + x[k][1] + x[k][2]) - thr_c[k]);
                const float sgm = __fdividef(1.0f, 1.0f + __expf(-z));
                dth[k] = fmaf(-gk[k], kSteSharpness * sgm * (1.0f - sgm), dth[k]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == pl.ST) { st = 0; ph ^= 1u; }
    }
    if (P2 && XS) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
#pragma unroll
            for (int u = 0; u < 13; ++u) {
                acc[k][2 * u] = acc2[k][u].x;
                acc[k][2 * u + 1] = acc2[k][u].y;
            }
            acc[k][26] = acc2[k][13].x;
        }
    }
    // warp reduction of the per-lane accumulators
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
    // per-CTA partial [C][28] (27 taps + the threshold gradient); channels other roles own get zeros
    for (int i = threadIdx.x; i < s.C * 28; i += blockDim.x) {
        const int c = i / 28, t = i % 28;
        float v = 0.0f;
        for (int wv = 0; wv < WARPS; ++wv) {  // fixed order
            const int u = role * WARPS + wv;
            if (u < CP * s.CW && u / s.CW == c / CH) v += red[(wv * CH + c % CH) * 28 + t];
        }
        partial[(size_t)blockIdx.x * s.C * 28 + i] = v;
    }
}

// g_conv_w[c][t] and g_thr[c] = sum over CTAs of partial[cta][c][28].  The fold is pure latency (148 dependent-free loads
// per element): a CTA of 256 threads takes 32 elements and eight slices of the CTA list (slice q: CTAs q, q + 8, ...),
// so every thread has its ~19 loads in flight at once; slices are combined in slice order -- a fixed order per shape.
__global__ void __launch_bounds__(256)
input_bwd_fold_kernel(int C, int nblk, const float *__restrict__ partial, float *__restrict__ g_conv_w,
                      float *__restrict__ g_thr) {
    __shared__ float red[8][32];
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), q = threadIdx.x >> 5;
    float v = 0.0f;
    if (i < C * 28) {
        float t[20];
        int k = q;
        while (k < nblk) {
#pragma unroll
            for (int u = 0; u < 20; ++u) t[u] = k + 8 * u < nblk ? __ldg(partial + (size_t)(k + 8 * u) * C * 28 + i) : 0.0f;
#pragma unroll
            for (int u = 0; u < 20; ++u) v += t[u];
            k += 160;
        }
    }
    red[q][threadIdx.x & 31] = v;
    __syncthreads();
    if (q == 0 && i < C * 28) {
        for (int s2 = 1; s2 < 8; ++s2) v += red[s2][threadIdx.x & 31];
        const int c = i / 28, t = i % 28;
        if (t < 27) g_conv_w[c * 27 + t] = v;
        else if (g_thr) g_thr[c] = v;
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// images [B][3][32][32] fp32 as a 3-D tensor (x, y, plane), box 32 x 40 x 3, 128-byte swizzle, zero fill out of bounds
static bool make_image_tmap(const float *images, int B, CUtensorMap *tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {32, 32, (cuuint64_t)B * 3};
    const cuuint64_t strides[2] = {128, 4096};
    const cuuint32_t box[3] = {32, (cuuint32_t)kSwzRows, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(images), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CH, int WARPS>
static int launch_conv_bwd(const nnue_shape &s, const InPlan &pl, const float *images, const float *dval,
                           const float *xpad, const float *conv_w, const float *thr, float *partial, cudaStream_t st) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (s.H == 32 && s.W == 32 && get_option(kOptInputSwizzle) && make_image_tmap(images, s.B, &tm)) {
        // swizzled staging: stage = 3 planes of 40 rows (1024-byte aligned) | g_bin row | activation row
        InPlan p2 = pl;
        p2.stage_off = (int)align_up((size_t)pl.stage_off, 1024);
        p2.stage_floats = (int)(align_up((size_t)kSwzImgBytes + 2 * (size_t)s.PP * 4, 1024) / 4);
        int ST = (int)((kMaxSmemOptin - (size_t)p2.stage_off) / ((size_t)p2.stage_floats * 4));
        if (ST > kInMaxStages) ST = kInMaxStages;
        if (ST >= 3) {
            p2.ST = ST;
            p2.smem = (size_t)p2.stage_off + (size_t)ST * p2.stage_floats * 4;
            auto k = xpad ? (get_option(kOptConvBwdPacked) ? conv_bwd_kernel<CH, WARPS, 1024, true, true, true>
                                                           : conv_bwd_kernel<CH, WARPS, 1024, true, true, false>)
                          : conv_bwd_kernel<CH, WARPS, 1024, false, true, false>;
            NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p2.smem));
            k<<<p2.grid, WARPS * 32, p2.smem, st>>>(s, images, dval, xpad, conv_w, thr, partial, p2, tm);
            NNUE_CHECK_LAUNCH("conv_bwd_kernel");
            return NNUE_OK;
        }
    }
    auto k = xpad ? (s.H * s.W == 1024 ? conv_bwd_kernel<CH, WARPS, 1024, true, false> : conv_bwd_kernel<CH, WARPS, 0, true, false>)
                  : (s.H * s.W == 1024 ? conv_bwd_kernel<CH, WARPS, 1024, false, false> : conv_bwd_kernel<CH, WARPS, 0, false, false>);
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k<<<pl.grid, WARPS * 32, pl.smem, st>>>(s, images, dval, xpad, conv_w, thr, partial, pl, tm);
    NNUE_CHECK_LAUNCH("conv_bwd_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_input_bwd_is_dense(const nnue_shape *s) { return s && plan_input_bwd(*s).fused ? 1 : 0; }

int nnue_ft_bwd_gbin(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *g_ft_d,
                     float *gbin_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !bits_s_d || !ft_w_d || !g_ft_d || !gbin_d) return NNUE_ERR_INVALID_ARG;
    if (!plan_input_bwd(*s).fused) return NNUE_ERR_UNSUPPORTED;
    if (ft_umma_ok(*s) && workspace_d && workspace_bytes >= ws_ft_gbin_umma(*s))  // tcgen05 contraction
        return launch_ft_bwd_gbin_umma(*s, bits_s_d, ft_w_d, g_ft_d, workspace_d, gbin_d, static_cast<cudaStream_t>(stream));
    if (ft_umma_ok(*s) && !dense_shape_ok(*s)) return NNUE_ERR_WORKSPACE;  // no CUDA-core form for this L1
    if (plan_ft_mma(*s).ok && workspace_d && workspace_bytes >= mma_wfrag_bytes(*s))  // tensor-core contraction
        return launch_ft_bwd_gbin_mma(*s, bits_s_d, ft_w_d, g_ft_d, static_cast<uint4 *>(workspace_d), gbin_d,
                                      static_cast<cudaStream_t>(stream));
    return launch_ft_bwd_dval_dense(*s, bits_s_d, ft_w_d, g_ft_d, gbin_d, static_cast<cudaStream_t>(stream));
}

int nnue_conv_bwd(const nnue_shape *s, const float *images_d, const float *gbin_d, const float *xpad_d,
                  const float *conv_w_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d, void *workspace_d,
                  size_t workspace_bytes, void *stream) {
    if (!s || !images_d || !gbin_d || !conv_w_d || !thr_d || !g_conv_w_d || !g_thr_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    const InPlan pl = plan_input_bwd(*s);
    if (!pl.fused) return NNUE_ERR_UNSUPPORTED;
    if (workspace_bytes < (size_t)pl.grid * s->C * 28 * 4) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace_d);
    int rc;
    if (pl.CH == 2) rc = launch_conv_bwd<2, 16>(*s, pl, images_d, gbin_d, xpad_d, conv_w_d, thr_d, partial, st);
    else rc = launch_conv_bwd<4, 8>(*s, pl, images_d, gbin_d, xpad_d, conv_w_d, thr_d, partial, st);
    if (rc != NNUE_OK) return rc;
    input_bwd_fold_kernel<<<ceil_div(s->C * 28, 32), 256, 0, st>>>(s->C, pl.grid, partial, g_conv_w_d, g_thr_d);
    NNUE_CHECK_LAUNCH("input_bwd_fold_kernel");
    return NNUE_OK;
}

int nnue_input_bwd_fused_ok(const nnue_shape *s) { return s && plan_input_bwd_fused(*s).ok ? 1 : 0; }

int nnue_input_bwd_fused(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *xpad_d,
                         const float *ft_w_d, const float *g_ft_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d,
                         void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !images_d || !bits_s_d || !xpad_d || !ft_w_d || !g_ft_d || !thr_d || !g_conv_w_d || !g_thr_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    const FusedPlan fp = plan_input_bwd_fused(*s);
    if (!fp.ok) return NNUE_ERR_UNSUPPORTED;
    if (workspace_bytes < fp.ws_bytes) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = nullptr;
    const int rc = launch_input_bwd_fused(*s, fp, images_d, bits_s_d, xpad_d, ft_w_d, g_ft_d, thr_d, workspace_d, &partial, st);
    if (rc != NNUE_OK) return rc;
    input_bwd_fold_kernel<<<ceil_div(s->C * 28, 32), 256, 0, st>>>(s->C, fp.nq, partial, g_conv_w_d, g_thr_d);
    NNUE_CHECK_LAUNCH("input_bwd_fold_kernel");
    return NNUE_OK;
}

int nnue_input_bwd_wants_activations(const nnue_shape *s) {
    if (!s) return 0;
    return plan_input_bwd(*s).fused ? 0 : 1;  // (the dense pair takes them through nnue_conv_bwd)
}

int nnue_input_bwd(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *ft_w_d,
                   const float *g_ft_d, const float *conv_w_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d,
                   void *workspace_d, size_t workspace_bytes, void *stream) {
    return nnue_input_bwd_stored(s, images_d, bits_s_d, nullptr, ft_w_d, g_ft_d, conv_w_d, thr_d, g_conv_w_d, g_thr_d,
                                 workspace_d, workspace_bytes, stream);
}

int nnue_input_bwd_stored(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *xpad_d,
                          const float *ft_w_d, const float *g_ft_d, const float *conv_w_d, const float *thr_d,
                          float *g_conv_w_d, float *g_thr_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !images_d || !bits_s_d || !ft_w_d || !g_ft_d || !conv_w_d || !thr_d || !g_conv_w_d || !g_thr_d ||
        !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_input_bwd(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace_d);
    const size_t plane = align_up((size_t)s->B * s->PP * 4, 256);
    const InPlan pl = plan_input_bwd(*s);
    if (pl.fused) {
        float *gbin = reinterpret_cast<float *>(ws);
        const bool umma = ft_umma_ok(*s);  // its operand scratch sits behind the plane; the conv gradient reuses it afterwards
        const int rc = nnue_ft_bwd_gbin(s, bits_s_d, ft_w_d, g_ft_d, gbin, umma ? ws + plane : nullptr,
                                        umma ? workspace_bytes - plane : 0, stream);
        if (rc != NNUE_OK) return rc;
        return nnue_conv_bwd(s, images_d, gbin, xpad_d, conv_w_d, thr_d, g_conv_w_d, g_thr_d, ws + plane,
                             workspace_bytes - plane, stream);
    }
    // General shapes.  The pre-threshold activations come from the forward (xpad_d) or are recomputed into scratch; the
    // value gradient goes to a scratch plane; the conv / threshold gradients are taken from row-staged images
    // (input_bwd_rows.cu) or, where that kernel does not apply, by the direct-gather kernel of extract.cu.
    float *dval = reinterpret_cast<float *>(ws + plane);
    char *rest = ws + 2 * plane;
    const size_t rest_bytes = workspace_bytes - 2 * plane;
    int rc;
    const float *xpad = xpad_d;
    if (!xpad) {
        rc = extract_xpad(*s, images_d, conv_w_d, thr_d, reinterpret_cast<float *>(ws), st);
        if (rc != NNUE_OK) return rc;
        xpad = reinterpret_cast<const float *>(ws);
    }
    const RowsPlan rp = plan_conv_bwd_rows(*s);
    const size_t part_bytes = (size_t)rp.grid * s->C * 28 * 4;
    if (rp.ok && ft_umma_ok(*s)) {  // dense value gradient, then ONE pass for both the conv and the threshold gradient
        rc = launch_ft_bwd_gbin_umma(*s, bits_s_d, ft_w_d, g_ft_d, rest + align_up(part_bytes, 256), dval, st);
        if (rc != NNUE_OK) return rc;
        float *partial = reinterpret_cast<float *>(rest);
        rc = launch_conv_bwd_rows(*s, rp, images_d, bits_s_d, dval, xpad, thr_d, partial, st);
        if (rc != NNUE_OK) return rc;
        input_bwd_fold_kernel<<<ceil_div(s->C * 28, 32), 256, 0, st>>>(s->C, rp.grid, partial, g_conv_w_d, g_thr_d);
        NNUE_CHECK_LAUNCH("input_bwd_fold_kernel");
        return NNUE_OK;
    }
    rc = nnue_ft_bwd_dval(s, bits_s_d, ft_w_d, g_ft_d, xpad, thr_d, dval, g_thr_d, rest, rest_bytes, stream);
    if (rc != NNUE_OK) return rc;
    if (rp.ok) {  // (the index-driven value gradient has already produced the threshold gradient)
        float *partial = reinterpret_cast<float *>(rest);
        rc = launch_conv_bwd_rows(*s, rp, images_d, bits_s_d, dval, xpad, thr_d, partial, st);
        if (rc != NNUE_OK) return rc;
        input_bwd_fold_kernel<<<ceil_div(s->C * 28, 32), 256, 0, st>>>(s->C, rp.grid, partial, g_conv_w_d, nullptr);
        NNUE_CHECK_LAUNCH("input_bwd_fold_kernel");
        return NNUE_OK;
    }
    return nnue_extract_bwd(s, images_d, bits_s_d, dval, g_conv_w_d, rest, rest_bytes, stream);
}

}  // extern "C"
