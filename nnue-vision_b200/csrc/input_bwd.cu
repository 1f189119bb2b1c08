// input_bwd.cu -- everything upstream of the feature transformer's input in ONE pass over the images:
//   dval[b,p]  = <W[min(p,F-1)], g_ft[b]>                      (autograd edge of nnue.py:602-603, 705)
//   g_bin      = dval at active positions, 0 elsewhere          (straight-through, nnue.py:28-33)
//   g_thr[c]   = -sum g_bin * k * sig * (1 - sig)               (nnue.py:36-52, k = 10)
//   g_conv_w   = conv2d_weight(images, g_bin)
// Neither dval nor the pre-threshold activations ever touch HBM: the conv is recomputed from the
// image the weight gradient needs anyway.
//
// Ownership: a warp owns CH channels of one cell word (32 conv cells, one per lane) for the whole
// kernel, so the table rows of its positions stay in registers (CH x L1 floats per lane) next to the
// CH x 27 conv-gradient accumulators.  The warps of a CTA cover WARPS such units; NH CTAs ("roles")
// cover all of them and walk the same sample stream.  A producer lane streams each sample's image
// planes, g_ft row and bitmask row into a ring of shared-memory stages with bulk TMA copies
// (full / empty mbarriers; the producer is lane 0 of warp 0, so the CTA is exactly 8 warps = two per
// scheduler and every lane may use up to 255 registers); consumers read g_ft as broadcast LDS.128 and image taps as LDS.32.
// Out-of-image taps (padding = 1) are redirected to a zeroed pad word behind each plane, so the
// inner loop has no predicates.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr float kSteSharpness = 10.0f;  // nnue.py:41

template <int L1, int CH, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
input_bwd_kernel(const nnue_shape s, const float *__restrict__ images, const uint32_t *__restrict__ bits_s,
                 const float *__restrict__ ft_w, const float *__restrict__ g_ft, const float *__restrict__ conv_w,
                 const float *__restrict__ thr, float *__restrict__ partial, const InPlan pl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kInMaxStages;
    float *s_cw = reinterpret_cast<float *>(smem_raw + 128);      // [C][28]
    float *s_thr = s_cw + s.C * 28;                                // [C]
    float *red = s_thr + align_up((size_t)s.C, 4);                 // [WARPS][CH][28]
    float *stages = reinterpret_cast<float *>(smem_raw + pl.stage_off);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    const int HW = s.H * s.W, HWp = HW + 4;

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
    for (int i = threadIdx.x; i < pl.ST * 3 * 4; i += blockDim.x) {
        const int st = i / 12, pln = (i % 12) / 4, k = i % 4;
        stages[(size_t)st * pl.stage_floats + pln * HWp + HW + k] = 0.0f;
    }
    __syncthreads();

    // Producer duty (lane 0 of warp 0): sample `ii` of this CTA's stream goes into stage ii % ST once every
    // consumer warp has released that stage.  It runs kInLag samples behind the ring so that warp 0 only
    // ever waits for work the other warps finished two iterations ago.
    const int n_mine = s.B > q ? (s.B - q + pl.nq - 1) / pl.nq : 0;
    auto produce = [&](int ii) {
        const int st = ii % pl.ST;
        if (ii >= pl.ST) mbar_wait(&empty[st], ((ii / pl.ST) - 1) & 1);
        const int b = q + ii * pl.nq;
        float *stg = stages + (size_t)st * pl.stage_floats;
        mbar_arrive_expect_tx(&full[st], (uint32_t)(3 * HW + L1 + s.NW) * 4u);
        const float *img = images + (size_t)b * 3 * HW;
#pragma unroll
        for (int pln = 0; pln < 3; ++pln) tma_bulk_g2s(stg + pln * HWp, img + pln * HW, (uint32_t)HW * 4u, &full[st]);
        tma_bulk_g2s(stg + 3 * HWp, g_ft + (size_t)b * L1, L1 * 4u, &full[st]);
        tma_bulk_g2s(stg + 3 * HWp + L1, bits_s + (size_t)b * s.NW, (uint32_t)s.NW * 4u, &full[st]);
    };
    const int ahead = pl.ST - (pl.ST > 2 ? kInLag : 1);  // samples in flight
    if (threadIdx.x == 0)
        for (int ii = 0; ii < ahead && ii < n_mine; ++ii) produce(ii);
    {  // ---- consumers ----
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
                    off9[kh * 3 + kw] = in ? iy * s.W + ix : HW;
                }
        }
        // my table rows, stationary in registers
        float wr[CH][L1];
        bool chan_ok[CH];
        int widx[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            chan_ok[k] = active && c0 + k < s.C;
            widx[k] = chan_ok[k] ? (c0 + k) * s.CW + j : 0;
            const int p = (c0 + k) * cells + cell;
            const int row = min(p, s.F - 1);  // clamp of nnue.py:701
            const bool ok = valid && chan_ok[k];
#pragma unroll
            for (int v = 0; v < L1 / 4; ++v) {
                const float4 t = ok ? __ldg(reinterpret_cast<const float4 *>(ft_w + (size_t)row * L1) + v)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                wr[k][4 * v + 0] = t.x; wr[k][4 * v + 1] = t.y; wr[k][4 * v + 2] = t.z; wr[k][4 * v + 3] = t.w;
            }
        }
        float acc[CH][27], dth[CH], thr_c[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            dth[k] = 0.0f;
            thr_c[k] = chan_ok[k] ? s_thr[c0 + k] : 0.0f;
#pragma unroll
            for (int t = 0; t < 27; ++t) acc[k][t] = 0.0f;
        }
        const float *cwk[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) cwk[k] = s_cw + (size_t)min(c0 + k, s.C - 1) * 28;

        int i = 0;
        for (int b = q; b < s.B; b += pl.nq, ++i) {
            const int st = i % pl.ST;
            if (threadIdx.x == 0 && i + ahead < n_mine) produce(i + ahead);
            __syncwarp();
            mbar_wait(&full[st], (i / pl.ST) & 1);
            if (active) {
                const float *stg = stages + (size_t)st * pl.stage_floats;
                const float4 *sg = reinterpret_cast<const float4 *>(stg + 3 * HWp);
                const uint32_t *sb = reinterpret_cast<const uint32_t *>(stg + 3 * HWp + L1);
                unsigned word[CH], any = 0;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    word[k] = chan_ok[k] ? sb[widx[k]] : 0u;
                    any |= word[k];
                }
                if (any) {  // warp-uniform
                    // dval: CH dots against the broadcast g_ft row, two chains per dot
                    float d0[CH], d1[CH];
#pragma unroll
                    for (int k = 0; k < CH; ++k) d0[k] = d1[k] = 0.0f;
#pragma unroll
                    for (int v = 0; v < L1 / 4; ++v) {
                        const float4 g = sg[v];
#pragma unroll
                        for (int k = 0; k < CH; ++k) {
                            d0[k] = fmaf(wr[k][4 * v + 0], g.x, d0[k]);
                            d1[k] = fmaf(wr[k][4 * v + 1], g.y, d1[k]);
                            d0[k] = fmaf(wr[k][4 * v + 2], g.z, d0[k]);
                            d1[k] = fmaf(wr[k][4 * v + 3], g.w, d1[k]);
                        }
                    }
                    float gk[CH], x[CH];
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        gk[k] = ((word[k] >> lane) & 1u) ? d0[k] + d1[k] : 0.0f;
                        x[k] = 0.0f;
                    }
                    // one pass over the 27 taps: recompute the conv and accumulate its weight gradient
#pragma unroll
                    for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                        for (int t9 = 0; t9 < 9; ++t9) {
                            const float pt = stg[ic * HWp + off9[t9]];
                            const int t = ic * 9 + t9;
#pragma unroll
                            for (int k = 0; k < CH; ++k) {
                                x[k] = fmaf(pt, cwk[k][t], x[k]);
                                acc[k][t] = fmaf(gk[k], pt, acc[k][t]);
                            }
                        }
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        const float z = kSteSharpness * (x[k] - thr_c[k]);
                        const float sgm = __fdividef(1.0f, 1.0f + __expf(-z));
                        dth[k] = fmaf(-gk[k], kSteSharpness * sgm * (1.0f - sgm), dth[k]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
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
    }
    __syncthreads();
    // per-CTA partial [C][28] (27 taps + the threshold gradient); channels other roles own get zeros
    const int CP = ceil_div(s.C, CH);
    for (int i = threadIdx.x; i < s.C * 28; i += blockDim.x) {
        const int c = i / 28, t = i % 28;
        float v = 0.0f;
        for (int wv = 0; wv < WARPS; ++wv) {  // fixed order
            const int unit = role * WARPS + wv;
            if (unit < CP * s.CW && unit / s.CW == c / CH) v += red[(wv * CH + c % CH) * 28 + t];
        }
        partial[(size_t)blockIdx.x * s.C * 28 + i] = v;
    }
}

// g_conv_w[c][t] and g_thr[c] = sum over CTAs (in CTA order) of partial[cta][c][28]
__global__ void input_bwd_fold_kernel(int C, int nblk, const float *__restrict__ partial, float *__restrict__ g_conv_w,
                                      float *__restrict__ g_thr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * 28) return;
    float v = 0.0f;
    for (int k = 0; k < nblk; ++k) v += partial[(size_t)k * C * 28 + i];
    const int c = i / 28, t = i % 28;
    if (t < 27) g_conv_w[c * 27 + t] = v;
    else g_thr[c] = v;
}

template <int L1, int CH, int WARPS>
static int launch_input_bwd(const nnue_shape &s, const InPlan &pl, const float *images, const uint32_t *bits_s,
                            const float *ft_w, const float *g_ft, const float *conv_w, const float *thr, float *partial,
                            cudaStream_t st) {
    auto k = input_bwd_kernel<L1, CH, WARPS>;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k<<<pl.grid, WARPS * 32, pl.smem, st>>>(s, images, bits_s, ft_w, g_ft, conv_w, thr, partial, pl);
    NNUE_CHECK_LAUNCH("input_bwd_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_input_bwd(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *ft_w_d,
                   const float *g_ft_d, const float *conv_w_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d,
                   void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !images_d || !bits_s_d || !ft_w_d || !g_ft_d || !conv_w_d || !thr_d || !g_conv_w_d || !g_thr_d ||
        !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_input_bwd(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const InPlan pl = plan_input_bwd(*s);
    if (pl.fused) {
        float *partial = static_cast<float *>(workspace_d);
        int rc;
        if (s->L1 == 64) rc = launch_input_bwd<64, 2, kInWarps>(*s, pl, images_d, bits_s_d, ft_w_d, g_ft_d, conv_w_d, thr_d, partial, st);
        else rc = launch_input_bwd<32, 2, kInWarps>(*s, pl, images_d, bits_s_d, ft_w_d, g_ft_d, conv_w_d, thr_d, partial, st);
        if (rc != NNUE_OK) return rc;
        input_bwd_fold_kernel<<<ceil_div(s->C * 28, 128), 128, 0, st>>>(s->C, pl.grid, partial, g_conv_w_d, g_thr_d);
        NNUE_CHECK_LAUNCH("input_bwd_fold_kernel");
        return NNUE_OK;
    }
    // General shapes: recompute the pre-threshold activations into scratch, then the two-kernel path.
    char *ws = static_cast<char *>(workspace_d);
    const size_t plane = align_up((size_t)s->B * s->PP * 4, 256);
    float *xpad = reinterpret_cast<float *>(ws);
    float *dval = reinterpret_cast<float *>(ws + plane);
    void *rest = ws + 2 * plane;
    const size_t rest_bytes = workspace_bytes - 2 * plane;
    int rc = extract_xpad(*s, images_d, conv_w_d, thr_d, xpad, st);
    if (rc != NNUE_OK) return rc;
    rc = nnue_ft_bwd_dval(s, bits_s_d, ft_w_d, g_ft_d, xpad, thr_d, dval, g_thr_d, rest, rest_bytes, stream);
    if (rc != NNUE_OK) return rc;
    return nnue_extract_bwd(s, images_d, bits_s_d, dval, g_conv_w_d, rest, rest_bytes, stream);
}

}  // extern "C"
