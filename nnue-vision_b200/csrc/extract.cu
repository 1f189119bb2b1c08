// extract.cu -- grid-feature extraction (conv 3x3 + threshold -> bitmasks) and its backward.
//
// Work decomposition: one warp owns a "unit" = (sample b, cell word j): 32 consecutive cells
// of the conv raster, one cell per lane.  A lane keeps its 3x3x3 input patch in registers and
// produces every channel for that cell, so __ballot_sync over the warp yields one 32-bit word
// of the sample-major bitmask per channel with no shared-memory traffic.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kExtThreads = 256;

// conv weights are staged in shared memory padded to 28 floats per channel so that a channel's
// 27 taps are read with 7 broadcast LDS.128
__device__ __forceinline__ void stage_conv_weights(float *sw, float *sthr, const float *conv_w, const float *thr,
                                                   int c0, int cn) {
    for (int i = threadIdx.x; i < cn * 28; i += blockDim.x) {
        const int c = i / 28, t = i % 28;
        sw[i] = t < 27 ? conv_w[(c0 + c) * 27 + t] : 0.0f;
    }
    for (int i = threadIdx.x; i < cn; i += blockDim.x) sthr[i] = thr[c0 + i];
}

// patch[ic*9 + kh*3 + kw] (PyTorch OIHW tap order), zero outside the image (padding = 1).
// Row / column validity and the 9 in-plane offsets are computed once; the three input channels
// reuse them (32-bit offsets: one image plane is < 2^31 elements).
__device__ __forceinline__ void load_patch(float (&patch)[28], const float *img, int H, int W, int oy, int ox,
                                           int stride, bool valid) {
    const int plane = H * W;
    const int y0 = oy * stride - 1, x0 = ox * stride - 1;
    bool rok[3], cok[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        rok[k] = valid && (unsigned)(y0 + k) < (unsigned)H;
        cok[k] = (unsigned)(x0 + k) < (unsigned)W;
    }
    const float *p0 = img + y0 * W + x0;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const bool in = rok[kh] && cok[kw];
            const float *p = p0 + kh * W + kw;
#pragma unroll
            for (int ic = 0; ic < 3; ++ic) patch[ic * 9 + kh * 3 + kw] = in ? __ldg(p + ic * plane) : 0.0f;
        }
    patch[27] = 0.0f;
}

// patch carries a 28th element fixed at 0 so the 7 x float4 weight reads need no tail case
__device__ __forceinline__ float conv_tap_sum(const float (&patch)[28], const float *w28) {
    const float4 *w4 = reinterpret_cast<const float4 *>(w28);
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        const float4 w = w4[q];
        acc = fmaf(patch[q * 4 + 0], w.x, acc);
        acc = fmaf(patch[q * 4 + 1], w.y, acc);
        acc = fmaf(patch[q * 4 + 2], w.z, acc);
        acc = fmaf(patch[q * 4 + 3], w.w, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(kExtThreads)
extract_fwd_kernel(const nnue_shape s, const float *__restrict__ images, const float *__restrict__ conv_w,
                   const float *__restrict__ thr, uint32_t *__restrict__ bits_s, float *__restrict__ xpad,
                   float *__restrict__ conv_out) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                 // [C][28]
    float *sthr = smem + s.C * 28;    // [C]
    stage_conv_weights(sw, sthr, conv_w, thr, 0, s.C);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int cells = s.Gh * s.Gw;
    const long long units = 1LL * s.B * s.CW;
    for (long long u = 1LL * blockIdx.x * wpb + warp; u < units; u += 1LL * gridDim.x * wpb) {
        const int b = (int)(u / s.CW), j = (int)(u % s.CW);
        const int cell = j * 32 + lane;
        const bool valid = cell < cells;
        const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;
        float patch[28];
        load_patch(patch, images + (size_t)b * 3 * s.H * s.W, s.H, s.W, oy, ox, s.stride, valid);
        for (int c = 0; c < s.C; ++c) {
            const float x = conv_tap_sum(patch, sw + c * 28);
            const unsigned word = __ballot_sync(kFull, valid && x > sthr[c]);
            if (bits_s && lane == 0) bits_s[(size_t)b * s.NW + c * s.CW + j] = word;
            if (xpad) xpad[(size_t)b * s.PP + (size_t)(c * s.CW + j) * 32 + lane] = x;  // coalesced 128 B
            if (conv_out && valid) conv_out[((size_t)b * s.C + c) * cells + cell] = x;
        }
    }
}

// ---- forward, fixed-word form -----------------------------------------------------------------------
// Same result as extract_fwd_kernel with the channel count a compile-time constant: a warp keeps ONE cell
// word j for the whole kernel and strides over the samples, so the 9 in-plane tap offsets and their
// validity are computed once, the channel loop is unrolled and the per-unit integer work disappears
// (ncu on the generic kernel: 743 instructions per (sample, word) unit, 70 % issue; this one ~400).
// Conv weights [C][27] and thresholds in the constant bank (copied device-to-device on the launching stream
// before the launch): the unrolled channel loop reads them through the uniform datapath, which keeps the
// 56 broadcast LDS.128 per unit of the generic kernel off the L1/shared-memory pipe the image loads need.
constexpr int kExtConstChannels = 32;
__constant__ float c_ext_w[kExtConstChannels * 27];
__constant__ float c_ext_thr[kExtConstChannels];

// PLANE = H*W when known at compile time (32 x 32 images: 1024), else 0.  With a fixed plane size a lane keeps nine
// 64-bit tap POINTERS and advances them once per sample; the three planes of a tap are then LDG immediates.  (ncu on
// the offset form: 9.8 M of 40 M executed instructions were 64-bit address arithmetic, 5.5 per load, at 70 % issue.)
template <int CT, int PLANE>
__global__ void __launch_bounds__(kExtThreads, 2)
extract_fwd_fixed_kernel(const nnue_shape s, const float *__restrict__ images, uint32_t *__restrict__ bits_s,
                         float *__restrict__ xpad) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (kExtThreads / 32) + (threadIdx.x >> 5), nw = gridDim.x * (kExtThreads / 32);
    const int j = gw % s.CW;                       // my cell word (needs nw % CW == 0)
    const int cells = s.Gh * s.Gw, plane = PLANE ? PLANE : s.H * s.W;
    const int cell = j * 32 + lane;
    const bool valid = cell < cells;
    const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;
    int off9[9];
    unsigned okm = 0;
    {
        const int y0 = oy * s.stride - 1, x0 = ox * s.stride - 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iy = y0 + kh, ix = x0 + kw;
                const bool in = valid && (unsigned)iy < (unsigned)s.H && (unsigned)ix < (unsigned)s.W;
                off9[kh * 3 + kw] = in ? iy * s.W + ix : 0;
                okm |= (in ? 1u : 0u) << (kh * 3 + kw);
            }
    }
    const int step = nw / s.CW;
    int b = gw / s.CW;
    const float *tap[9];  // PLANE only: my nine taps of the sample being fetched (plane 0)
    if (PLANE) {
#pragma unroll
        for (int t9 = 0; t9 < 9; ++t9) tap[t9] = images + (size_t)min(b, s.B - 1) * 3 * PLANE + off9[t9];
    }
    // the next sample's taps are in flight while this sample's channels are computed
    auto fetch = [&](float (&p)[27], int bb) {
        if (PLANE) {  // (bb is the sample the tap pointers stand on)
#pragma unroll
            for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) p[ic * 9 + t9] = __ldg(tap[t9] + ic * PLANE);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) tap[t9] += (size_t)step * 3 * PLANE;
        } else {
            const float *img = images + (size_t)bb * 3 * plane;
#pragma unroll
            for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) p[ic * 9 + t9] = __ldg(img + ic * plane + off9[t9]);
        }
    };
    float nxt[27];
    if (b < s.B) fetch(nxt, b);
    for (; b < s.B; b += step) {
        float patch[27];
#pragma unroll
        for (int t = 0; t < 27; ++t) patch[t] = ((okm >> (t % 9)) & 1u) ? nxt[t] : 0.0f;
        if (b + step < s.B) fetch(nxt, b + step);
        uint32_t *brow = bits_s + (size_t)b * s.NW + j;
        float *xrow = xpad ? xpad + (size_t)b * s.PP + (size_t)j * 32 + lane : nullptr;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float x = 0.0f;  // same tap order as conv_tap_sum
#pragma unroll
            for (int t = 0; t < 27; ++t) x = fmaf(patch[t], c_ext_w[c * 27 + t], x);
            const unsigned word = __ballot_sync(kFull, valid && x > c_ext_thr[c]);
            if (lane == 0) brow[c * s.CW] = word;
            if (xrow) xrow[(size_t)c * s.CW * 32] = x;
        }
    }
}

// ---- forward, TMA-staged form (CIFAR-sized images) ---------------------------------------------------
// Same result as extract_fwd_kernel.  A warp owns kExtCH channels of one cell word for the whole kernel and
// keeps their 3x3x3 taps in registers (no weight traffic at all in the loop); the producer (lane 0 of
// warp 0) streams each sample's three image planes into a ring of shared-memory stages with bulk TMA
// copies; out-of-image taps (padding = 1) point at a zeroed pad word behind each plane.  NH CTAs
// ("roles") cover all (channel group, cell word) units of a sample and walk the same sample stream.
template <int HWT>
__global__ void __launch_bounds__(kExtWarps * 32, 1)
extract_fwd_tma_kernel(const nnue_shape s, const float *__restrict__ images, const float *__restrict__ conv_w,
                       const float *__restrict__ thr, uint32_t *__restrict__ bits_s, float *__restrict__ xpad,
                       const ExtPlan pl) {
    constexpr int CH = kExtCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + kInMaxStages;
    float *stages = reinterpret_cast<float *>(smem_raw + kInHeader);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = blockIdx.x % pl.NH, q = blockIdx.x / pl.NH;
    const int HW = HWT ? HWT : s.H * s.W, HWp = HW + 4;

    if (threadIdx.x == 0) {
        for (int i = 0; i < pl.ST; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kExtWarps);
        }
        mbar_fence_init();
    }
    for (int i = threadIdx.x; i < pl.ST * 3 * 4; i += blockDim.x) {
        const int st = i / 12, pln = (i % 12) / 4, k = i % 4;
        stages[(size_t)st * pl.stage_floats + pln * HWp + HW + k] = 0.0f;
    }
    __syncthreads();

    const int n_mine = s.B > q ? (s.B - q + pl.nq - 1) / pl.nq : 0;
    auto produce = [&](int ii, int st, uint32_t ph) {
        if (ii >= pl.ST) mbar_wait(&empty[st], ph);
        const int b = q + ii * pl.nq;
        float *stg = stages + (size_t)st * pl.stage_floats;
        mbar_arrive_expect_tx(&full[st], (uint32_t)(3 * HW) * 4u);
        const float *img = images + (size_t)b * 3 * HW;
#pragma unroll
        for (int pln = 0; pln < 3; ++pln) tma_bulk_g2s(stg + pln * HWp, img + pln * HW, (uint32_t)HW * 4u, &full[st]);
    };
    const int ahead = pl.ST - (pl.ST > 2 ? kInLag : 1);
    if (threadIdx.x == 0)
        for (int ii = 0; ii < ahead && ii < n_mine; ++ii) produce(ii, ii, 0);
    int p_st = ahead % pl.ST;
    uint32_t p_ph = 1u;

    const int cells = s.Gh * s.Gw;
    const int CG = ceil_div(s.C, CH);
    const int unit = role * kExtWarps + warp;  // (channel group, cell word)
    const bool active = unit < CG * s.CW;
    const int cg = active ? unit / s.CW : 0, j = active ? unit % s.CW : 0;
    const int c0 = cg * CH;
    const int cell = j * 32 + lane;
    const bool valid = active && cell < cells;
    const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;
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
    float wk[CH][27], thr_c[CH];
    bool chan_ok[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        chan_ok[k] = active && c0 + k < s.C;
        const int c = min(c0 + k, s.C - 1);
        thr_c[k] = __ldg(thr + c);
#pragma unroll
        for (int t = 0; t < 27; ++t) wk[k][t] = __ldg(conv_w + c * 27 + t);
    }

    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_mine; ++i) {
        if (threadIdx.x == 0 && i + ahead < n_mine) produce(i + ahead, p_st, p_ph);
        if (++p_st == pl.ST) { p_st = 0; p_ph ^= 1u; }
        __syncwarp();
        mbar_wait(&full[st], ph);
        if (active) {
            const int b = q + i * pl.nq;
            const float *stg = stages + (size_t)st * pl.stage_floats;
            // same accumulation order as conv_tap_sum: taps in OIHW order, one FMA chain per channel
            float x[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) x[k] = 0.0f;
#pragma unroll
            for (int ic = 0; ic < 3; ++ic)
#pragma unroll
                for (int t9 = 0; t9 < 9; ++t9) {
                    const float pt = stg[ic * HWp + off9[t9]];
#pragma unroll
                    for (int k = 0; k < CH; ++k) x[k] = fmaf(pt, wk[k][ic * 9 + t9], x[k]);
                }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const unsigned word = __ballot_sync(kFull, valid && x[k] > thr_c[k]);
                if (chan_ok[k]) {  // warp-uniform
                    const int widx = (c0 + k) * s.CW + j;
                    if (lane == 0) bits_s[(size_t)b * s.NW + widx] = word;
                    if (xpad) xpad[(size_t)b * s.PP + (size_t)widx * 32 + lane] = x[k];
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == pl.ST) { st = 0; ph ^= 1u; }
    }
}

// Position-major copy of the bitmask: bits_t[pp][bw] holds, for padded position pp, one bit per
// sample of the 32-sample group bw.  This transpose IS the sort by feature that the weight
// gradient's segment reduction consumes.  One CTA transposes 8 sample groups x 32 words.
constexpr int kTrGroups = 8;
__global__ void __launch_bounds__(1024)
bits_transpose_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, uint32_t *__restrict__ bits_t) {
    __shared__ uint32_t tile[kTrGroups][32][33];
    const int k = threadIdx.x, wi = threadIdx.y;  // blockDim = (32, 32)
    const int w0 = blockIdx.x * 32, g0 = blockIdx.y * kTrGroups;
    // load: thread (k, wi) -> sample row wi of each group, word column k (coalesced over k)
    for (int g = 0; g < kTrGroups; ++g) {
        const int b = (g0 + g) * 32 + wi, w = w0 + k;
        tile[g][wi][k] = (b < s.B && w < s.NW) ? bits_s[(size_t)b * s.NW + w] : 0u;
    }
    __syncthreads();
    const int w = w0 + wi;
    if (w >= s.NW) return;
    for (int g = 0; g < kTrGroups; ++g) {
        if (g0 + g >= s.BW) break;
        unsigned out = 0;
#pragma unroll
        for (int sidx = 0; sidx < 32; ++sidx) out |= ((tile[g][sidx][wi] >> k) & 1u) << sidx;
        bits_t[((size_t)w * 32 + k) * s.BW + g0 + g] = out;
    }
}

__global__ void bits_count_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, int32_t *__restrict__ nnz) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= s.B) return;
    int n = 0;
    for (int w = lane; w < s.NW; w += 32) n += __popc(bits_s[(size_t)b * s.NW + w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(kFull, n, o);
    if (lane == 0) nnz[b] = n;
}

// bits -> padded (indices, values) as NNUE._to_sparse_features lays them out (nnue.py:590-635)
__global__ void sparse_from_bits_kernel(const nnue_shape s, const uint32_t *__restrict__ bits_s, int K,
                                        int64_t *__restrict__ idx, float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= s.B) return;
    const int cells = s.Gh * s.Gw;
    int off = 0;
    for (int w = 0; w < s.NW; ++w) {
        const unsigned word = bits_s[(size_t)b * s.NW + w];
        if (!word) continue;
        const int c = w / s.CW, j = w % s.CW;
        if ((word >> lane) & 1u) {
            const int pos = off + __popc(word & ((1u << lane) - 1u));
            if (pos < K) {
                idx[(size_t)b * K + pos] = (int64_t)c * cells + j * 32 + lane;
                val[(size_t)b * K + pos] = 1.0f;
            }
        }
        off += __popc(word);
    }
    for (int p = off + lane; p < K; p += 32) {
        idx[(size_t)b * K + p] = -1;
        val[(size_t)b * K + p] = 0.0f;
    }
}

// ---- backward: conv weight gradient --------------------------------------------------------
// g_conv_w[c][t] = sum over (b, cell) of g_bin[b, c, cell] * patch[b, cell][t], g_bin = dval at active
// positions and 0 elsewhere (the straight-through estimator passes g_bin to the conv unchanged; the
// threshold gradient is produced by nnue_ft_bwd_dval).  A warp owns kExbCCH channels and keeps their
// 27 tap accumulators in registers while it streams over cell words (one cell per lane, the lane's
// 3x3x3 patch in registers); the `cpc` channel chunks of a CTA walk the SAME cell words side by side,
// so the image is pulled from HBM once and the sibling warps hit L1.
__global__ void __launch_bounds__(kExbThreads, 2)
extract_bwd_kernel(const nnue_shape s, const float *__restrict__ images, const uint32_t *__restrict__ bits_s,
                   const float *__restrict__ dval, float *__restrict__ partial, int cpc) {
    __shared__ float red[kExbWarps][kExbCCH * 27];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int streams = kExbWarps / cpc;
    const int chunk_local = warp % cpc, stream = warp / cpc;
    const int cbase = blockIdx.y * cpc * kExbCCH;            // first channel of this CTA
    const int c0 = cbase + chunk_local * kExbCCH;            // first channel of this warp
    const int cn = max(0, min(kExbCCH, s.C - c0));
    const int cta_cn = max(0, min(cpc * kExbCCH, s.C - cbase));

    const int cells = s.Gh * s.Gw;
    float acc[kExbCCH][27];
#pragma unroll
    for (int cc = 0; cc < kExbCCH; ++cc)
#pragma unroll
        for (int t = 0; t < 27; ++t) acc[cc][t] = 0.0f;
    const int units = s.B * s.CW;
    if (cn > 0) {
        for (int u = blockIdx.x * streams + stream; u < units; u += gridDim.x * streams) {
            const int b = u / s.CW, j = u % s.CW;
            unsigned words[kExbCCH];
            unsigned any = 0;
#pragma unroll
            for (int cc = 0; cc < kExbCCH; ++cc) {
                words[cc] = cc < cn ? __ldg(bits_s + (size_t)b * s.NW + (c0 + cc) * s.CW + j) : 0u;
                any |= words[cc];
            }
            if (!any) continue;  // warp-uniform
            const int cell = j * 32 + lane;
            const bool valid = cell < cells;
            const int oy = valid ? cell / s.Gw : 0, ox = valid ? cell % s.Gw : 0;
            float patch[28];
            load_patch(patch, images + (size_t)b * 3 * s.H * s.W, s.H, s.W, oy, ox, s.stride, valid);
#pragma unroll
            for (int cc = 0; cc < kExbCCH; ++cc) {
                if (!words[cc]) continue;  // warp-uniform
                const bool on = (words[cc] >> lane) & 1u;
                const float g = on ? __ldg(dval + (size_t)b * s.PP + (size_t)((c0 + cc) * s.CW + j) * 32 + lane) : 0.0f;
#pragma unroll
                for (int t = 0; t < 27; ++t) acc[cc][t] = fmaf(g, patch[t], acc[cc][t]);
            }
        }
    }
    // block reduction: warp shuffle, then across the streams of each chunk through shared memory (fixed order)
#pragma unroll
    for (int cc = 0; cc < kExbCCH; ++cc)
#pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float v = warp_sum(acc[cc][t]);
            if (lane == 0) red[warp][cc * 27 + t] = v;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < cta_cn * 27; i += blockDim.x) {
        const int cl = i / (kExbCCH * 27), jj = i % (kExbCCH * 27);
        float v = 0.0f;
        for (int st = 0; st < streams; ++st) v += red[st * cpc + cl][jj];
        partial[((size_t)blockIdx.x * s.C + cbase) * 27 + i] = v;
    }
}

// out[i] = sum over blocks (in block order) of partial[blk][i]
__global__ void fold_rows_kernel(int n, int nblk, const float *__restrict__ partial, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = 0.0f;
    for (int k = 0; k < nblk; ++k) v += partial[(size_t)k * n + i];
    out[i] = v;
}

static int launch_extract_fwd(const nnue_shape &s, const float *images, const float *conv_w, const float *thr,
                              uint32_t *bits_s, float *xpad, float *conv_out, cudaStream_t st) {
    const ExtPlan ep = plan_extract_tma(s);
    if (ep.ok && bits_s && !conv_out) {
        auto k = s.H * s.W == 1024 ? extract_fwd_tma_kernel<1024> : extract_fwd_tma_kernel<0>;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ep.smem));
        k<<<ep.grid, kExtWarps * 32, ep.smem, st>>>(s, images, conv_w, thr, bits_s, xpad, ep);
        NNUE_CHECK_LAUNCH("extract_fwd_tma_kernel");
        return NNUE_OK;
    }
    if (bits_s && !conv_out && get_option(kOptExtractFixed) && (s.C == 4 || s.C == 8 || s.C == 16 || s.C == 32)) {
        // warps per grid: a multiple of CW, about 32 per SM, no more than one per unit
        long long warps = 32LL * kNumSMs / s.CW * s.CW;
        if (warps > 1LL * s.B * s.CW) warps = 1LL * s.B * s.CW;
        const int wpb = kExtThreads / 32;
        const int grid = (int)((warps + wpb - 1) / wpb);
        if ((1LL * grid * wpb) % s.CW == 0) {
            NNUE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_ext_w, conv_w, (size_t)s.C * 27 * 4, 0, cudaMemcpyDeviceToDevice, st));
            NNUE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_ext_thr, thr, (size_t)s.C * 4, 0, cudaMemcpyDeviceToDevice, st));
            const bool p32 = s.H * s.W == 1024;  // 32 x 32 images: tap-pointer form
            switch (s.C) {
                case 4: (p32 ? extract_fwd_fixed_kernel<4, 1024> : extract_fwd_fixed_kernel<4, 0>)<<<grid, kExtThreads, 0, st>>>(s, images, bits_s, xpad); break;
                case 8: (p32 ? extract_fwd_fixed_kernel<8, 1024> : extract_fwd_fixed_kernel<8, 0>)<<<grid, kExtThreads, 0, st>>>(s, images, bits_s, xpad); break;
                case 16: (p32 ? extract_fwd_fixed_kernel<16, 1024> : extract_fwd_fixed_kernel<16, 0>)<<<grid, kExtThreads, 0, st>>>(s, images, bits_s, xpad); break;
                default: (p32 ? extract_fwd_fixed_kernel<32, 1024> : extract_fwd_fixed_kernel<32, 0>)<<<grid, kExtThreads, 0, st>>>(s, images, bits_s, xpad); break;
            }
            NNUE_CHECK_LAUNCH("extract_fwd_fixed_kernel");
            return NNUE_OK;
        }
    }
    const size_t smem = (size_t)s.C * 29 * sizeof(float);
    if (smem > 48 * 1024) return NNUE_ERR_UNSUPPORTED;  // C <= 423 channels
    const long long units = 1LL * s.B * s.CW;
    const int wpb = kExtThreads / 32;
    long long grid = (units + wpb - 1) / wpb;
    const long long cap = 32LL * kNumSMs;
    if (grid > cap) grid = cap;
    extract_fwd_kernel<<<(int)grid, kExtThreads, smem, st>>>(s, images, conv_w, thr, bits_s, xpad, conv_out);
    NNUE_CHECK_LAUNCH("extract_fwd_kernel");
    return NNUE_OK;
}

int extract_xpad(const nnue_shape &s, const float *images, const float *conv_w, const float *thr, float *xpad,
                 cudaStream_t st) {
    return launch_extract_fwd(s, images, conv_w, thr, nullptr, xpad, nullptr, st);
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_extract_fwd(const nnue_shape *s, const float *images_d, const float *conv_w_d, const float *thr_d,
                     uint32_t *bits_s_d, uint32_t *bits_t_d, float *xpad_d, float *conv_out_d, int32_t *nnz_d,
                     void *stream) {
    if (!s || !images_d || !conv_w_d || !thr_d || !bits_s_d) return NNUE_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = launch_extract_fwd(*s, images_d, conv_w_d, thr_d, bits_s_d, xpad_d, conv_out_d, st);
    if (rc != NNUE_OK) return rc;
    if (bits_t_d) {
        dim3 g(ceil_div(s->NW, 32), ceil_div(s->BW, kTrGroups)), blk(32, 32);
        bits_transpose_kernel<<<g, blk, 0, st>>>(*s, bits_s_d, bits_t_d);
        NNUE_CHECK_LAUNCH("bits_transpose_kernel");
    }
    if (nnz_d) {
        bits_count_kernel<<<ceil_div(s->B, 8), 256, 0, st>>>(*s, bits_s_d, nnz_d);
        NNUE_CHECK_LAUNCH("bits_count_kernel");
    }
    return NNUE_OK;
}

int nnue_sparse_from_bits(const nnue_shape *s, const uint32_t *bits_s_d, int K, int64_t *idx_d, float *val_d,
                          void *stream) {
    if (!s || !bits_s_d || !idx_d || !val_d || K < 1) return NNUE_ERR_INVALID_ARG;
    sparse_from_bits_kernel<<<ceil_div(s->B, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(*s, bits_s_d, K, idx_d,
                                                                                                val_d);
    NNUE_CHECK_LAUNCH("sparse_from_bits_kernel");
    return NNUE_OK;
}

int nnue_extract_bwd(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *dval_d,
                     float *g_conv_w_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !images_d || !bits_s_d || !dval_d || !g_conv_w_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_extract_bwd(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace_d);
    const int gx = exb_grid_x(*s);
    dim3 grid(gx, exb_grid_y(s->C));
    extract_bwd_kernel<<<grid, kExbThreads, 0, st>>>(*s, images_d, bits_s_d, dval_d, partial, exb_cpc(s->C));
    NNUE_CHECK_LAUNCH("extract_bwd_kernel");
    fold_rows_kernel<<<ceil_div(s->C * 27, 128), 128, 0, st>>>(s->C * 27, gx, partial, g_conv_w_d);
    NNUE_CHECK_LAUNCH("fold_rows_kernel");
    return NNUE_OK;
}

}  // extern "C"
