// qinfer_fast.cu -- two more members of the integer-inference family (same integers as q_infer_kernel, qinfer.cu;
// reference: engine/src/nnue_engine.cpp:48-157 conv, 188-211 active set, 704-734 evaluate, simd_scalar.cpp:78-133):
//
//  * q_infer_cta_kernel      small batches.  One CTA of 16 warps per sample instead of one warp: a call with one image
//                            is bound by the dependent-load chain of its ~400 table rows, so the rows (and the conv
//                            words before them) are spread over the warps and the int16x2 partial accumulators are
//                            folded through shared memory (addition mod 2^16 is order-independent: same result).
//  * q_conv_bits32_kernel    large batches of 32 x 32 images at conv stride 4 (the 8 x 8 raster of every grid of 9..11
//                            squares, config D included): conv + threshold -> active-feature bitmask + density, the
//                            first of the three launches of the tensor-core form.  The image rows the conv touches
//                            (23 of 32) arrive by bulk TMA in half-sample units, a lane reads its 3 x 3 x 3 taps as nine
//                            conflict-free LDS.128 (a cell step is 48 bytes: the eight lanes of a quarter warp cover
//                            all 32 banks), releases the stage before any arithmetic, and the conv taps are
//                            constant-bank operands of the IMADs.  The engine's truncate-divide, clamp and threshold
//                            compare collapse into ONE integer compare against a bound worked out on the host
//                            (see conv_bound): no division in the kernel.
#include "plan.cuh"
#include "qinfer.cuh"

namespace nnue {

// ================================================ small batches ==================================================
constexpr int kQcWarps = 16;

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// out[o] = post(bias[o] + sum_k dp4a(in4[k], w[k][o])) for o < n_out, 0 for the padding outputs up to n_pad.
// Deep layers give every output to a warp (lanes split k, integer sums are exact in any order); shallow ones to a thread.
template <typename Post>
__device__ __forceinline__ void cta_dense(const int32_t *in4, int K, const int32_t *__restrict__ w, const int32_t *__restrict__ b,
                                          int n_out, int n_pad, int8_t *out, Post post) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (K >= 16) {
        for (int o = warp; o < n_pad; o += kQcWarps) {
            int a = 0;
            if (o < n_out)
                for (int k = lane; k < K; k += 32) a = __dp4a(in4[k], __ldg(w + (size_t)k * n_out + o), a);
            a = warp_sum_int(a);
            if (lane == 0) out[o] = o < n_out ? (int8_t)post(a + __ldg(b + o)) : (int8_t)0;
        }
    } else {
        for (int o = threadIdx.x; o < n_pad; o += kQcWarps * 32) {
            int v = 0;
            if (o < n_out) {
                int a = __ldg(b + o);
                for (int k = 0; k < K; ++k) a = __dp4a(in4[k], __ldg(w + (size_t)k * n_out + o), a);
                v = post(a);
            }
            out[o] = (int8_t)v;
        }
    }
}

// MAXW: 32-bit accumulator words per lane (L1p / 2 <= 32 * MAXW)
template <int MAXW>
__global__ void __launch_bounds__(kQcWarps * 32)
q_infer_cta_kernel(const QParams q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nwords = q.L1p / 2, NWq = q.OC * q.CWq;
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(smem_raw);                 // [OC][CWq] over the whole G x G buffer
    uint32_t *s_part = s_bits + NWq;                                           // [warps][nwords] int16x2 partial sums
    uint32_t *s_acc = s_part + kQcWarps * nwords;                              // [nwords]
    int8_t *s_pw = reinterpret_cast<int8_t *>(s_acc + nwords);                 // [4 K1]
    int8_t *s_h1 = s_pw + q.K1 * 4;                                            // [4 K2]
    int8_t *s_h2 = s_h1 + q.K2 * 4;                                            // [4 K3]
    int *s_count = reinterpret_cast<int *>(s_h2 + q.K3 * 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const int cells = q.oh * q.ow, CWr = (cells + 31) / 32;
    const float *img = q.images + (size_t)b * q.H * q.W * 3;
    if (threadIdx.x == 0) *s_count = 0;
    __syncthreads();

    // ---- conv + threshold: unit = (cell word j, group of channels); every word of the buffer is written ----
    {
        const int n_groups0 = min(q.OC, max(1, kQcWarps / q.CWq));
        const int per = (q.OC + n_groups0 - 1) / n_groups0, n_groups = (q.OC + per - 1) / per;
        const bool tail_on = 0.0f > q.threshold;  // cells past the raster stay 0 in the engine's buffer (nnue_engine.cpp:720)
        int n_active = 0;
        for (int u = warp; u < q.CWq * n_groups; u += kQcWarps) {
            const int j = u % q.CWq, oc0 = (u / q.CWq) * per, oc1 = min(q.OC, oc0 + per);
            const int cell = j * 32 + lane;
            const unsigned tail = tail_on ? __ballot_sync(kFull, cell >= cells && cell < q.G2) : 0u;
            if (j >= CWr) {  // warp-uniform: a word entirely past the raster needs no conv
                for (int oc = oc0; oc < oc1; ++oc) {
                    const unsigned wd = oc < 64 ? tail : 0u;
                    n_active += __popc(wd);
                    if (lane == 0) s_bits[oc * q.CWq + j] = wd;
                }
                continue;
            }
            const bool valid = cell < cells;
            const int oy = valid ? cell / q.ow : 0, ox = valid ? cell % q.ow : 0;
            int xq[27];  // [kh][kw][ic], truncated (int32)(pixel * scale), 0 in the padding
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int iy = oy * q.stride + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ix = ox * q.stride + kw - 1;
                    const bool in = valid && iy >= 0 && iy < q.H && ix >= 0 && ix < q.W;
#pragma unroll
                    for (int ic = 0; ic < 3; ++ic) {
                        const float px = in ? __ldg(img + ((size_t)iy * q.W + ix) * 3 + ic) : 0.0f;
                        xq[(kh * 3 + kw) * 3 + ic] = __float2int_rz(__fmul_rn(px, q.conv_scale));
                    }
                }
            }
            for (int oc = oc0; oc < oc1; ++oc) {
                int a = __ldg(q.conv_b + oc), a1 = 0, a2 = 0;
                const int32_t *w = q.conv_w + oc * 28;
#pragma unroll
                for (int t = 0; t < 27; t += 3) {
                    a += xq[t] * __ldg(w + t);
                    a1 += xq[t + 1] * __ldg(w + t + 1);
                    a2 += xq[t + 2] * __ldg(w + t + 2);
                }
                a += a1 + a2;
                const int v = clampi(a / q.conv_iscale, -127, 127);
                unsigned wd = __ballot_sync(kFull, valid && (float)v > q.threshold) | tail;
                if (oc >= 64) wd = 0u;  // the engine's feature loop stops at channel 64 (nnue_engine.cpp:196)
                n_active += __popc(wd);
                if (lane == 0) s_bits[oc * q.CWq + j] = wd;
            }
        }
        if (lane == 0 && n_active) atomicAdd(s_count, n_active);
    }
    __syncthreads();

    // ---- accumulate: the buffer's words over the warps (raster words first: that is where the bits are) ----
    uint32_t acc[MAXW];
#pragma unroll
    for (int i = 0; i < MAXW; ++i) acc[i] = 0u;
    {
        constexpr int U = MAXW <= 4 ? 8 : 4;  // rows in flight per warp
        const uint32_t *ftw32 = reinterpret_cast<const uint32_t *>(q.ft_w);
        const int CWt = q.CWq - CWr;
        for (int u = warp; u < q.OC * q.CWq; u += kQcWarps) {
            int oc, j;
            if (u < q.OC * CWr) { oc = u / CWr; j = u % CWr; }
            else { const int t = u - q.OC * CWr; oc = t / CWt; j = CWr + t % CWt; }
            unsigned word = s_bits[oc * q.CWq + j];
            const unsigned fbase = ((unsigned)(j * 32) * (unsigned)q.OC + (unsigned)oc) * (unsigned)nwords + (unsigned)lane;
            const unsigned fstep = (unsigned)q.OC * (unsigned)nwords;
            while (word) {
                int k[U];
#pragma unroll
                for (int x = 0; x < U; ++x) {
                    k[x] = word ? __ffs(word) - 1 : -1;
                    word &= word - 1;
                }
#pragma unroll
                for (int i = 0; i < MAXW; ++i) {
                    if (i * 32 + lane < nwords) {
                        uint32_t v[U];
#pragma unroll
                        for (int x = 0; x < U; ++x) v[x] = k[x] >= 0 ? __ldg(ftw32 + fbase + (unsigned)k[x] * fstep + i * 32) : 0u;
#pragma unroll
                        for (int x = 0; x < U; ++x) acc[i] = __vadd2(acc[i], v[x]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXW; ++i)
        if (i * 32 + lane < nwords) s_part[warp * nwords + i * 32 + lane] = acc[i];
    __syncthreads();
    {
        const uint32_t *ftb32 = reinterpret_cast<const uint32_t *>(q.ft_b);
        for (int t = threadIdx.x; t < nwords; t += kQcWarps * 32) {
            uint32_t v = __ldg(ftb32 + t);
#pragma unroll
            for (int w = 0; w < kQcWarps; ++w) v = __vadd2(v, s_part[w * nwords + t]);
            s_acc[t] = v;
        }
    }
    __syncthreads();

    // ---- clipped ReLU + pairwise (nnue_engine.cpp:726-729, 490-500) ----
    const int16_t *a16 = reinterpret_cast<const int16_t *>(s_acc);
    const int half = q.L1 / 2;
    for (int i = threadIdx.x; i < q.K1 * 4; i += kQcWarps * 32) {
        int v = 0;
        if (i < half) {
            const int x = clampi(a16[i], 0, q.qone), y = clampi(a16[i + half], 0, q.qone);
            v = clampi((x * y) / 128, 0, 127);
        } else if (i < 2 * half) {
            v = clampi(clampi(a16[i - half], 0, q.qone), 0, 127);
        }
        s_pw[i] = (int8_t)v;
    }
    __syncthreads();
    // L1: float divide then truncate (simd_scalar.cpp:131-133), clamp 0..127
    const float l1_scale = q.l1_scale;
    cta_dense(reinterpret_cast<const int32_t *>(s_pw), q.K1, q.w1, q.b1, q.L2, q.K2 * 4, s_h1,
              [l1_scale](int a) { return clampi(__float2int_rz(__fdiv_rn(__int2float_rn(a), l1_scale)), 0, 127); });
    __syncthreads();
    // L2: integer divide (truncating), clamp +-127, ReLU (nnue_engine.cpp:512-523)
    const int l2_iscale = q.l2_iscale;
    cta_dense(reinterpret_cast<const int32_t *>(s_h1), q.K2, q.w2, q.b2, q.L3, q.K3 * 4, s_h2,
              [l2_iscale](int a) { return max(0, clampi(a / l2_iscale, -127, 127)); });
    __syncthreads();
    // output: (float)acc / output_scale (nnue_engine.cpp:526-533)
    const int32_t *h24 = reinterpret_cast<const int32_t *>(s_h2);
    for (int c = threadIdx.x; c < q.NC; c += kQcWarps * 32) {
        int a = __ldg(q.bo + c);
        for (int k = 0; k < q.K3; ++k) a = __dp4a(h24[k], __ldg(q.wo + (size_t)k * q.NC + c), a);
        q.logits[(size_t)b * q.NC + c] = __fdiv_rn(__int2float_rn(a), q.out_scale);
    }
    if (q.density && threadIdx.x == 0)
        q.density[b] = __fdiv_rn(__int2float_rn(*s_count), __int2float_rn(q.F));  // nnue_inference.cpp:54
}

static size_t q_cta_smem(const QParams &q) {
    const size_t nwords = (size_t)q.L1p / 2;
    return (size_t)q.OC * q.CWq * 4 + (kQcWarps + 1) * nwords * 4 + (size_t)(q.K1 + q.K2 + q.K3) * 4 + 16;
}

// NNUE_ERR_UNSUPPORTED when the sample's scratch does not fit one CTA (the caller falls back to the warp-per-sample kernel)
int launch_q_infer_cta(const QParams &q, cudaStream_t st) {
    const size_t smem = q_cta_smem(q);
    const int nwords = q.L1p / 2;
    if (smem > 200 * 1024 || nwords > 1024) return NNUE_ERR_UNSUPPORTED;
#define NNUE_QCLAUNCH(MAXW)                                                                                          \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            NNUE_CUDA_TRY(cudaFuncSetAttribute(q_infer_cta_kernel<MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        q_infer_cta_kernel<MAXW><<<q.B, kQcWarps * 32, smem, st>>>(q);                                               \
    } while (0)
    if (nwords <= 32) NNUE_QCLAUNCH(1);
    else if (nwords <= 128) NNUE_QCLAUNCH(4);
    else if (nwords <= 512) NNUE_QCLAUNCH(16);
    else NNUE_QCLAUNCH(32);
#undef NNUE_QCLAUNCH
    NNUE_CHECK_LAUNCH("q_infer_cta_kernel");
    return NNUE_OK;
}

// ====================================== 32 x 32 images, conv stride 4: bitmask ======================================
constexpr int kCbWarps = 14;                       // consumer warps; warps kCbWarps, kCbWarps + 1 are the TMA producers
constexpr int kCbThreads = (kCbWarps + 2) * 32;
constexpr int kCbRow = 32 * 3 * 4;                 // bytes of one HWC image row (384)
constexpr int kCbStageBytes = 16 + 12 * kCbRow;    // 16 bytes of slack in front (the ix = -1 read of the first cell) + 12 rows
constexpr int kCbStages = 44;                      // 203 KB: stages are released right after the tap loads, so ~all of it is in flight

constexpr int kCbMaxOC = 32;
// conv taps [oc][28] ([kh][kw][ic] as the engine indexes the file's bytes, nnue_engine.cpp:69; entry 27 = the bias) in the
// constant bank: every IMAD of the conv takes its tap as a c[][] operand.  Refreshed from the model's device copy by a
// stream-ordered copy in front of each launch (the float extraction does the same with its weights, extract.cu).
__constant__ int c_qconv_taps[kCbMaxOC * 28];
struct QConvArgs {
    const float *images;
    uint32_t *bits_out;   // [B][OC][CWq]
    float *density;       // [B] or null
    int B, CWq, G2, F;
    float conv_scale;
    int a_min;            // active iff conv sum >= a_min (mode 0); mode 1: never, mode 2: always
    int mode;
    int tail_on;          // 0 > threshold: the buffer cells past the raster are active
};

// The engine computes v = clamp(a / s, -127, 127) (C++ truncating division, s > 0) and fires iff (float)v > thr
// (nnue_engine.cpp:93-103, 199).  With t = the smallest integer above thr:  t > 127 never fires, t <= -127 always does,
// t >= 1: a / s >= t <=> a >= t s;  t <= 0: trunc(a / s) >= t <=> a > (t - 1) s.  Returns the mode, *a_min the bound.
static int conv_bound(float thr, int s, int *a_min) {
    *a_min = 0;
    if (!(thr == thr)) return 1;  // NaN compares false
    if (thr >= 127.0f) return 1;
    if (thr < -127.0f) return 2;
    const long long t = (long long)floorf(thr) + 1;  // in [-127, 127]
    if (t <= -127) return 2;
    const long long bound = t >= 1 ? t * (long long)s : (t - 1) * (long long)s + 1;
    if (bound > 2147483647LL) return 1;
    if (bound < -2147483648LL) return 2;
    *a_min = (int)bound;
    return 0;
}

template <int OC>
__global__ void __launch_bounds__(kCbThreads, 1)
q_conv_bits32_kernel(const QConvArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);   // [stages] TMA -> consumer
    uint64_t *empty = full + kCbStages;                        // [stages] consumer -> TMA
    unsigned char *ring = smem_raw + 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kCbStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // this CTA's samples: blockIdx.x, + gridDim.x, ...; local sample i -> consumer warp i % kCbWarps, stages 2i and 2i + 1
    const int n_local = a.B > (int)blockIdx.x ? (a.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp >= kCbWarps) {
        // ---- producers: warp kCbWarps + h feeds half h of every sample (stage 2 i + h): the four 3-row groups its cells
        // touch (image rows 4 oy - 1 .. 4 oy + 1 -> stage rows 3 g .. 3 g + 2), one bulk copy each, issued by ONE lane
        // (a bulk copy is a warp-level instruction: several active lanes would only be serialised, and a single warp
        // feeding both halves was the bottleneck of the first version -- 55 % of the consumers' samples sat on `full`)
        const int h = warp - kCbWarps;
        if (lane == 0) {
            const unsigned char *src = reinterpret_cast<const unsigned char *>(a.images) + (size_t)blockIdx.x * (32 * kCbRow) +
                                       (h ? 15 * kCbRow : 0);
            const size_t src_step = (size_t)gridDim.x * (32 * kCbRow);
            const uint32_t tx = h ? 12 * kCbRow : 11 * kCbRow;
            for (int i = 0; i < n_local; ++i, src += src_step) {
                const int n = 2 * i + h, st = n % kCbStages;
                if (n >= kCbStages) mbar_wait(&empty[st], ((n / kCbStages) - 1) & 1);
                unsigned char *dst = ring + (size_t)st * kCbStageBytes + 16;
                mbar_arrive_expect_tx(&full[st], tx);
                if (h == 0) {  // rows 0..1 (row -1 is padding, never read as data), 3..5, 7..9, 11..13
                    tma_bulk_g2s(dst + kCbRow, src, 2 * kCbRow, &full[st]);
#pragma unroll
                    for (int g = 1; g < 4; ++g) tma_bulk_g2s(dst + 3 * g * kCbRow, src + (4 * g - 1) * kCbRow, 3 * kCbRow, &full[st]);
                } else {       // rows 15..17, 19..21, 23..25, 27..29
#pragma unroll
                    for (int g = 0; g < 4; ++g) tma_bulk_g2s(dst + 3 * g * kCbRow, src + 4 * g * kCbRow, 3 * kCbRow, &full[st]);
                }
            }
        }
        return;
    }

    // ---- consumers: lane = cell of the half (4 raster rows x 8 columns): oy = 4 half + lane / 8, ox = lane % 8 ----
    const int ox = lane & 7, gy = lane >> 3;
    const uint32_t my_off = 16 + (uint32_t)(3 * gy) * kCbRow + (uint32_t)ox * 48 - 16;  // 16 bytes before pixel 4 ox of stage row 3 gy
    for (int i = warp; i < n_local; i += kCbWarps) {
        const int b = (int)blockIdx.x + i * (int)gridDim.x;
        uint32_t my_word = 0u;  // lane l ends up with word l of the sample's [OC][CWq] bitmask (l < OC * CWq; looped beyond)
        int n_active = 0;
#pragma unroll 1  // (one copy of the conv in the instruction stream: the taps stay just-in-time uniform-register loads)
        for (int h = 0; h < 2; ++h) {
            const int n = 2 * i + h, st = n % kCbStages;
            mbar_wait(&full[st], (n / kCbStages) & 1);
            const unsigned char *p = ring + (size_t)st * kCbStageBytes + my_off;
            float4 v[9];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int c = 0; c < 3; ++c) v[kh * 3 + c] = *reinterpret_cast<const float4 *>(p + kh * kCbRow + c * 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);  // the stage is free again before any arithmetic happens
            // taps of row kh: pixel 4 ox - 1 = v[3 kh].yzw, pixel 4 ox = v[3 kh + 1].xyz, pixel 4 ox + 1 = v[3 kh + 1].w, v[3 kh + 2].xy
            int xq[27];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float4 f0 = v[kh * 3], f1 = v[kh * 3 + 1], f2 = v[kh * 3 + 2];
                const float px[9] = {f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y};
                const bool row_in = kh > 0 || h > 0 || gy > 0;  // image row -1 is padding
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const bool in = row_in && (t >= 3 || ox > 0);  // image column -1 is padding
                    xq[kh * 9 + t] = in ? __float2int_rz(__fmul_rn(px[t], a.conv_scale)) : 0;
                }
            }
#pragma unroll
            for (int oc = 0; oc < OC; ++oc) {
                int s0 = c_qconv_taps[oc * 28 + 27], s1 = 0, s2 = 0;
#pragma unroll
                for (int t = 0; t < 27; t += 3) {
                    s0 += xq[t] * c_qconv_taps[oc * 28 + t];
                    s1 += xq[t + 1] * c_qconv_taps[oc * 28 + t + 1];
                    s2 += xq[t + 2] * c_qconv_taps[oc * 28 + t + 2];
                }
                const int sum = s0 + (s1 + s2);
                const bool on = a.mode == 0 ? sum >= a.a_min : a.mode == 2;
                const unsigned wd = __ballot_sync(kFull, on);
                n_active += __popc(wd);
                if (OC * 4 <= 32 && oc * a.CWq + h == lane) my_word = wd;  // the record fits a warp: stored below
                if (OC * 4 > 32 && lane == 0) a.bits_out[((size_t)b * OC + oc) * a.CWq + h] = wd;
            }
        }
        // the sample's [OC][CWq] record: raster words from the lanes that kept them (one coalesced store when the record
        // fits a warp), words past the raster (cells 64 .. G2 - 1) all active iff 0 > threshold
        uint32_t *orow = a.bits_out + (size_t)b * OC * a.CWq;
        for (int w0 = 0; w0 < OC * a.CWq; w0 += 32) {
            const int wI = w0 + lane;
            if (wI < OC * a.CWq) {
                const int j = wI % a.CWq;
                if (j >= 2) {
                    const int rem = a.G2 - 32 * j;
                    orow[wI] = (a.tail_on && rem > 0) ? (rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1u)) : 0u;
                } else if (OC * 4 <= 32) {
                    orow[wI] = my_word;
                }
            }
        }
        if (a.tail_on) n_active += OC * max(0, a.G2 - 64);
        if (a.density && lane == 0) a.density[b] = __fdiv_rn(__int2float_rn(n_active), __int2float_rn(a.F));  // nnue_inference.cpp:54
    }
}

bool q_conv_bits32_ok(const QParams &q) {
    return q.H == 32 && q.W == 32 && q.stride == 4 && q.oh == 8 && q.ow == 8 && q.G2 >= 64 && q.CWq >= 2 &&
           (q.OC == 4 || q.OC == 8 || q.OC == 16 || q.OC == 32) && q.conv_iscale > 0 &&
           (reinterpret_cast<uintptr_t>(q.images) & 15) == 0;
}

template <int OC>
static int launch_cb(const QParams &q, const int32_t *taps_d, cudaStream_t st) {
    NNUE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_qconv_taps, taps_d, (size_t)OC * 28 * 4, 0, cudaMemcpyDeviceToDevice, st));
    QConvArgs a{};
    a.images = q.images; a.bits_out = q.bits_out; a.density = q.density;
    a.B = q.B; a.CWq = q.CWq; a.G2 = q.G2; a.F = q.F; a.conv_scale = q.conv_scale;
    a.mode = conv_bound(q.threshold, q.conv_iscale, &a.a_min);
    a.tail_on = 0.0f > q.threshold ? 1 : 0;
    constexpr size_t smem = 1024 + (size_t)kCbStages * kCbStageBytes;
    NNUE_CUDA_TRY(cudaFuncSetAttribute(q_conv_bits32_kernel<OC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = min(kNumSMs, ceil_div(q.B, kCbWarps));
    q_conv_bits32_kernel<OC><<<grid, kCbThreads, smem, st>>>(a);
    NNUE_CHECK_LAUNCH("q_conv_bits32_kernel");
    return NNUE_OK;
}

// taps_d [OC][28]: the model's device copy of the taps with the bias in entry 27 of each channel
int launch_q_conv_bits32(const QParams &q, const int32_t *taps_d, cudaStream_t st) {
    switch (q.OC) {
        case 4: return launch_cb<4>(q, taps_d, st);
        case 8: return launch_cb<8>(q, taps_d, st);
        case 16: return launch_cb<16>(q, taps_d, st);
        case 32: return launch_cb<32>(q, taps_d, st);
    }
    return NNUE_ERR_UNSUPPORTED;
}

}  // namespace nnue

extern "C" int nnue_q_conv_bound(float threshold, int conv_scale, int32_t *a_min) {
    if (!a_min || conv_scale < 1) return NNUE_ERR_INVALID_ARG;
    int bound = 0;
    const int mode = nnue::conv_bound(threshold, conv_scale, &bound);
    *a_min = bound;
    return mode;
}
