// qinfer.cu -- quantized integer inference, bit-exact against serialize.py + engine/ of the
// reference: .nnue v2 loader (host) and one fused kernel per batch:
//   conv (int32) -> int8 clamp -> threshold -> int16 wrap-around accumulate -> clipped ReLU ->
//   pairwise -> L1 (float divide, truncate) -> L2 (integer divide) -> output / 64.
// One warp owns a sample; the active set never leaves registers (ballot words), the accumulator
// is packed int16x2 (__vadd2 wraps exactly like the engine's int16 adds), the dense layers are
// __dp4a over weights repacked at load time (all activations are in [0,127] by construction).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "plan.cuh"
#include "qinfer.cuh"

struct nnue_qmodel {
    int F, L1, L2, L3, NC, OC, G, n_buckets;
    float nnue2score, quantized_one, threshold;
    float conv_scale, ft_scale;
    int L1p;   // L1 rounded up to 8 (table rows are 16-byte aligned)
    int K1, K2, K3;  // dp4a depth (groups of 4 inputs) of the three dense layers
    // device blobs
    int32_t *conv_w;   // [OC][28] taps in the engine's [kh][kw][ic] order, widened to int32
    int32_t *conv_b;   // [OC]
    int32_t *conv_taps = nullptr;  // [OC][28] taps with the bias in entry 27 (the fixed-shape conv kernel's constant-bank image)
    int16_t *ft_w;     // [F][L1p]
    int16_t *ft_b;     // [L1p]  bias truncated to int16 (simd_scalar.cpp:82-84)
    int32_t *ft_b32 = nullptr;  // [L1] the same bias as stored (the tensor-core accumulate wraps at the end)
    struct Stack {
        float l1_scale, l2_scale, out_scale;
        int32_t *w1, *b1;  // w1 [K1][L2] dp4a words: inputs 4k..4k+3 of output o
        int32_t *w2, *b2;  // w2 [K2][L3]
        int32_t *wo, *bo;  // wo [K3][NC]
        // raw blocks for the single-score LayerStack::forward behind evaluate_incremental (nnue_engine.cpp:382-478);
        // null when the file lacks what that path indexes (bias L2 of the combined layer)
        float l1_fact_scale;
        int8_t *w1_raw = nullptr;    // [(L2 + 1)][L1]
        int32_t *b1_raw = nullptr;   // [L2 + 1]
        int8_t *wf_row = nullptr;    // row L2 of the L1-factoriser weights [L1]
        int32_t bf_L2 = 0;           // its bias
        int8_t *w2_raw = nullptr;    // [L3][2 * L2]
        int8_t *wo_row = nullptr;    // output row 0 [L3]
        int32_t bo0 = 0;
    };
    std::vector<Stack> stacks;
    std::vector<void *> allocs;
    // scratch for nnue_q_infer_host (one instance per host thread, like NNUEEvaluator)
    mutable float *h_img = nullptr, *h_logits = nullptr, *h_density = nullptr;
    mutable size_t h_cap_img = 0, h_cap_b = 0;
    mutable cudaStream_t h_stream = nullptr;
    // tensor-core form for large batches (L1 a multiple of 64): the table as high/low-byte bf16 tiles in UMMA order,
    // K index = oc * CWq * 32 + cell over the whole G x G buffer.  The device entry points take their scratch from the
    // caller (nnue_q_workspace_bytes); only nnue_q_infer_host keeps its own (s_ws, with its other host-path buffers)
    unsigned char *tc_tiles = nullptr;
    int CWq = 0;                            // 32-bit words per channel: ceil(G*G / 32)
    mutable void *s_ws = nullptr;
    mutable size_t s_cap = 0;               // bytes
};

namespace nnue {

struct Reader {
    const unsigned char *p, *end;
    bool ok = true;
    bool take(void *dst, size_t n) {
        if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
        memcpy(dst, p, n);
        p += n;
        return true;
    }
    bool skip(size_t n) {
        if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
        p += n;
        return true;
    }
    size_t remaining() const { return ok ? (size_t)(end - p) : 0; }
    // may a payload of `count` elements of `elem` bytes still follow?  (checked BEFORE anything is allocated for it)
    bool fits(uint64_t count, uint64_t elem) const { return ok && count <= remaining() / (elem ? elem : 1); }
    uint32_t u32() { uint32_t v = 0; take(&v, 4); return v; }
    float f32() { float v = 0; take(&v, 4); return v; }
};

template <typename T>
static int upload(nnue_qmodel *m, const std::vector<T> &h, T **out) {
    void *d = nullptr;
    NNUE_CUDA_TRY(cudaMalloc(&d, h.size() * sizeof(T) + 16));
    m->allocs.push_back(d);
    NNUE_CUDA_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<T *>(d);
    return NNUE_OK;
}

// repack int8 [n_out][row_stride] (first k_in columns used) into dp4a words [ceil(k_in/4)][n_out]
static std::vector<int32_t> pack_dp4a(const std::vector<int8_t> &w, int n_out, int k_in, int row_stride) {
    const int K4 = (k_in + 3) / 4;
    std::vector<int32_t> out((size_t)K4 * n_out, 0);
    for (int o = 0; o < n_out; ++o)
        for (int k = 0; k < k_in; ++k) {
            const uint32_t byte = (uint8_t)w[(size_t)o * row_stride + k];
            out[(size_t)(k / 4) * n_out + o] |= (int32_t)(byte << (8 * (k % 4)));
        }
    return out;
}

// Format: serialize.py:30-63 (header), 103-136 (conv), 394-420 (FT), 423-491 (stack).  Checks
// mirror NNUEEvaluator::load_model / LayerStack::load_from_stream (nnue_engine.cpp:544-657, 283-380).
static int parse_and_upload(const unsigned char *bytes, size_t n, nnue_qmodel *m) {
    Reader r{bytes, bytes + n};
    char magic[4];
    if (!r.take(magic, 4) || memcmp(magic, "NNUE", 4) != 0) return NNUE_ERR_FORMAT;
    if (r.u32() != 2 || !r.ok) return NNUE_ERR_FORMAT;
    const uint32_t F = r.u32(), L1 = r.u32(), L2 = r.u32(), L3 = r.u32(), nb = r.u32();
    m->nnue2score = r.f32(); m->quantized_one = r.f32(); m->threshold = r.f32();
    (void)r.u32();  // conv layer type
    m->conv_scale = r.f32();
    const uint32_t OC = r.u32(), IC = r.u32(), KH = r.u32(), KW = r.u32();
    if (!r.ok || IC != 3 || KH != 3 || KW != 3 || OC == 0 || OC > (1u << 20) || !r.fits((uint64_t)OC * 27, 1)) return NNUE_ERR_FORMAT;
    if (L2 > (1u << 16) || L3 > (1u << 16)) return NNUE_ERR_FORMAT;  // (the engine's layer sizes are small integers)
    std::vector<int8_t> cw((size_t)OC * 27);
    if (!r.take(cw.data(), cw.size())) return NNUE_ERR_FORMAT;
    if (r.u32() != OC || !r.fits(OC, 4)) return NNUE_ERR_FORMAT;
    std::vector<int32_t> cb(OC);
    if (!r.take(cb.data(), (size_t)OC * 4)) return NNUE_ERR_FORMAT;
    if (F == 0 || F % OC) return NNUE_ERR_FORMAT;
    const int G = (int)sqrt((double)(F / OC));  // nnue_engine.cpp:599
    if ((uint32_t)(G * G) * OC != F) return NNUE_ERR_FORMAT;
    m->ft_scale = r.f32();
    if (r.u32() != F || r.u32() != L1 || !r.ok || L1 == 0) return NNUE_ERR_FORMAT;
    if ((uint64_t)F * L1 > (1ull << 31)) return NNUE_ERR_UNSUPPORTED;
    if (!r.fits((uint64_t)F * L1, 2)) return NNUE_ERR_FORMAT;  // truncated file: refuse before allocating the table
    std::vector<int16_t> fw((size_t)F * L1);
    if (!r.take(fw.data(), fw.size() * 2)) return NNUE_ERR_FORMAT;
    if (r.u32() != L1 || !r.fits(L1, 4)) return NNUE_ERR_FORMAT;
    std::vector<int32_t> fb(L1);
    if (!r.take(fb.data(), (size_t)L1 * 4)) return NNUE_ERR_FORMAT;
    if (nb < 1 || nb > 4096) return NNUE_ERR_FORMAT;

    m->F = (int)F; m->L1 = (int)L1; m->L2 = (int)L2; m->L3 = (int)L3; m->OC = (int)OC; m->G = G;
    m->n_buckets = (int)nb;
    m->L1p = (int)((L1 + 7) / 8 * 8);
    m->K1 = (int)((L1 + 3) / 4); m->K2 = (int)((L2 + 3) / 4); m->K3 = (int)((L3 + 3) / 4);

    // conv taps: the engine indexes the file's bytes as [oc][kh][kw][ic] (nnue_engine.cpp:69)
    std::vector<int32_t> cw32((size_t)OC * 28, 0);
    for (uint32_t oc = 0; oc < OC; ++oc)
        for (int t = 0; t < 27; ++t) cw32[(size_t)oc * 28 + t] = cw[(size_t)oc * 27 + t];
    std::vector<int16_t> fwp((size_t)F * m->L1p, 0), fbp((size_t)m->L1p, 0);
    for (uint32_t f = 0; f < F; ++f) memcpy(&fwp[(size_t)f * m->L1p], &fw[(size_t)f * L1], (size_t)L1 * 2);
    for (uint32_t i = 0; i < L1; ++i) fbp[i] = (int16_t)fb[i];
    std::vector<int32_t> taps28 = cw32;
    for (uint32_t oc = 0; oc < OC; ++oc) taps28[(size_t)oc * 28 + 27] = cb[oc];
    int rc;
    if ((rc = upload(m, taps28, &m->conv_taps))) return rc;
    if ((rc = upload(m, cw32, &m->conv_w)) || (rc = upload(m, cb, &m->conv_b)) || (rc = upload(m, fwp, &m->ft_w)) ||
        (rc = upload(m, fbp, &m->ft_b)) || (rc = upload(m, fb, &m->ft_b32)))
        return rc;

    if (L1 % 64 == 0 && (uint64_t)OC * ((uint64_t)(G * G + 31) / 32) * 32 * L1 * 4 <= (1ull << 31)) {
        // tiles [L1 / 64][k-step][128 rows x 16 k] bf16, canonical K-major: row n = term * 64 + column, term 0 = high
        // byte (signed), term 1 = low byte (unsigned): w = 256 * hi + lo, both exact in bf16
        const int CWq = (G * G + 31) / 32, NWq = (int)OC * CWq, n_ks = 2 * NWq;
        std::vector<uint16_t> tiles((size_t)(L1 / 64) * n_ks * 128 * 16, 0);
        auto bf16_of_int = [](int v) { float f = (float)v; uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); };
        for (int kk = 0; kk < 32 * NWq; ++kk) {
            const int oc = (kk >> 5) / CWq, cell = ((kk >> 5) % CWq) * 32 + (kk & 31);
            if (cell >= G * G) continue;
            const int16_t *row = &fw[((size_t)cell * OC + oc) * L1];
            const int ks = kk / 16, k = kk % 16;
            for (uint32_t col = 0; col < L1; ++col) {
                const int w = row[col], lo = w & 0xFF, hi = (w - lo) / 256;
                const size_t tile = ((size_t)(col / 64) * n_ks + ks) * (128 * 16);
                const int c = (int)(col % 64);
                for (int t = 0; t < 2; ++t) {
                    const int n = t * 64 + c;
                    tiles[tile + (size_t)(k / 8) * (128 * 8) + (size_t)(n / 8) * 64 + (size_t)(n % 8) * 8 + (k % 8)] =
                        bf16_of_int(t == 0 ? hi : lo);
                }
            }
        }
        uint16_t *d_tiles = nullptr;
        if ((rc = upload(m, tiles, &d_tiles))) return rc;
        m->tc_tiles = reinterpret_cast<unsigned char *>(d_tiles);
        m->CWq = CWq;
    }

    for (uint32_t b = 0; b < nb; ++b) {
        nnue_qmodel::Stack st{};
        st.l1_scale = r.f32(); st.l2_scale = r.f32(); st.out_scale = r.f32();
        st.l1_fact_scale = r.f32();
        uint32_t rows = r.u32(), cols = r.u32();
        if (!r.ok || rows != L2 + 1 || cols != L1 || L2 < 1 || L3 < 1 || !r.fits((uint64_t)rows * cols, 1)) return NNUE_ERR_FORMAT;
        std::vector<int8_t> w1((size_t)rows * cols);
        if (!r.take(w1.data(), w1.size())) return NNUE_ERR_FORMAT;
        uint32_t nbias = r.u32();
        if (!r.ok || nbias < L2 || !r.fits(nbias, 4)) return NNUE_ERR_FORMAT;
        std::vector<int32_t> b1(nbias);
        if (!r.take(b1.data(), (size_t)nbias * 4)) return NNUE_ERR_FORMAT;
        rows = r.u32(); cols = r.u32();  // L1 factoriser: unused by the multiclass head, row L2 feeds the single score
        if (!r.ok || cols != L1 || rows <= L2 || !r.fits((uint64_t)rows * cols, 1)) return NNUE_ERR_FORMAT;
        std::vector<int8_t> wf((size_t)rows * cols);
        if (!r.take(wf.data(), wf.size())) return NNUE_ERR_FORMAT;
        nbias = r.u32();
        if (!r.fits(nbias, 4)) return NNUE_ERR_FORMAT;
        std::vector<int32_t> bfv(nbias);
        if (!r.take(bfv.data(), (size_t)nbias * 4)) return NNUE_ERR_FORMAT;
        rows = r.u32(); cols = r.u32();
        if (!r.ok || cols != 2 * L2 || rows != L3 || !r.fits((uint64_t)rows * cols, 1)) return NNUE_ERR_FORMAT;
        std::vector<int8_t> w2((size_t)rows * cols);
        if (!r.take(w2.data(), w2.size())) return NNUE_ERR_FORMAT;
        nbias = r.u32();
        if (!r.ok || nbias < L3 || !r.fits(nbias, 4)) return NNUE_ERR_FORMAT;
        std::vector<int32_t> b2(nbias);
        if (!r.take(b2.data(), (size_t)nbias * 4)) return NNUE_ERR_FORMAT;
        rows = r.u32(); cols = r.u32();
        if (!r.ok || cols != L3 || rows < 1 || rows > (1u << 24) || !r.fits((uint64_t)rows * cols, 1)) return NNUE_ERR_FORMAT;
        if (b == 0) m->NC = (int)rows;
        else if ((int)rows != m->NC) return NNUE_ERR_FORMAT;
        std::vector<int8_t> wo((size_t)rows * cols);
        if (!r.take(wo.data(), wo.size())) return NNUE_ERR_FORMAT;
        nbias = r.u32();
        if (!r.ok || nbias < rows || !r.fits(nbias, 4)) return NNUE_ERR_FORMAT;
        std::vector<int32_t> bo(nbias);
        if (!r.take(bo.data(), (size_t)nbias * 4)) return NNUE_ERR_FORMAT;
        if ((rc = upload(m, pack_dp4a(w1, (int)L2, (int)L1, (int)L1), &st.w1)) || (rc = upload(m, b1, &st.b1)) ||
            (rc = upload(m, pack_dp4a(w2, (int)L3, (int)L2, (int)(2 * L2)), &st.w2)) || (rc = upload(m, b2, &st.b2)) ||
            (rc = upload(m, pack_dp4a(wo, m->NC, (int)L3, (int)L3), &st.wo)) || (rc = upload(m, bo, &st.bo)))
            return rc;
        if (b1.size() >= (size_t)L2 + 1 && bfv.size() > (size_t)L2) {
            std::vector<int32_t> b1x(b1.begin(), b1.begin() + L2 + 1);
            std::vector<int8_t> wfr(wf.begin() + (size_t)L2 * L1, wf.begin() + (size_t)(L2 + 1) * L1);
            std::vector<int8_t> wor(wo.begin(), wo.begin() + L3);
            if ((rc = upload(m, w1, &st.w1_raw)) || (rc = upload(m, b1x, &st.b1_raw)) || (rc = upload(m, wfr, &st.wf_row)) ||
                (rc = upload(m, w2, &st.w2_raw)) || (rc = upload(m, wor, &st.wo_row)))
                return rc;
            st.bf_L2 = bfv[L2];
            st.bo0 = bo[0];
        }
        m->stacks.push_back(st);
    }
    return NNUE_OK;
}

constexpr int kQWarps = 8;

// MAXW: 32-bit accumulator words per lane (L1p/2 <= 32*MAXW)
// MODE 0: the whole path in one kernel.  Large batches split it around the tensor-core accumulate
// (launch_q_accumulate_umma, ft_umma.cu): MODE 1 = conv + threshold -> active-feature bitmask (+ density),
// MODE 2 = clipped ReLU + pairwise + dense layers from the int16 accumulators.
template <int MAXW, int MODE>
__global__ void __launch_bounds__(kQWarps * 32)
q_infer_kernel(const QParams q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int32_t *s_cw = reinterpret_cast<int32_t *>(smem_raw);       // [OC][28]
    int32_t *s_cb = s_cw + q.OC * 28;                             // [OC]
    const int act_bytes = q.L1p * 2 + q.K1 * 4 + q.K2 * 4 + q.K3 * 4;
    unsigned char *s_act = reinterpret_cast<unsigned char *>(s_cb + q.OC) + (threadIdx.x >> 5) * act_bytes;
    for (int i = threadIdx.x; i < q.OC * 28; i += blockDim.x) s_cw[i] = q.conv_w[i];
    for (int i = threadIdx.x; i < q.OC; i += blockDim.x) s_cb[i] = q.conv_b[i];
    __syncthreads();

    uint32_t *s_acc = reinterpret_cast<uint32_t *>(s_act);                  // int16x2 [L1p/2]
    int8_t *s_pw = reinterpret_cast<int8_t *>(s_act + q.L1p * 2);           // [4*K1]
    int8_t *s_h1 = s_pw + q.K1 * 4;                                         // [4*K2]
    int8_t *s_h2 = s_h1 + q.K2 * 4;                                         // [4*K3]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwords = q.L1p / 2;
    const int cells = q.oh * q.ow;
    const uint32_t *ftw32 = reinterpret_cast<const uint32_t *>(q.ft_w);
    const uint32_t *ftb32 = reinterpret_cast<const uint32_t *>(q.ft_b);

    for (int b = blockIdx.x * kQWarps + warp; b < q.B; b += gridDim.x * kQWarps) {
        const float *img = q.images + (size_t)b * q.H * q.W * 3;
        uint32_t acc[MAXW];
#pragma unroll
        for (int i = 0; i < MAXW; ++i) acc[i] = (i * 32 + lane < nwords) ? __ldg(ftb32 + i * 32 + lane) : 0u;
        int n_active = 0;

        if (MODE == 2) {  // accumulators come from the tensor-core kernel
            const uint32_t *arow = reinterpret_cast<const uint32_t *>(q.acc_in) + (size_t)b * (q.L1 / 2);
#pragma unroll
            for (int i = 0; i < MAXW; ++i) acc[i] = (i * 32 + lane < q.L1 / 2) ? __ldg(arow + i * 32 + lane) : 0u;
        }
        // Q1 + Q2 + Q3: conv cell per lane, ballot per channel, accumulate rows of set bits
        // (MODE 1 walks the whole G x G buffer: cells past the conv raster stay 0 and are active iff 0 > threshold)
        for (int cell0 = 0; MODE != 2 && cell0 < (MODE == 1 ? q.G2 : cells); cell0 += 32) {
            const int cell = cell0 + lane;
            const bool valid = cell < cells;
            if (MODE == 1 && cell0 >= cells) {  // warp-uniform: a block entirely past the raster needs no conv
                const unsigned word = (0.0f > q.threshold) ? __ballot_sync(kFull, cell < q.G2) : 0u;
                for (int oc = 0; oc < q.OC; ++oc) {
                    const unsigned wd = oc < 64 ? word : 0u;
                    n_active += __popc(wd);
                    if (lane == 0) q.bits_out[((size_t)b * q.OC + oc) * q.CWq + (cell0 >> 5)] = wd;
                }
                continue;
            }
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
            for (int oc = 0; oc < q.OC; ++oc) {
                // four independent partial sums (integer addition is associative: same result, 4x shorter IMAD chains)
                int a = s_cb[oc], a1 = 0, a2 = 0, a3 = 0;
                const int4 *w4 = reinterpret_cast<const int4 *>(s_cw + oc * 28);  // 28 taps = 7 broadcast LDS.128
#pragma unroll
                for (int t4 = 0; t4 < 7; ++t4) {
                    const int4 w = w4[t4];
                    a += xq[4 * t4] * w.x;
                    a1 += xq[4 * t4 + 1] * w.y;
                    a2 += xq[4 * t4 + 2] * w.z;
                    if (t4 < 6) a3 += xq[4 * t4 + 3] * w.w;
                }
                a += (a1 + a2) + a3;
                const int v = clampi(a / q.conv_iscale, -127, 127);
                const bool on = valid ? (oc < 64 && (float)v > q.threshold)
                                      : (MODE == 1 && cell < q.G2 && oc < 64 && 0.0f > q.threshold);
                unsigned word = __ballot_sync(kFull, on);
                n_active += __popc(word);
                if (MODE == 1) {
                    if (lane == 0) q.bits_out[((size_t)b * q.OC + oc) * q.CWq + (cell0 >> 5)] = word;
                    continue;
                }
                // Rows of the set bits, four at a time: the loads of a group are issued together (the one-at-a-time loop
                // stalled on every row: ncu, long-scoreboard on the add) and the index math stays in 32 bits
                // (F * L1p / 2 < 2^31 is checked at load).  Addition mod 2^16 is order-independent.
                const unsigned fbase = ((unsigned)cell0 * (unsigned)q.OC + (unsigned)oc) * (unsigned)nwords + (unsigned)lane;
                const unsigned fstep = (unsigned)q.OC * (unsigned)nwords;
                while (word) {
                    int k[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        k[u] = word ? __ffs(word) - 1 : -1;
                        word &= word - 1;  // (0 & 0xffffffff stays 0)
                    }
#pragma unroll
                    for (int i = 0; i < MAXW; ++i) {
                        if (i * 32 + lane < nwords) {
                            uint32_t v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = k[u] >= 0 ? __ldg(ftw32 + fbase + (unsigned)k[u] * fstep + i * 32) : 0u;
#pragma unroll
                            for (int u = 0; u < 4; ++u) acc[i] = __vadd2(acc[i], v[u]);
                        }
                    }
                }
            }
        }
        if (MODE == 1) {
            if (q.density && lane == 0)
                q.density[b] = __fdiv_rn(__int2float_rn(n_active), __int2float_rn(q.F));  // nnue_inference.cpp:54
            continue;
        }
        // buffer entries past the conv raster stay 0 (nnue_engine.cpp:720): active iff 0 > threshold
        if (MODE == 0 && 0.0f > q.threshold) {
            for (int f = cells * q.OC; f < q.F; ++f) {
                if (f % q.OC >= 64) continue;
                ++n_active;
                const uint32_t *row = ftw32 + (size_t)f * nwords;
#pragma unroll
                for (int i = 0; i < MAXW; ++i)
                    if (i * 32 + lane < nwords) acc[i] = __vadd2(acc[i], __ldg(row + i * 32 + lane));
            }
        }
#pragma unroll
        for (int i = 0; i < MAXW; ++i)
            if (i * 32 + lane < nwords) s_acc[i * 32 + lane] = acc[i];
        __syncwarp();

        // Q4 clipped ReLU + Q5 pairwise (nnue_engine.cpp:726-729, 490-500)
        const int16_t *a16 = reinterpret_cast<const int16_t *>(s_acc);
        const int half = q.L1 / 2;
        for (int i = lane; i < q.K1 * 4; i += 32) {
            int v = 0;
            if (i < half) {
                const int x = clampi(a16[i], 0, q.qone), y = clampi(a16[i + half], 0, q.qone);
                v = clampi((x * y) / 128, 0, 127);
            } else if (i < 2 * half) {
                v = clampi(clampi(a16[i - half], 0, q.qone), 0, 127);
            }
            s_pw[i] = (int8_t)v;
        }
        __syncwarp();
        // L1: float divide then truncate (simd_scalar.cpp:131-133), clamp 0..127
        const int32_t *pw4 = reinterpret_cast<const int32_t *>(s_pw);
        for (int o = lane; o < q.K2 * 4; o += 32) {
            int v = 0;
            if (o < q.L2) {
                int a = __ldg(q.b1 + o);
                for (int k = 0; k < q.K1; ++k) a = __dp4a(pw4[k], __ldg(q.w1 + (size_t)k * q.L2 + o), a);
                v = clampi(__float2int_rz(__fdiv_rn(__int2float_rn(a), q.l1_scale)), 0, 127);
            }
            s_h1[o] = (int8_t)v;
        }
        __syncwarp();
        // L2: integer divide (truncating), clamp +-127, ReLU (nnue_engine.cpp:512-523)
        const int32_t *h14 = reinterpret_cast<const int32_t *>(s_h1);
        for (int o = lane; o < q.K3 * 4; o += 32) {
            int v = 0;
            if (o < q.L3) {
                int a = __ldg(q.b2 + o);
                for (int k = 0; k < q.K2; ++k) a = __dp4a(h14[k], __ldg(q.w2 + (size_t)k * q.L3 + o), a);
                v = max(0, clampi(a / q.l2_iscale, -127, 127));
            }
            s_h2[o] = (int8_t)v;
        }
        __syncwarp();
        // output: (float)acc / output_scale (nnue_engine.cpp:526-533)
        const int32_t *h24 = reinterpret_cast<const int32_t *>(s_h2);
        for (int c = lane; c < q.NC; c += 32) {
            int a = __ldg(q.bo + c);
            for (int k = 0; k < q.K3; ++k) a = __dp4a(h24[k], __ldg(q.wo + (size_t)k * q.NC + c), a);
            q.logits[(size_t)b * q.NC + c] = __fdiv_rn(__int2float_rn(a), q.out_scale);
        }
        if (MODE == 0 && q.density && lane == 0)
            q.density[b] = __fdiv_rn(__int2float_rn(n_active), __int2float_rn(q.F));  // nnue_inference.cpp:54
        __syncwarp();
    }
}

static int fill_params(const nnue_qmodel *m, int B, int H, int W, int bucket, QParams *q) {
    if (!m || B < 1 || H < 1 || W < 1) return NNUE_ERR_INVALID_ARG;
    if (bucket < 0) return NNUE_ERR_INVALID_ARG;
    if (bucket >= m->n_buckets) bucket = 0;  // nnue_engine.cpp:705-707
    // engine stride rule: ceil((H-1)/(G-1)), collapse to one cell when G == 1 (nnue_engine.cpp:710-718)
    int stride = m->G > 1 ? (H - 1 + m->G - 2) / (m->G - 1) : (H > 1 ? H : 1);
    if (stride < 1) stride = 1;
    q->B = B; q->H = H; q->W = W; q->stride = stride;
    q->oh = (H - 1) / stride + 1; q->ow = (W - 1) / stride + 1;
    if (1LL * q->oh * q->ow * m->OC > m->F) return NNUE_ERR_RASTER;
    q->F = m->F; q->L1 = m->L1; q->L2 = m->L2; q->L3 = m->L3; q->NC = m->NC; q->OC = m->OC;
    q->L1p = m->L1p; q->K1 = m->K1; q->K2 = m->K2; q->K3 = m->K3;
    q->threshold = m->threshold; q->conv_scale = m->conv_scale;
    q->conv_iscale = (int)m->conv_scale;
    q->qone = (int)(int16_t)m->quantized_one;
    const nnue_qmodel::Stack &st = m->stacks[(size_t)bucket];
    q->l2_iscale = (int)st.l2_scale; q->l1_scale = st.l1_scale; q->out_scale = st.out_scale;
    if (q->conv_iscale == 0 || q->l2_iscale == 0) return NNUE_ERR_FORMAT;
    q->conv_w = m->conv_w; q->conv_b = m->conv_b; q->ft_w = m->ft_w; q->ft_b = m->ft_b;
    q->w1 = st.w1; q->b1 = st.b1; q->w2 = st.w2; q->b2 = st.b2; q->wo = st.wo; q->bo = st.bo;
    q->G2 = m->F / m->OC; q->CWq = (q->G2 + 31) / 32;  // the whole G x G buffer as bitmask words per channel
    return NNUE_OK;
}

static int launch_q_infer(const QParams &q, int mode, cudaStream_t st) {
    const size_t act = (size_t)q.L1p * 2 + (size_t)(q.K1 + q.K2 + q.K3) * 4;
    const size_t smem = (size_t)q.OC * 29 * 4 + kQWarps * act;
    if (smem > 200 * 1024) return NNUE_ERR_UNSUPPORTED;
    int grid = ceil_div(q.B, kQWarps);
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    const int nwords = q.L1p / 2;
#define NNUE_QLAUNCH(MAXW)                                                                                   \
    do {                                                                                                     \
        auto k = mode == 0 ? q_infer_kernel<MAXW, 0> : mode == 1 ? q_infer_kernel<MAXW, 1> : q_infer_kernel<MAXW, 2>; \
        if (smem > 48 * 1024)                                                                                \
            NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k<<<grid, kQWarps * 32, smem, st>>>(q);                                                              \
    } while (0)
    if (nwords <= 32) NNUE_QLAUNCH(1);
    else if (nwords <= 128) NNUE_QLAUNCH(4);
    else if (nwords <= 512) NNUE_QLAUNCH(16);
    else if (nwords <= 1024) NNUE_QLAUNCH(32);
    else return NNUE_ERR_UNSUPPORTED;
#undef NNUE_QLAUNCH
    NNUE_CHECK_LAUNCH("q_infer_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

void nnue_q_free(nnue_qmodel *m) {
    if (!m) return;
    for (void *p : m->allocs) cudaFree(p);
    if (m->h_img) cudaFree(m->h_img);
    if (m->h_logits) cudaFree(m->h_logits);
    if (m->h_density) cudaFree(m->h_density);
    if (m->h_stream) cudaStreamDestroy(m->h_stream);
    if (m->s_ws) cudaFree(m->s_ws);
    delete m;
}

int nnue_q_load_memory(const void *bytes_h, size_t n_bytes, nnue_qmodel **out) {
    if (!bytes_h || !out) return NNUE_ERR_INVALID_ARG;
    *out = nullptr;
    // nothing may unwind through the C boundary: a malformed file answers NNUE_ERR_FORMAT the way
    // NNUEEvaluator::load_model answers false (nnue_engine.cpp:544-657)
    nnue_qmodel *m = nullptr;
    int rc;
    try {
        m = new nnue_qmodel();
        rc = parse_and_upload(static_cast<const unsigned char *>(bytes_h), n_bytes, m);
    } catch (...) {
        rc = NNUE_ERR_FORMAT;
    }
    if (rc != NNUE_OK) { nnue_q_free(m); return rc; }
    *out = m;
    return NNUE_OK;
}

int nnue_q_load(const char *path, nnue_qmodel **out) {
    if (!path || !out) return NNUE_ERR_INVALID_ARG;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return NNUE_ERR_IO;
    std::vector<unsigned char> buf;
    bool err = false;
    try {
        std::vector<unsigned char> chunk(1 << 20);
        size_t got;
        while ((got = fread(chunk.data(), 1, chunk.size(), f)) > 0) buf.insert(buf.end(), chunk.begin(), chunk.begin() + got);
        err = ferror(f) != 0;
    } catch (...) {
        err = true;
    }
    fclose(f);
    if (err) return NNUE_ERR_IO;
    return nnue_q_load_memory(buf.data(), buf.size(), out);
}

int nnue_q_dims(const nnue_qmodel *m, int32_t *dims, float *visual_threshold) {
    if (!m || !dims) return NNUE_ERR_INVALID_ARG;
    dims[0] = m->F; dims[1] = m->L1; dims[2] = m->L2; dims[3] = m->L3; dims[4] = m->NC; dims[5] = m->OC;
    dims[6] = m->G; dims[7] = m->n_buckets;
    if (visual_threshold) *visual_threshold = m->threshold;
    return NNUE_OK;
}

// scratch of the split (large-batch) form: bitmask [B][OC * CWq] u32 | accumulators [B][L1] i16
size_t nnue_q_workspace_bytes(const nnue_qmodel *m, int B) {
    if (!m || B < 1 || !m->tc_tiles) return 0;
    return align_up((size_t)B * m->OC * m->CWq * 4, 256) + align_up((size_t)B * m->L1 * 2, 256);
}

int nnue_q_infer_ws(const nnue_qmodel *m, const float *images_d, int B, int H, int W, int bucket, float *logits_d,
                    float *density_d, void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!images_d || !logits_d) return NNUE_ERR_INVALID_ARG;
    QParams q{};
    const int rc = fill_params(m, B, H, W, bucket, &q);
    if (rc != NNUE_OK) return rc;
    q.images = images_d; q.logits = logits_d; q.density = density_d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int min_b = get_option(kOptQTcMinBatch);
    if (m->tc_tiles && min_b > 0 && B >= min_b && workspace_d && workspace_bytes >= nnue_q_workspace_bytes(m, B)) {
        // large batches: bitmask -> tcgen05 accumulate -> layer stack (three launches, same integers) on the caller's scratch
        uint32_t *s_bits = static_cast<uint32_t *>(workspace_d);
        int16_t *s_acc = reinterpret_cast<int16_t *>(static_cast<char *>(workspace_d) + align_up((size_t)B * m->OC * m->CWq * 4, 256));
        q.bits_out = s_bits; q.acc_in = s_acc;
        int rc2 = (get_option(kOptQConvFixed) && q_conv_bits32_ok(q)) ? launch_q_conv_bits32(q, m->conv_taps, st)
                                                                      : launch_q_infer(q, 1, st);
        if (rc2 != NNUE_OK) return rc2;
        // the accumulate walks only the words that can hold a bit: a non-negative threshold leaves the cells past the conv
        // raster inactive (nnue_engine.cpp:720); small stacks on a 64-wide accumulator finish in the same kernel
        QAccArgs qa{};
        qa.cw_all = m->CWq;
        qa.cw_used = (0.0f > q.threshold) ? m->CWq : min(m->CWq, (q.oh * q.ow + 31) / 32);
        qa.n_words_used = m->OC * qa.cw_used;
        // (option q_stack_fused = the smallest batch that takes the one-kernel form, 0 = never: with few 128-sample tiles the
        // serial epilogue of a CTA costs more than the extra launch -- measured 24.0 vs 22.3 us at 2048, 64.6 vs 70.0 at 16384)
        const int fuse_min = get_option(kOptQStackFused);
        const bool fused = fuse_min > 0 && B >= fuse_min && q_stack_fused_ok(m->L1, m->L2, m->L3, m->NC);
        if (fused) {
            qa.w1 = q.w1; qa.b1 = q.b1; qa.w2 = q.w2; qa.b2 = q.b2; qa.wo = q.wo; qa.bo = q.bo;
            qa.L2 = q.L2; qa.L3 = q.L3; qa.NC = q.NC; qa.K2 = q.K2; qa.K3 = q.K3; qa.qone = q.qone; qa.l2_iscale = q.l2_iscale;
            qa.l1_scale = q.l1_scale; qa.out_scale = q.out_scale; qa.logits = logits_d;
        }
        rc2 = launch_q_accumulate_umma(B, m->OC * m->CWq, m->L1, s_bits, m->tc_tiles,
                                       reinterpret_cast<const int32_t *>(m->ft_b32), s_acc, qa, st);
        if (rc2 != NNUE_OK || fused) return rc2;
        return launch_q_infer(q, 2, st);
    }
    if (B <= get_option(kOptQCtaMaxBatch)) {  // small batches: a CTA per sample (falls through when its scratch does not fit)
        const int rc2 = launch_q_infer_cta(q, st);
        if (rc2 != NNUE_ERR_UNSUPPORTED) return rc2;
    }
    return launch_q_infer(q, 0, st);
}

// without scratch: the one-kernel form at every batch size (never allocates, never touches the model)
int nnue_q_infer(const nnue_qmodel *m, const float *images_d, int B, int H, int W, int bucket, float *logits_d,
                 float *density_d, void *stream) {
    return nnue_q_infer_ws(m, images_d, B, H, W, bucket, logits_d, density_d, nullptr, 0, stream);
}

}  // extern "C"

namespace nnue {

// ---- incremental accumulators (NNUEEvaluator::refresh_accumulator / update_features / evaluate_incremental,
//      nnue_engine.cpp:739-821), batched over S independent streams -----------------------------------------------
// acc [S][L1] int16.  A warp owns a stream: refresh starts from (int16)bias, otherwise from the stored accumulator;
// rows of `removed` are subtracted and rows of `added` are added with int16 wrap-around (simd_scalar.cpp:97-113);
// indices outside [0, F) are ignored (nnue_engine.cpp:215, 234).
__global__ void __launch_bounds__(256)
q_acc_apply_kernel(int S, int F, int L1, int L1p, const int16_t *__restrict__ ft_w, const int16_t *__restrict__ ft_b, int refresh,
                   const int32_t *__restrict__ add_off, const int32_t *__restrict__ add_idx,
                   const int32_t *__restrict__ rem_off, const int32_t *__restrict__ rem_idx, int16_t *__restrict__ acc) {
    const int lane = threadIdx.x & 31;
    const int s = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (s >= S) return;
    int16_t *row = acc + (size_t)s * L1;
    for (int c0 = lane; c0 < L1; c0 += 32) {  // a lane owns columns lane, lane + 32, ...
        int v = refresh ? (int)ft_b[c0] : (int)row[c0];
        if (rem_off)
            for (int k = rem_off[s]; k < rem_off[s + 1]; ++k) {
                const int f = rem_idx[k];
                if (f >= 0 && f < F) v -= (int)__ldg(ft_w + (size_t)f * L1p + c0);
            }
        if (add_off)
            for (int k = add_off[s]; k < add_off[s + 1]; ++k) {
                const int f = add_idx[k];
                if (f >= 0 && f < F) v += (int)__ldg(ft_w + (size_t)f * L1p + c0);
            }
        row[c0] = (int16_t)v;  // exact int arithmetic, then the wrap the engine's int16 adds perform step by step
    }
}

struct QLegacy {
    int L1, L2, L3, qone;
    float l1_scale, l2_scale, out_scale, l1_fact_scale;
    const int8_t *w1, *wf, *w2, *wo;
    const int32_t *b1, *b2;
    int32_t bf, bo0;
};
__device__ __forceinline__ int dense_out(int acc, float scale) {  // simd_scalar.cpp:131-133: float divide, truncate, 0..127
    return max(0, min(127, __float2int_rz(__fdiv_rn(__int2float_rn(acc), scale))));
}
// simd_avx2.cpp:114-152 -- the form LayerStack::forward takes on every AVX2 host (nnue_engine.cpp:393-397, 453-457),
// i.e. what the reference engine computes wherever it is built today: the accumulator vector starts as
// set1_epi32(bias) and all eight lanes are summed, so the bias counts EIGHT times, and the quotient is an integer
// division.  Reproduced as is (the parity tests compare against exactly that build of the reference engine).
__device__ __forceinline__ int dense_out_avx2(int dot, int bias, float scale) {
    const int iscale = (int)scale;
    return max(0, min(127, (8 * bias + dot) / (iscale ? iscale : 1)));
}
// score[s] = LayerStack::forward(clipped accumulator) (nnue_engine.cpp:382-478): combined layer (L2 + 1 outputs),
// row L2 of the factoriser, squared / linear expansion, L2 layer over 2 * L2 inputs, output row 0; a warp per stream
__global__ void __launch_bounds__(256)
q_acc_score_kernel(int S, const QLegacy q, const int16_t *__restrict__ acc, float *__restrict__ score) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * (blockDim.x >> 5) + warp;
    int16_t *in = reinterpret_cast<int16_t *>(smem_raw) + (size_t)warp * (q.L1 + 3 * q.L2 + q.L3 + 8);
    int16_t *comb = in + q.L1;          // [L2 + 1]
    int16_t *expd = comb + q.L2 + 1;    // [2 * L2]
    int16_t *l2o = expd + 2 * q.L2;     // [L3]
    if (s >= S) return;
    for (int i = lane; i < q.L1; i += 32) in[i] = (int16_t)max(0, min(q.qone, (int)acc[(size_t)s * q.L1 + i]));
    __syncwarp();
    for (int o = lane; o <= q.L2; o += 32) {
        int a = 0;
        for (int i = 0; i < q.L1; ++i) a += (int)in[i] * (int)__ldg(q.w1 + (size_t)o * q.L1 + i);
        comb[o] = (int16_t)dense_out_avx2(a, q.b1[o], q.l1_scale);
    }
    int af = 0;  // factoriser row L2: lanes split the inputs
    for (int i = lane; i < q.L1; i += 32) af += (int)in[i] * (int)__ldg(q.wf + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) af += __shfl_xor_sync(kFull, af, o);
    __syncwarp();
    for (int i = lane; i < q.L2; i += 32) {
        const int c = comb[i];
        expd[i] = (int16_t)max(0, min(127, (c * c * 127) / 128));
        expd[i + q.L2] = (int16_t)c;
    }
    __syncwarp();
    for (int o = lane; o < q.L3; o += 32) {
        int a = 0;
        for (int i = 0; i < 2 * q.L2; ++i) a += (int)expd[i] * (int)__ldg(q.w2 + (size_t)o * 2 * q.L2 + i);
        l2o[o] = (int16_t)dense_out_avx2(a, q.b2[o], q.l2_scale);
    }
    __syncwarp();
    if (lane == 0) {
        int a = q.bo0;
        for (int j = 0; j < q.L3; ++j) a += (int)l2o[j] * (int)__ldg(q.wo + j);
        const float l3c = __fdiv_rn(__int2float_rn(a), q.out_scale);
        const float l1f = __fdiv_rn(__int2float_rn(dense_out(q.bf + af, q.l1_fact_scale)), q.l1_fact_scale);
        const float l1c = __fdiv_rn(__int2float_rn((int)comb[q.L2]), q.l1_scale);
        score[s] = __fadd_rn(__fadd_rn(l3c, l1f), l1c);  // nnue_engine.cpp:477
    }
}

}  // namespace nnue

extern "C" {

int nnue_q_acc_apply(const nnue_qmodel *m, int S, int refresh, const int32_t *add_off_d, const int32_t *add_idx_d,
                     const int32_t *rem_off_d, const int32_t *rem_idx_d, int16_t *acc_d, void *stream) {
    if (!m || S < 1 || !acc_d || (add_off_d && !add_idx_d) || (rem_off_d && !rem_idx_d)) return NNUE_ERR_INVALID_ARG;
    q_acc_apply_kernel<<<ceil_div(S, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        S, m->F, m->L1, m->L1p, m->ft_w, m->ft_b, refresh ? 1 : 0, add_off_d, add_idx_d, rem_off_d, rem_idx_d, acc_d);
    NNUE_CHECK_LAUNCH("q_acc_apply_kernel");
    return NNUE_OK;
}

int nnue_q_acc_score(const nnue_qmodel *m, int S, const int16_t *acc_d, int bucket, float *score_d, void *stream) {
    if (!m || S < 1 || !acc_d || !score_d || bucket < 0) return NNUE_ERR_INVALID_ARG;
    if (bucket >= m->n_buckets) bucket = 0;  // nnue_engine.cpp:741-743
    const nnue_qmodel::Stack &st = m->stacks[(size_t)bucket];
    if (!st.w1_raw) return NNUE_ERR_UNSUPPORTED;
    QLegacy q{};
    q.L1 = m->L1; q.L2 = m->L2; q.L3 = m->L3; q.qone = (int)(int16_t)m->quantized_one;
    q.l1_scale = st.l1_scale; q.l2_scale = st.l2_scale; q.out_scale = st.out_scale; q.l1_fact_scale = st.l1_fact_scale;
    q.w1 = st.w1_raw; q.wf = st.wf_row; q.w2 = st.w2_raw; q.wo = st.wo_row; q.b1 = st.b1_raw; q.b2 = st.b2;
    q.bf = st.bf_L2; q.bo0 = st.bo0;
    const size_t smem = 8 * (size_t)(q.L1 + 3 * q.L2 + q.L3 + 8) * 2;
    if (smem > 200 * 1024) return NNUE_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        NNUE_CUDA_TRY(cudaFuncSetAttribute(q_acc_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    q_acc_score_kernel<<<ceil_div(S, 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(S, q, acc_d, score_d);
    NNUE_CHECK_LAUNCH("q_acc_score_kernel");
    return NNUE_OK;
}

int nnue_q_infer_host(const nnue_qmodel *m, const float *images_h, int B, int H, int W, int bucket, float *logits_h,
                      float *density_h) {
    if (!m || !images_h || !logits_h || B < 1 || H < 1 || W < 1) return NNUE_ERR_INVALID_ARG;
    const size_t img_bytes = (size_t)B * H * W * 3 * 4;
    if (!m->h_stream) NNUE_CUDA_TRY(cudaStreamCreateWithFlags(&m->h_stream, cudaStreamNonBlocking));
    if (img_bytes > m->h_cap_img) {
        if (m->h_img) { cudaFree(m->h_img); m->h_img = nullptr; m->h_cap_img = 0; }
        NNUE_CUDA_TRY(cudaMalloc(&m->h_img, img_bytes));
        m->h_cap_img = img_bytes;
    }
    if ((size_t)B > m->h_cap_b) {
        if (m->h_logits) { cudaFree(m->h_logits); cudaFree(m->h_density); m->h_logits = nullptr; m->h_cap_b = 0; }
        NNUE_CUDA_TRY(cudaMalloc(&m->h_logits, (size_t)B * m->NC * 4));
        NNUE_CUDA_TRY(cudaMalloc(&m->h_density, (size_t)B * 4));
        m->h_cap_b = (size_t)B;
    }
    const size_t ws_bytes = nnue_q_workspace_bytes(m, B);
    if (ws_bytes > m->s_cap) {
        if (m->s_ws) { cudaFree(m->s_ws); m->s_ws = nullptr; m->s_cap = 0; }
        NNUE_CUDA_TRY(cudaMalloc(&m->s_ws, ws_bytes));
        m->s_cap = ws_bytes;
    }
    NNUE_CUDA_TRY(cudaMemcpyAsync(m->h_img, images_h, img_bytes, cudaMemcpyHostToDevice, m->h_stream));
    const int rc = nnue_q_infer_ws(m, m->h_img, B, H, W, bucket, m->h_logits, m->h_density, m->s_ws, m->s_cap, m->h_stream);
    if (rc != NNUE_OK) return rc;
    NNUE_CUDA_TRY(cudaMemcpyAsync(logits_h, m->h_logits, (size_t)B * m->NC * 4, cudaMemcpyDeviceToHost, m->h_stream));
    if (density_h)
        NNUE_CUDA_TRY(cudaMemcpyAsync(density_h, m->h_density, (size_t)B * 4, cudaMemcpyDeviceToHost, m->h_stream));
    NNUE_CUDA_TRY(cudaStreamSynchronize(m->h_stream));
    return NNUE_OK;
}

}  // extern "C"
