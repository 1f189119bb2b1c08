// head_fused.cu -- the training step of the head in ONE kernel for small layer stacks:
//   pairwise product -> Linear/ReLU -> Linear/ReLU -> Linear -> mean cross-entropy
//   -> gradient of every head parameter and of the feature-transformer output.
// (nnue.py:660-669, 713-738; train.py:250-254.)  The stack of the benchmark configuration is
// 64 -> 32 -> 8 -> 10: 2.4 k weights, 5 kFLOP per sample.  That is far too little arithmetic for a
// tensor-core tile and the separate layer kernels spend their time on launches and on writing and
// re-reading activations, so here one thread carries one sample through forward, loss and backward
// entirely in registers (weights are broadcast LDS.128 from shared memory), and only the
// parameter gradients, which sum over samples, go through shared memory: the CTA parks its
// 128 samples' activations / activation gradients as tiles and contracts them with a register-tiled
// outer-product loop.  Per-CTA partial gradients are folded in CTA order by a second kernel
// (deterministic, no atomics).  fp32 FMA throughout (the 1e-5 parity bar rules out bf16 / tf32).
//
// Compile-time maxima (L1P, L2P, L3P, NCP) with zero padding serve every smaller stack.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

// Compiler-only fence: keeps ptxas from hoisting the (cheap, broadcast) weight loads of later loop
// iterations above the current one and then spilling them.
#define NNUE_SCHED_FENCE() asm volatile("" ::: "memory")

struct HeadTrainArgs {
    int B, L1, L2, L3, NC;
    const float *ft_out;        // [B, L1]
    const int64_t *labels;      // [B]
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    float inv_count;
    float *g_ft;                // [B, L1]
    float *partial;             // [grid][kHeadPartial]
};

template <int L1P, int L2P, int L3P, int NCP>
struct HeadLayout {
    static constexpr int HP = L1P / 2;
    // weights (floats)
    static constexpr int oW1 = 0, oB1 = oW1 + L2P * L1P, oW2 = oB1 + L2P, oB2 = oW2 + L3P * L2P,
                         oW3 = oB2 + L3P, oB3 = oW3 + NCP * L3P, nW = oB3 + NCP;
    // tiles, row strides padded by 4 floats: 16-byte aligned rows, conflict-free STS.128 by row owners
    static constexpr int sL0 = L1P + 4, sG1 = L2P + 4, sA1 = L2P + 4, sG2 = L3P + 4, sA2 = L3P + 4, sGL = NCP + 4;
    static constexpr int oL0 = (nW + 3) / 4 * 4, oG1 = oL0 + kHeadTile * sL0, oA1 = oG1 + kHeadTile * sG1,
                         oG2 = oA1 + kHeadTile * sA1, oA2 = oG2 + kHeadTile * sG2, oGL = oA2 + kHeadTile * sA2,
                         oRed = oGL + kHeadTile * sGL, oAcc = oRed + 8, total = oAcc + kHeadPartial;
    // per-CTA partial gradient block
    static constexpr int pW1 = 0, pB1 = pW1 + L2P * L1P, pW2 = pB1 + L2P, pB2 = pW2 + L3P * L2P, pW3 = pB2 + L3P,
                         pB3 = pW3 + NCP * L3P, pLoss = pB3 + NCP, pTotal = (pLoss + 1 + 3) / 4 * 4;
};

template <int L1P, int L2P, int L3P, int NCP>
__global__ void __launch_bounds__(kHeadTile, 1)
head_train_kernel(const HeadTrainArgs a) {
    using Lay = HeadLayout<L1P, L2P, L3P, NCP>;
    constexpr int HP = Lay::HP;
    static_assert(kHeadTile == 128 && L1P == 64 && L2P == 32 && L3P == 8 && NCP == 16,
                  "the parameter-gradient thread mapping below is written for a 128-thread CTA and 64/32/8/16");
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int h = a.L1 / 2;

    // ---- stage the weights, zero padded; W1 columns follow the register layout of l0:
    //      slot i < HP = product a_i * b_i (real column i), slot HP + i = a_i (real column h + i)
    for (int e = tid; e < L2P * L1P; e += kHeadTile) {
        const int o = e / L1P, sl = e % L1P;
        const int i = sl < HP ? sl : sl - HP;
        const int col = sl < HP ? i : h + i;
        sm[Lay::oW1 + e] = (o < a.L2 && i < h) ? a.w1[(size_t)o * a.L1 + col] : 0.0f;
    }
    for (int e = tid; e < L2P; e += kHeadTile) sm[Lay::oB1 + e] = e < a.L2 ? a.b1[e] : 0.0f;
    for (int e = tid; e < L3P * L2P; e += kHeadTile) {
        const int o = e / L2P, i = e % L2P;
        sm[Lay::oW2 + e] = (o < a.L3 && i < a.L2) ? a.w2[(size_t)o * a.L2 + i] : 0.0f;
    }
    for (int e = tid; e < L3P; e += kHeadTile) sm[Lay::oB2 + e] = e < a.L3 ? a.b2[e] : 0.0f;
    for (int e = tid; e < NCP * L3P; e += kHeadTile) {
        const int o = e / L3P, i = e % L3P;
        sm[Lay::oW3 + e] = (o < a.NC && i < a.L3) ? a.w3[(size_t)o * a.L3 + i] : 0.0f;
    }
    for (int e = tid; e < NCP; e += kHeadTile) sm[Lay::oB3 + e] = e < a.NC ? a.b3[e] : 0.0f;
    __syncthreads();

    // parameter-gradient block of this CTA (every element has exactly one owner thread), kept in shared
    // memory across the CTA's tiles so the sample phase has the whole register file
    float *gacc = sm + Lay::oAcc;
    for (int e = tid; e < Lay::pTotal; e += kHeadTile) gacc[e] = 0.0f;
    float loss_acc = 0.0f;

    const int ntiles = ceil_div(a.B, kHeadTile);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // =========================== one sample per thread ===========================
        const int b = tile * kHeadTile + tid;
        const bool live = b < a.B;
        float av[HP], bv[HP];  // the two halves of the feature-transformer output
        {
            const float *row = a.ft_out + (size_t)(live ? b : 0) * a.L1;
            const bool vec = (a.L1 == L1P);
            if (vec) {
#pragma unroll
                for (int v = 0; v < HP / 4; ++v) {
                    const float4 x = __ldg(reinterpret_cast<const float4 *>(row) + v);
                    const float4 y = __ldg(reinterpret_cast<const float4 *>(row + HP) + v);
                    av[4 * v] = x.x; av[4 * v + 1] = x.y; av[4 * v + 2] = x.z; av[4 * v + 3] = x.w;
                    bv[4 * v] = y.x; bv[4 * v + 1] = y.y; bv[4 * v + 2] = y.z; bv[4 * v + 3] = y.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < HP; ++i) {
                    av[i] = i < h ? __ldg(row + i) : 0.0f;
                    bv[i] = i < h ? __ldg(row + h + i) : 0.0f;
                }
            }
            if (!live) {
#pragma unroll
                for (int i = 0; i < HP; ++i) av[i] = bv[i] = 0.0f;
            }
        }
        // l0 tile row (product slots, then the a slots)
        {
            float4 *dst = reinterpret_cast<float4 *>(sm + Lay::oL0 + tid * Lay::sL0);
#pragma unroll
            for (int v = 0; v < HP / 4; ++v)
                dst[v] = make_float4(av[4 * v] * bv[4 * v], av[4 * v + 1] * bv[4 * v + 1], av[4 * v + 2] * bv[4 * v + 2],
                                     av[4 * v + 3] * bv[4 * v + 3]);
#pragma unroll
            for (int v = 0; v < HP / 4; ++v) dst[HP / 4 + v] = make_float4(av[4 * v], av[4 * v + 1], av[4 * v + 2], av[4 * v + 3]);
        }
        // layer 1: four outputs at a time, post-ReLU activations go straight to my row of the A1 tile
        {
            float4 *arow = reinterpret_cast<float4 *>(sm + Lay::oA1 + tid * Lay::sA1);
#pragma unroll 1
            for (int o4 = 0; o4 < L2P / 4; ++o4) {
                float z[4];
#pragma unroll
                for (int oo = 0; oo < 4; ++oo) {
                    const int o = o4 * 4 + oo;
                    const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW1 + o * L1P);
                    float z0 = sm[Lay::oB1 + o], z1 = 0.0f;
#pragma unroll
                    for (int v = 0; v < HP / 4; ++v) {
                        const float4 wp = w[v], wa = w[HP / 4 + v];
                        z0 = fmaf(av[4 * v] * bv[4 * v], wp.x, z0);
                        z0 = fmaf(av[4 * v + 1] * bv[4 * v + 1], wp.y, z0);
                        z0 = fmaf(av[4 * v + 2] * bv[4 * v + 2], wp.z, z0);
                        z0 = fmaf(av[4 * v + 3] * bv[4 * v + 3], wp.w, z0);
                        z1 = fmaf(av[4 * v], wa.x, z1);
                        z1 = fmaf(av[4 * v + 1], wa.y, z1);
                        z1 = fmaf(av[4 * v + 2], wa.z, z1);
                        z1 = fmaf(av[4 * v + 3], wa.w, z1);
                    }
                    z[oo] = fmaxf(z0 + z1, 0.0f);
                }
                arow[o4] = make_float4(z[0], z[1], z[2], z[3]);
            }
        }
        // layer 2 (activations of layer 1 come back from my tile row)
        float act1[L2P];
        {
            const float4 *arow = reinterpret_cast<const float4 *>(sm + Lay::oA1 + tid * Lay::sA1);
#pragma unroll
            for (int v = 0; v < L2P / 4; ++v) {
                const float4 t = arow[v];
                act1[4 * v] = t.x; act1[4 * v + 1] = t.y; act1[4 * v + 2] = t.z; act1[4 * v + 3] = t.w;
            }
        }
        float act2[L3P];
#pragma unroll
        for (int o = 0; o < L3P; ++o) {
            const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW2 + o * L2P);
            float z = sm[Lay::oB2 + o];
#pragma unroll
            for (int v = 0; v < L2P / 4; ++v) {
                const float4 ww = w[v];
                z = fmaf(act1[4 * v], ww.x, z); z = fmaf(act1[4 * v + 1], ww.y, z);
                z = fmaf(act1[4 * v + 2], ww.z, z); z = fmaf(act1[4 * v + 3], ww.w, z);
            }
            act2[o] = fmaxf(z, 0.0f);
            NNUE_SCHED_FENCE();
        }
        // output layer + mean cross-entropy and its gradient
        float gl[NCP];
        {
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NCP; ++c) {
                const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW3 + c * L3P);
                float z = sm[Lay::oB3 + c];
#pragma unroll
                for (int v = 0; v < L3P / 4; ++v) {
                    const float4 ww = w[v];
                    z = fmaf(act2[4 * v], ww.x, z); z = fmaf(act2[4 * v + 1], ww.y, z);
                    z = fmaf(act2[4 * v + 2], ww.z, z); z = fmaf(act2[4 * v + 3], ww.w, z);
                }
                gl[c] = c < a.NC ? z : -INFINITY;
                mx = fmaxf(mx, gl[c]);
                if (c % 4 == 3) NNUE_SCHED_FENCE();
            }
            const int y = live ? min(max((int)a.labels[b], 0), a.NC - 1) : 0;
            float se = 0.0f, ly = 0.0f;
#pragma unroll
            for (int c = 0; c < NCP; ++c) {
                if (c == y) ly = gl[c];
                gl[c] = expf(gl[c] - mx);  // exp(-inf) = 0 for the padded classes
                se += gl[c];
            }
            if (live) loss_acc += (mx + logf(se)) - ly;
            const float inv = live ? a.inv_count / se : 0.0f;
#pragma unroll
            for (int c = 0; c < NCP; ++c) gl[c] = live ? fmaf(gl[c], inv, c == y ? -a.inv_count : 0.0f) : 0.0f;
        }
        // backward through the output layer and layer 2
        float g2[L3P];
#pragma unroll
        for (int k = 0; k < L3P; ++k) g2[k] = 0.0f;
#pragma unroll
        for (int c = 0; c < NCP; ++c) {
            const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW3 + c * L3P);
#pragma unroll
            for (int v = 0; v < L3P / 4; ++v) {
                const float4 ww = w[v];
                g2[4 * v] = fmaf(gl[c], ww.x, g2[4 * v]); g2[4 * v + 1] = fmaf(gl[c], ww.y, g2[4 * v + 1]);
                g2[4 * v + 2] = fmaf(gl[c], ww.z, g2[4 * v + 2]); g2[4 * v + 3] = fmaf(gl[c], ww.w, g2[4 * v + 3]);
            }
            if (c % 4 == 3) NNUE_SCHED_FENCE();
        }
#pragma unroll
        for (int k = 0; k < L3P; ++k) g2[k] = act2[k] > 0.0f ? g2[k] : 0.0f;
        float g1[L2P];
#pragma unroll
        for (int o = 0; o < L2P; ++o) g1[o] = 0.0f;
#pragma unroll
        for (int k = 0; k < L3P; ++k) {
            const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW2 + k * L2P);
#pragma unroll
            for (int v = 0; v < L2P / 4; ++v) {
                const float4 ww = w[v];
                g1[4 * v] = fmaf(g2[k], ww.x, g1[4 * v]); g1[4 * v + 1] = fmaf(g2[k], ww.y, g1[4 * v + 1]);
                g1[4 * v + 2] = fmaf(g2[k], ww.z, g1[4 * v + 2]); g1[4 * v + 3] = fmaf(g2[k], ww.w, g1[4 * v + 3]);
            }
            NNUE_SCHED_FENCE();
        }
#pragma unroll
        for (int o = 0; o < L2P; ++o) g1[o] = act1[o] > 0.0f ? g1[o] : 0.0f;
        // park the rows the parameter gradients contract over
        {
            float4 *d = reinterpret_cast<float4 *>(sm + Lay::oG1 + tid * Lay::sG1);
#pragma unroll
            for (int v = 0; v < L2P / 4; ++v) d[v] = make_float4(g1[4 * v], g1[4 * v + 1], g1[4 * v + 2], g1[4 * v + 3]);
            float4 *f = reinterpret_cast<float4 *>(sm + Lay::oG2 + tid * Lay::sG2);
            float4 *g = reinterpret_cast<float4 *>(sm + Lay::oA2 + tid * Lay::sA2);
#pragma unroll
            for (int v = 0; v < L3P / 4; ++v) {
                f[v] = make_float4(g2[4 * v], g2[4 * v + 1], g2[4 * v + 2], g2[4 * v + 3]);
                g[v] = make_float4(act2[4 * v], act2[4 * v + 1], act2[4 * v + 2], act2[4 * v + 3]);
            }
            float4 *l = reinterpret_cast<float4 *>(sm + Lay::oGL + tid * Lay::sGL);
#pragma unroll
            for (int v = 0; v < NCP / 4; ++v) l[v] = make_float4(gl[4 * v], gl[4 * v + 1], gl[4 * v + 2], gl[4 * v + 3]);
        }
        // backward through layer 1 and the pairwise product: g_ft row
        {
            float gp[HP], ga[HP];  // gradient of the product slots / of the a slots
#pragma unroll
            for (int i = 0; i < HP; ++i) gp[i] = ga[i] = 0.0f;
            const float4 *grow = reinterpret_cast<const float4 *>(sm + Lay::oG1 + tid * Lay::sG1);
#pragma unroll 1
            for (int o4 = 0; o4 < L2P / 4; ++o4) {
                const float4 g4 = grow[o4];
                const float gq[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int oo = 0; oo < 4; ++oo) {
                    const float4 *w = reinterpret_cast<const float4 *>(sm + Lay::oW1 + (o4 * 4 + oo) * L1P);
                    const float go = gq[oo];
#pragma unroll
                    for (int v = 0; v < HP / 4; ++v) {
                        const float4 wp = w[v], wa = w[HP / 4 + v];
                        gp[4 * v] = fmaf(go, wp.x, gp[4 * v]); gp[4 * v + 1] = fmaf(go, wp.y, gp[4 * v + 1]);
                        gp[4 * v + 2] = fmaf(go, wp.z, gp[4 * v + 2]); gp[4 * v + 3] = fmaf(go, wp.w, gp[4 * v + 3]);
                        ga[4 * v] = fmaf(go, wa.x, ga[4 * v]); ga[4 * v + 1] = fmaf(go, wa.y, ga[4 * v + 1]);
                        ga[4 * v + 2] = fmaf(go, wa.z, ga[4 * v + 2]); ga[4 * v + 3] = fmaf(go, wa.w, ga[4 * v + 3]);
                    }
                }
            }
            // g_ft[:, i] = g_l0[:, i] * ft[:, i+h] + g_l0[:, i+h];  g_ft[:, i+h] = g_l0[:, i] * ft[:, i]
            if (live) {
                float *row = a.g_ft + (size_t)b * a.L1;
                if (a.L1 == L1P) {
#pragma unroll
                    for (int v = 0; v < HP / 4; ++v) {
                        reinterpret_cast<float4 *>(row)[v] =
                            make_float4(fmaf(gp[4 * v], bv[4 * v], ga[4 * v]), fmaf(gp[4 * v + 1], bv[4 * v + 1], ga[4 * v + 1]),
                                        fmaf(gp[4 * v + 2], bv[4 * v + 2], ga[4 * v + 2]),
                                        fmaf(gp[4 * v + 3], bv[4 * v + 3], ga[4 * v + 3]));
                        reinterpret_cast<float4 *>(row + HP)[v] = make_float4(gp[4 * v] * av[4 * v], gp[4 * v + 1] * av[4 * v + 1],
                                                                              gp[4 * v + 2] * av[4 * v + 2],
                                                                              gp[4 * v + 3] * av[4 * v + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < HP; ++i)
                        if (i < h) {
                            row[i] = fmaf(gp[i], bv[i], ga[i]);
                            row[h + i] = gp[i] * av[i];
                        }
                }
            }
        }
        __syncthreads();
        // =========================== parameter gradients of this tile ===========================
        {
            const int og = tid / 16, ig = tid % 16;        // dW1[4 og .. +3][4 ig .. +3]
            const int o2 = tid / 16, i2 = (tid % 16) * 2;  // dW2[o2][i2, i2 + 1]
            const int c3 = tid / 8, k3 = tid % 8;          // dW3[c3][k3]
            // bias gradients: threads 0..31 -> b1, 32..39 -> b2, 40..55 -> b3
            const float *bsrc = tid < L2P ? sm + Lay::oG1 + tid
                               : tid < L2P + L3P ? sm + Lay::oG2 + (tid - L2P)
                                                 : sm + Lay::oGL + (tid < L2P + L3P + NCP ? tid - L2P - L3P : 0);
            const int bstride = tid < L2P ? Lay::sG1 : tid < L2P + L3P ? Lay::sG2 : Lay::sGL;
            const bool has_bias = tid < L2P + L3P + NCP;
            float dw1[4][4], dw2[2] = {0.0f, 0.0f}, dw3 = 0.0f, dbias = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dw1[i][jj] = 0.0f;
#pragma unroll 4
            for (int r = 0; r < kHeadTile; ++r) {
                const float4 g = *reinterpret_cast<const float4 *>(sm + Lay::oG1 + r * Lay::sG1 + og * 4);
                const float4 x = *reinterpret_cast<const float4 *>(sm + Lay::oL0 + r * Lay::sL0 + ig * 4);
                dw1[0][0] = fmaf(g.x, x.x, dw1[0][0]); dw1[0][1] = fmaf(g.x, x.y, dw1[0][1]);
                dw1[0][2] = fmaf(g.x, x.z, dw1[0][2]); dw1[0][3] = fmaf(g.x, x.w, dw1[0][3]);
                dw1[1][0] = fmaf(g.y, x.x, dw1[1][0]); dw1[1][1] = fmaf(g.y, x.y, dw1[1][1]);
                dw1[1][2] = fmaf(g.y, x.z, dw1[1][2]); dw1[1][3] = fmaf(g.y, x.w, dw1[1][3]);
                dw1[2][0] = fmaf(g.z, x.x, dw1[2][0]); dw1[2][1] = fmaf(g.z, x.y, dw1[2][1]);
                dw1[2][2] = fmaf(g.z, x.z, dw1[2][2]); dw1[2][3] = fmaf(g.z, x.w, dw1[2][3]);
                dw1[3][0] = fmaf(g.w, x.x, dw1[3][0]); dw1[3][1] = fmaf(g.w, x.y, dw1[3][1]);
                dw1[3][2] = fmaf(g.w, x.z, dw1[3][2]); dw1[3][3] = fmaf(g.w, x.w, dw1[3][3]);
                const float gg2 = sm[Lay::oG2 + r * Lay::sG2 + o2];
                const float2 a1 = *reinterpret_cast<const float2 *>(sm + Lay::oA1 + r * Lay::sA1 + i2);
                dw2[0] = fmaf(gg2, a1.x, dw2[0]);
                dw2[1] = fmaf(gg2, a1.y, dw2[1]);
                dw3 = fmaf(sm[Lay::oGL + r * Lay::sGL + c3], sm[Lay::oA2 + r * Lay::sA2 + k3], dw3);
                if (has_bias) dbias += bsrc[r * bstride];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 *d = reinterpret_cast<float4 *>(gacc + Lay::pW1 + (og * 4 + i) * L1P + ig * 4);
                const float4 c = *d;
                *d = make_float4(c.x + dw1[i][0], c.y + dw1[i][1], c.z + dw1[i][2], c.w + dw1[i][3]);
            }
            gacc[Lay::pW2 + o2 * L2P + i2] += dw2[0];
            gacc[Lay::pW2 + o2 * L2P + i2 + 1] += dw2[1];
            gacc[Lay::pW3 + c3 * L3P + k3] += dw3;
            if (tid < L2P) gacc[Lay::pB1 + tid] += dbias;
            else if (tid < L2P + L3P) gacc[Lay::pB2 + tid - L2P] += dbias;
            else if (has_bias) gacc[Lay::pB3 + tid - L2P - L3P] += dbias;
        }
        __syncthreads();  // tiles are rewritten by the next iteration
    }

    // ---- per-CTA partial block -------------------------------------------------------------------
    float *out = a.partial + (size_t)blockIdx.x * Lay::pTotal;
    for (int e = tid; e < Lay::pLoss; e += kHeadTile) out[e] = gacc[e];
    // loss: fixed-order sum over the CTA
    float v = warp_sum(loss_acc);
    if ((tid & 31) == 0) sm[Lay::oRed + (tid >> 5)] = v;
    __syncthreads();
    if (tid == 0) out[Lay::pLoss] = (sm[Lay::oRed] + sm[Lay::oRed + 1]) + (sm[Lay::oRed + 2] + sm[Lay::oRed + 3]);
}

// Sum the per-CTA blocks in CTA order and scatter them into the unpadded parameter gradients.
template <int L1P, int L2P, int L3P, int NCP>
__global__ void head_train_fold_kernel(int nblk, const float *__restrict__ partial, int L1, int L2, int L3, int NC,
                                       float inv_count, float *g_w1, float *g_b1, float *g_w2, float *g_b2, float *g_w3,
                                       float *g_b3, float *loss) {
    using Lay = HeadLayout<L1P, L2P, L3P, NCP>;
    // 256 threads = 32 elements x 8 slices of the CTA list (slice q: CTAs q, q + 8, ...): every thread has all its loads
    // in flight at once (the fold is pure latency); slices are combined in slice order -- a fixed order per shape
    __shared__ float red[8][32];
    const int e = blockIdx.x * 32 + (threadIdx.x & 31), q = threadIdx.x >> 5;
    float v = 0.0f;
    if (e <= Lay::pLoss) {
        float t[20];
        int k = q;
        while (k < nblk) {
#pragma unroll
            for (int u = 0; u < 20; ++u) t[u] = k + 8 * u < nblk ? __ldg(partial + (size_t)(k + 8 * u) * Lay::pTotal + e) : 0.0f;
#pragma unroll
            for (int u = 0; u < 20; ++u) v += t[u];
            k += 160;
        }
    }
    red[q][threadIdx.x & 31] = v;
    __syncthreads();
    if (q != 0 || e > Lay::pLoss) return;
    for (int s2 = 1; s2 < 8; ++s2) v += red[s2][threadIdx.x & 31];
    const int h = L1 / 2;
    if (e < Lay::pB1) {
        const int o = e / L1P, sl = e % L1P;
        const int i = sl < Lay::HP ? sl : sl - Lay::HP;
        if (o < L2 && i < h) g_w1[(size_t)o * L1 + (sl < Lay::HP ? i : h + i)] = v;
    } else if (e < Lay::pW2) {
        if (e - Lay::pB1 < L2) g_b1[e - Lay::pB1] = v;
    } else if (e < Lay::pB2) {
        const int o = (e - Lay::pW2) / L2P, i = (e - Lay::pW2) % L2P;
        if (o < L3 && i < L2) g_w2[(size_t)o * L2 + i] = v;
    } else if (e < Lay::pW3) {
        if (e - Lay::pB2 < L3) g_b2[e - Lay::pB2] = v;
    } else if (e < Lay::pB3) {
        const int o = (e - Lay::pW3) / L3P, i = (e - Lay::pW3) % L3P;
        if (o < NC && i < L3) g_w3[(size_t)o * L3 + i] = v;
    } else if (e < Lay::pLoss) {
        if (e - Lay::pB3 < NC) g_b3[e - Lay::pB3] = v;
    } else if (loss) {
        loss[0] = v * inv_count;
    }
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_head_is_fused(const nnue_shape *s) { return s && head_train_fused_ok(*s) ? 1 : 0; }
int nnue_head_uses_umma(const nnue_shape *s) { return s && !head_train_fused_ok(*s) && head_umma_ok(*s) ? 1 : 0; }

size_t nnue_head_side_workspace_bytes(const nnue_shape *s) { return s ? ws_head_side(*s) : 0; }

int nnue_head_train(const nnue_shape *s, const float *ft_out_d, const int64_t *labels_d, float inv_count,
                    const float *w1_d, const float *b1_d, const float *w2_d, const float *b2_d, const float *w3_d,
                    const float *b3_d, float *loss_d, float *g_ft_d, float *g_w1_d, float *g_b1_d, float *g_w2_d,
                    float *g_b2_d, float *g_w3_d, float *g_b3_d, void *workspace_d, size_t workspace_bytes,
                    void *stream) {
    return nnue_head_train_overlapped(s, ft_out_d, labels_d, inv_count, w1_d, b1_d, w2_d, b2_d, w3_d, b3_d, loss_d, g_ft_d,
                                      g_w1_d, g_b1_d, g_w2_d, g_b2_d, g_w3_d, g_b3_d, workspace_d, workspace_bytes, stream,
                                      nullptr, 0, nullptr);
}

int nnue_head_train_overlapped(const nnue_shape *s, const float *ft_out_d, const int64_t *labels_d, float inv_count,
                               const float *w1_d, const float *b1_d, const float *w2_d, const float *b2_d, const float *w3_d,
                               const float *b3_d, float *loss_d, float *g_ft_d, float *g_w1_d, float *g_b1_d, float *g_w2_d,
                               float *g_b2_d, float *g_w3_d, float *g_b3_d, void *workspace_d, size_t workspace_bytes,
                               void *stream, void *side_workspace_d, size_t side_workspace_bytes, void *side_stream) {
    if (!s || !ft_out_d || !labels_d || !w1_d || !b1_d || !w2_d || !b2_d || !w3_d || !b3_d || !loss_d || !g_ft_d ||
        !g_w1_d || !g_b1_d || !g_w2_d || !g_b2_d || !g_w3_d || !g_b3_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_head_train(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (head_train_fused_ok(*s)) {
        using Lay = HeadLayout<64, 32, 8, 16>;
        static_assert(Lay::pTotal == kHeadPartial, "plan.cuh kHeadPartial must match the partial block layout");
        HeadTrainArgs a{};
        a.B = s->B; a.L1 = s->L1; a.L2 = s->L2; a.L3 = s->L3; a.NC = s->NC;
        a.ft_out = ft_out_d; a.labels = labels_d;
        a.w1 = w1_d; a.b1 = b1_d; a.w2 = w2_d; a.b2 = b2_d; a.w3 = w3_d; a.b3 = b3_d;
        a.inv_count = inv_count;
        a.g_ft = g_ft_d;
        a.partial = static_cast<float *>(workspace_d);
        const int grid = head_train_grid(*s);
        const size_t smem = (size_t)Lay::total * sizeof(float);
        auto k = head_train_kernel<64, 32, 8, 16>;
        NNUE_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kHeadTile, smem, st>>>(a);
        NNUE_CHECK_LAUNCH("head_train_kernel");
        head_train_fold_kernel<64, 32, 8, 16><<<ceil_div(Lay::pLoss + 1, 32), 256, 0, st>>>(
            grid, a.partial, s->L1, s->L2, s->L3, s->NC, inv_count, g_w1_d, g_b1_d, g_w2_d, g_b2_d, g_w3_d, g_b3_d,
            loss_d);
        NNUE_CHECK_LAUNCH("head_train_fold_kernel");
        return NNUE_OK;
    }
    char *ws = static_cast<char *>(workspace_d);
    auto carve = [&](size_t bytes) { float *p = reinterpret_cast<float *>(ws); ws += align_up(bytes, 256); return p; };
    const size_t B = s->B;
    if (head_mid_ok(*s)) {
        // wide first layer, small tail: layer 1 forward (tensor cores when wide), then ONE kernel for layers 2 - 3, the
        // loss and their gradients (head_mid.cu), then layer 1 backward.  act1 | per-CTA partials | backward scratch | forward scratch
        float *act1 = carve(B * s->L2 * 4);
        float *midp = carve((size_t)head_mid_grid(*s) * kHeadMidPartial * 4);
        const size_t rest = workspace_bytes - (size_t)(ws - static_cast<char *>(workspace_d));
        const size_t fwd_ws = ws_head_umma_fwd(*s);
        void *fwd_scratch = fwd_ws && rest >= ws_head_bwd(*s) + fwd_ws ? ws + align_up(ws_head_bwd(*s), 256) : nullptr;
        int rc = head_layer1_fwd(s, ft_out_d, w1_d, b1_d, act1, fwd_scratch, fwd_scratch ? fwd_ws : 0, st);
        if (rc != NNUE_OK) return rc;
        HeadBwdWs bw = carve_head_bwd(*s, ws);
        // with a side stream (and enough side scratch) the layer-1 weight-gradient chain leaves the caller's stream: g_z1
        // then lives in the side scratch, which no later stage of the caller's stream writes
        HeadSide side{};
        const bool use_side = side_stream && side_workspace_d && ws_head_side(*s) > 0 && side_workspace_bytes >= ws_head_side(*s);
        if (use_side) {
            static thread_local cudaEvent_t ev[16] = {};  // one per device, reused by every call (a record replaces the last)
            int dev = 0;
            NNUE_CUDA_TRY(cudaGetDevice(&dev));
            if (dev < 0 || dev >= 16) return NNUE_ERR_UNSUPPORTED;
            if (!ev[dev]) NNUE_CUDA_TRY(cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming));
            side.stream = static_cast<cudaStream_t>(side_stream);
            side.ready = ev[dev];
            bw.g_act1 = static_cast<float *>(side_workspace_d);
            side.ws = static_cast<char *>(side_workspace_d) + align_up(B * s->L2 * 4, 256);
            side.ws_bytes = side_workspace_bytes - align_up(B * s->L2 * 4, 256);
        }
        rc = launch_head_mid(*s, act1, labels_d, inv_count, w2_d, b2_d, w3_d, b3_d, loss_d, bw.g_act1, g_w2_d, g_b2_d, g_w3_d,
                             g_b3_d, g_b1_d, midp, st);
        if (rc != NNUE_OK) return rc;
        // (the tensor-core form takes g_b1 from head_mid; the FMA form's weight-gradient GEMM produces it alongside g_w1)
        return head_bwd_layer1(s, bw, ft_out_d, w1_d, g_w1_d, head_umma_ok(*s) ? nullptr : g_b1_d, g_ft_d, st,
                               use_side ? &side : nullptr);
    }
    // other stacks: the layer kernels of head.cu, with the activations in scratch
    float *act1 = carve(B * s->L2 * 4), *act2 = carve(B * s->L3 * 4), *logits = carve(B * s->NC * 4);
    float *g_logits = carve(B * s->NC * 4), *per = carve(B * 4);
    const size_t rest = workspace_bytes - (size_t)(ws - static_cast<char *>(workspace_d));
    // (the forward's tensor-core scratch sits behind the backward's: both fit in ws_head_train)
    const size_t fwd_ws = ws_head_umma_fwd(*s);
    void *fwd_scratch = fwd_ws && rest >= ws_head_bwd(*s) + fwd_ws ? ws + align_up(ws_head_bwd(*s), 256) : nullptr;
    int rc = head_fwd_ws(s, ft_out_d, w1_d, b1_d, w2_d, b2_d, w3_d, b3_d, act1, act2, logits, fwd_scratch,
                         fwd_scratch ? fwd_ws : 0, static_cast<cudaStream_t>(stream));
    if (rc != NNUE_OK) return rc;
    rc = nnue_ce_fwd_bwd(s->B, s->NC, logits, labels_d, inv_count, nullptr, loss_d, per, g_logits, nullptr, 0, stream);
    if (rc != NNUE_OK) return rc;
    return nnue_head_bwd(s, g_logits, ft_out_d, act1, act2, w1_d, w2_d, w3_d, g_w1_d, g_b1_d, g_w2_d, g_b2_d, g_w3_d,
                         g_b3_d, g_ft_d, ws, rest, stream);
}

}  // extern "C"
