// head_mid.cu -- layers 2 and 3 of a wide stack, the mean cross-entropy and their whole backward in ONE kernel
// (nnue.py:713-738 `SimpleClassifier` layers l2 / output, train.py:250-254 `F.cross_entropy`, and autograd's nodes for them).
//
// The reference's "real" config (config/train_nnue.py: 1024 -> 128 -> 32 -> 10) keeps its first layer on the tensor cores
// (gemm_umma.cu: a dense 13 GFLOP contraction).  Behind it sit 128 -> 32 -> 10: 4.4 k weights, 13 kFLOP per sample -- no
// tile is worth issuing and the nine separate kernels that used to run them (two skinny GEMMs, CE, two split-K weight
// gradients with their folds, two masked input gradients) took 190 us of the 0.96 ms step at batch 16384, almost all of it
// launch tails and re-reading [B][32] / [B][10] activations.  Here a persistent CTA takes 128 samples at a time:
//
//   act1 tile -> shared memory -> z2 = act1 W2^T + b2 -> ReLU -> logits -> softmax / loss / g_logits
//             -> g_W3, g_b3, g_z2 -> g_W2, g_b2 -> g_z1 = (g_z2 W2) * (act1 > 0) -> global (+ its column sums = g_b1)
//
// with register-tiled fp32 FMA contractions over shared-memory tiles (row strides chosen so that every LDS.128 of a
// quarter warp covers distinct banks).  Parameter gradients accumulate in shared memory across the CTA's tiles (every
// element has one owner thread) and leave as one partial block per CTA, folded in CTA order by a second kernel: no
// atomics, bit-identical run to run.  fp32 throughout (the 1e-5 parity bar).
#include <math.h>

#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kMidTB = 128;       // samples per tile
constexpr int kMidThreads = 256;
constexpr int kMidL2P = 128, kMidL3P = 32, kMidNCP = 16;  // compile-time maxima, zero padded
constexpr int kMidSA = kMidL2P + 4;   // row stride of the act1 tile and of W2 (floats): rows r, r + 1, ... start 4 banks apart
constexpr int kMidSZ = kMidL3P + 1;   // act2 / g_z2 tiles
constexpr int kMidSG = kMidNCP + 1;   // g_logits tile

struct MidLayout {
    static constexpr int oW2 = 0, oW3 = oW2 + kMidL3P * kMidSA, oB2 = oW3 + kMidNCP * kMidSZ, oB3 = oB2 + kMidL3P,
                         oA1 = (oB3 + kMidNCP + 3) / 4 * 4, oZ2 = oA1 + kMidTB * kMidSA, oGZ = oZ2 + kMidTB * kMidSZ,
                         oGL = oGZ + kMidTB * kMidSZ, oAcc = (oGL + kMidTB * kMidSG + 3) / 4 * 4;
    // per-CTA partial block
    static constexpr int pW2 = 0, pB2 = pW2 + kMidL3P * kMidL2P, pW3 = pB2 + kMidL3P, pB3 = pW3 + kMidNCP * kMidL3P,
                         pB1 = pB3 + kMidNCP, pLoss = pB1 + kMidL2P, pTotal = (pLoss + 1 + 3) / 4 * 4;
    static constexpr int total = oAcc + pTotal;
};
static_assert(MidLayout::pTotal == kHeadMidPartial, "plan.cuh kHeadMidPartial must match the partial block layout");

struct HeadMidArgs {
    int B, L2, L3, NC;
    const float *act1;          // [B, L2] post-ReLU output of layer 1
    const int64_t *labels;      // [B]
    const float *w2, *b2, *w3, *b3;
    float inv_count;
    float *g_z1;                // [B, L2] gradient of the loss w.r.t. layer 1's pre-activation
    float *partial;             // [grid][pTotal]
};

__global__ void __launch_bounds__(kMidThreads, 1)
head_mid_kernel(const HeadMidArgs a) {
    using Lay = MidLayout;
    extern __shared__ __align__(16) float sm[];
    float *W2s = sm + Lay::oW2, *W3s = sm + Lay::oW3, *b2s = sm + Lay::oB2, *b3s = sm + Lay::oB3;
    float *A1 = sm + Lay::oA1, *Z2 = sm + Lay::oZ2, *GZ = sm + Lay::oGZ, *GL = sm + Lay::oGL, *acc = sm + Lay::oAcc;
    const int tid = threadIdx.x;

    // ---- weights, zero padded ----
    // (loads of a batch are issued together, then stored: a load -> store loop pays one memory latency per iteration)
    {
        float4 v[4];  // W2 [L3][L2] as float4 (L2 % 4 == 0): 32 rows x 32 chunks = 4 per thread; the 4 pad floats of a row stay unread
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * kMidThreads, o = e / (kMidL2P / 4), c4 = e % (kMidL2P / 4);
            v[u] = (o < a.L3 && 4 * c4 < a.L2) ? __ldg(reinterpret_cast<const float4 *>(a.w2 + (size_t)o * a.L2) + c4)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = tid + u * kMidThreads, o = e / (kMidL2P / 4), c4 = e % (kMidL2P / 4);
            *reinterpret_cast<float4 *>(W2s + o * kMidSA + 4 * c4) = v[u];
        }
        float w[3];   // W3 [NC][L3]: 16 x 33 padded slots = 528 <= 3 per thread
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = tid + u * kMidThreads, c = e / kMidSZ, k = e % kMidSZ;
            w[u] = (e < kMidNCP * kMidSZ && c < a.NC && k < a.L3) ? __ldg(a.w3 + (size_t)c * a.L3 + k) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = tid + u * kMidThreads;
            if (e < kMidNCP * kMidSZ) W3s[e] = w[u];
        }
    }
    if (tid < kMidL3P) b2s[tid] = tid < a.L3 ? __ldg(a.b2 + tid) : 0.0f;
    if (tid < kMidNCP) b3s[tid] = tid < a.NC ? __ldg(a.b3 + tid) : 0.0f;
    for (int e = tid; e < Lay::pTotal; e += kMidThreads) acc[e] = 0.0f;
    float loss_acc = 0.0f;  // threads 0..127: the losses of "their" sample of every tile
    __syncthreads();

    const int sg = tid >> 3, oc = tid & 7;  // the 4 x 4 register tiles below: samples sg + 32 i, outputs oc + 8 j
    const int ntiles = ceil_div(a.B, kMidTB);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b0 = tile * kMidTB;
        // ---- act1 tile (rows past the batch and columns past L2 are zero) ----
#pragma unroll
        for (int e0 = 0; e0 < kMidTB * (kMidL2P / 4); e0 += 8 * kMidThreads) {  // sixteen 16-byte chunks per thread, eight in flight
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * kMidThreads + tid, r = e / (kMidL2P / 4), c4 = e % (kMidL2P / 4);
                v[u] = (b0 + r < a.B && 4 * c4 < a.L2) ? __ldg(reinterpret_cast<const float4 *>(a.act1 + (size_t)(b0 + r) * a.L2) + c4)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * kMidThreads + tid, r = e / (kMidL2P / 4), c4 = e % (kMidL2P / 4);
                *reinterpret_cast<float4 *>(A1 + r * kMidSA + 4 * c4) = v[u];
            }
        }
        __syncthreads();

        // ---- layer 2: act2[s][o] = relu(b2[o] + sum_k act1[s][k] W2[o][k]) ----
        {
            float z[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) z[i][j] = 0.0f;
#pragma unroll 4
            for (int k4 = 0; k4 < kMidL2P / 4; ++k4) {
                float4 av[4], wv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4 *>(A1 + (sg + 32 * i) * kMidSA + 4 * k4);
#pragma unroll
                for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4 *>(W2s + (oc + 8 * j) * kMidSA + 4 * k4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        z[i][j] = fmaf(av[i].x, wv[j].x, z[i][j]);
                        z[i][j] = fmaf(av[i].y, wv[j].y, z[i][j]);
                        z[i][j] = fmaf(av[i].z, wv[j].z, z[i][j]);
                        z[i][j] = fmaf(av[i].w, wv[j].w, z[i][j]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) Z2[(sg + 32 * i) * kMidSZ + oc + 8 * j] = fmaxf(z[i][j] + b2s[oc + 8 * j], 0.0f);
        }
        __syncthreads();

        // ---- output layer + mean cross-entropy: one sample per thread (threads 0..127) ----
        if (tid < kMidTB) {
            const int s = tid, b = b0 + s;
            float x[kMidL3P];
#pragma unroll
            for (int k = 0; k < kMidL3P; ++k) x[k] = Z2[s * kMidSZ + k];
            float lg[kMidNCP];
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < kMidNCP; ++c) {
                float v = b3s[c];
#pragma unroll
                for (int k = 0; k < kMidL3P; ++k) v = fmaf(x[k], W3s[c * kMidSZ + k], v);
                lg[c] = v;
                if (c < a.NC) mx = fmaxf(mx, v);
            }
            float se = 0.0f;
#pragma unroll
            for (int c = 0; c < kMidNCP; ++c)
                if (c < a.NC) se += expf(lg[c] - mx);
            const bool live = b < a.B;
            const int y = live ? min(max((int)a.labels[b], 0), a.NC - 1) : 0;  // (a target outside [0, NC) must not read out of bounds)
            const float inv = 1.0f / se;
            float ly = 0.0f;
#pragma unroll
            for (int c = 0; c < kMidNCP; ++c) {
                if (c == y) ly = lg[c];
                const float p = expf(lg[c] - mx) * inv;
                GL[s * kMidSG + c] = (live && c < a.NC) ? (p - (c == y ? 1.0f : 0.0f)) * a.inv_count : 0.0f;
            }
            if (live) loss_acc += (mx + logf(se)) - ly;
        }
        __syncthreads();

        // ---- g_z2 = (g_logits W3) * (act2 > 0); g_W3 += g_logits^T act2; g_b3 += column sums ----
        {
            const int s = tid >> 1, o0 = (tid & 1) * (kMidL3P / 2);
            float gl[kMidNCP];
#pragma unroll
            for (int c = 0; c < kMidNCP; ++c) gl[c] = GL[s * kMidSG + c];
#pragma unroll
            for (int o = 0; o < kMidL3P / 2; ++o) {
                float v = 0.0f;
#pragma unroll
                for (int c = 0; c < kMidNCP; ++c) v = fmaf(gl[c], W3s[c * kMidSZ + o0 + o], v);
                GZ[s * kMidSZ + o0 + o] = Z2[s * kMidSZ + o0 + o] > 0.0f ? v : 0.0f;
            }
            // owner of g_W3[c][k0], g_W3[c][k0 + 1]
            const int c = tid >> 4, k0 = (tid & 15) * 2;
            float w0 = 0.0f, w1 = 0.0f;
#pragma unroll 8
            for (int r = 0; r < kMidTB; ++r) {
                const float g = GL[r * kMidSG + c];
                w0 = fmaf(g, Z2[r * kMidSZ + k0], w0);
                w1 = fmaf(g, Z2[r * kMidSZ + k0 + 1], w1);
            }
            acc[Lay::pW3 + c * kMidL3P + k0] += w0;
            acc[Lay::pW3 + c * kMidL3P + k0 + 1] += w1;
            if (tid < kMidNCP) {
                float v = 0.0f;
#pragma unroll 8
                for (int r = 0; r < kMidTB; ++r) v += GL[r * kMidSG + tid];
                acc[Lay::pB3 + tid] += v;
            }
        }
        __syncthreads();

        // ---- g_W2[o][k] += sum_s g_z2[s][o] act1[s][k]  (outputs oc + 8 j, columns 4 sg .. 4 sg + 3); g_b2 ----
        {
            float w[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int u = 0; u < 4; ++u) w[j][u] = 0.0f;
#pragma unroll 4
            for (int r = 0; r < kMidTB; ++r) {
                const float4 av = *reinterpret_cast<const float4 *>(A1 + r * kMidSA + 4 * sg);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float g = GZ[r * kMidSZ + oc + 8 * j];
                    w[j][0] = fmaf(g, av.x, w[j][0]);
                    w[j][1] = fmaf(g, av.y, w[j][1]);
                    w[j][2] = fmaf(g, av.z, w[j][2]);
                    w[j][3] = fmaf(g, av.w, w[j][3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 *dst = reinterpret_cast<float4 *>(acc + Lay::pW2 + (oc + 8 * j) * kMidL2P + 4 * sg);
                float4 v = *dst;
                v.x += w[j][0]; v.y += w[j][1]; v.z += w[j][2]; v.w += w[j][3];
                *dst = v;
            }
            if (tid < kMidL3P) {
                float v = 0.0f;
#pragma unroll 8
                for (int r = 0; r < kMidTB; ++r) v += GZ[r * kMidSZ + tid];
                acc[Lay::pB2 + tid] += v;
            }
        }
        // ---- g_z1[s][k] = (act1[s][k] > 0) ? sum_o g_z2[s][o] W2[o][k] : 0  (samples sg + 32 i, columns 4 (oc + 8 j) .. + 3) ----
        {
            float4 g1[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) g1[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int o = 0; o < kMidL3P; ++o) {
                float gz[4];
                float4 wv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) gz[i] = GZ[(sg + 32 * i) * kMidSZ + o];
#pragma unroll
                for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float4 *>(W2s + o * kMidSA + 4 * (oc + 8 * j));
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        g1[i][j].x = fmaf(gz[i], wv[j].x, g1[i][j].x);
                        g1[i][j].y = fmaf(gz[i], wv[j].y, g1[i][j].y);
                        g1[i][j].z = fmaf(gz[i], wv[j].z, g1[i][j].z);
                        g1[i][j].w = fmaf(gz[i], wv[j].w, g1[i][j].w);
                    }
            }
            float4 cs[4];  // column sums of this thread's four samples (the bias gradient of layer 1)
#pragma unroll
            for (int j = 0; j < 4; ++j) cs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s = sg + 32 * i;
                if (b0 + s < a.B) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = 4 * (oc + 8 * j);
                        if (k < a.L2) {
                            const float4 av = *reinterpret_cast<const float4 *>(A1 + s * kMidSA + k);
                            float4 v = g1[i][j];
                            v.x = av.x > 0.0f ? v.x : 0.0f; v.y = av.y > 0.0f ? v.y : 0.0f;
                            v.z = av.z > 0.0f ? v.z : 0.0f; v.w = av.w > 0.0f ? v.w : 0.0f;
                            *reinterpret_cast<float4 *>(a.g_z1 + (size_t)(b0 + s) * a.L2 + k) = v;
                            cs[j] = f4_add(cs[j], v);
                        }
                    }
                }
            }
            // (the act2 tile is dead since the g_z2 phase: it carries the 32 sample groups' column sums to their owners)
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4 *>(Z2 + sg * kMidL2P + 4 * (oc + 8 * j)) = cs[j];
        }
        __syncthreads();  // the next tile overwrites A1 / GZ / GL (and, after its own barrier, Z2)
        if (tid < kMidL2P) {
            float v = 0.0f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) v += Z2[r * kMidL2P + tid];
            acc[Lay::pB1 + tid] += v;
        }
    }

    // ---- the CTA's partial block; the loss in thread order (fixed) ----
    float *red = GL;  // (idle now)
    if (tid < kMidTB) red[tid] = loss_acc;
    __syncthreads();
    if (tid == 0) {
        float v = 0.0f;
        for (int r = 0; r < kMidTB; ++r) v += red[r];
        acc[Lay::pLoss] = v;
    }
    __syncthreads();
    float *out = a.partial + (size_t)blockIdx.x * Lay::pTotal;
    for (int e = tid; e < Lay::pTotal; e += kMidThreads) out[e] = acc[e];
}

// every element of the partial block summed over the CTAs in CTA order, then scattered to the parameter gradients
__global__ void __launch_bounds__(256)
head_mid_fold_kernel(int nblk, const float *__restrict__ partial, int L2, int L3, int NC, float inv_count, float *__restrict__ g_w2,
                     float *__restrict__ g_b2, float *__restrict__ g_w3, float *__restrict__ g_b3, float *__restrict__ g_b1, float *__restrict__ loss) {
    using Lay = MidLayout;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > Lay::pLoss) return;
    float v = 0.0f;
    int k = 0;
    for (; k + 8 <= nblk; k += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldg(partial + (size_t)(k + u) * Lay::pTotal + e);
#pragma unroll
        for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; k < nblk; ++k) v += __ldg(partial + (size_t)k * Lay::pTotal + e);
    if (e < Lay::pB2) {
        const int o = e / kMidL2P, i = e % kMidL2P;
        if (o < L3 && i < L2) g_w2[(size_t)o * L2 + i] = v;
    } else if (e < Lay::pW3) {
        if (e - Lay::pB2 < L3) g_b2[e - Lay::pB2] = v;
    } else if (e < Lay::pB3) {
        const int c = (e - Lay::pW3) / kMidL3P, i = (e - Lay::pW3) % kMidL3P;
        if (c < NC && i < L3) g_w3[(size_t)c * L3 + i] = v;
    } else if (e < Lay::pB1) {
        if (e - Lay::pB3 < NC) g_b3[e - Lay::pB3] = v;
    } else if (e < Lay::pLoss) {
        if (e - Lay::pB1 < L2) g_b1[e - Lay::pB1] = v;
    } else if (loss) {
        loss[0] = v * inv_count;
    }
}

// act1 [B][L2] -> loss, g_z1 [B][L2], g_w2, g_b2, g_w3, g_b3 and g_b1 (= column sums of g_z1);
// partial: head_mid_grid(s) * kHeadMidPartial floats
int launch_head_mid(const nnue_shape &s, const float *act1, const int64_t *labels, float inv_count, const float *w2, const float *b2,
                    const float *w3, const float *b3, float *loss, float *g_z1, float *g_w2, float *g_b2, float *g_w3, float *g_b3,
                    float *g_b1, float *partial, cudaStream_t st) {
    if (!head_mid_ok(s)) return NNUE_ERR_UNSUPPORTED;
    HeadMidArgs a{};
    a.B = s.B; a.L2 = s.L2; a.L3 = s.L3; a.NC = s.NC;
    a.act1 = act1; a.labels = labels; a.w2 = w2; a.b2 = b2; a.w3 = w3; a.b3 = b3;
    a.inv_count = inv_count; a.g_z1 = g_z1; a.partial = partial;
    const int grid = head_mid_grid(s);
    constexpr size_t smem = (size_t)MidLayout::total * sizeof(float);
    NNUE_CUDA_TRY(cudaFuncSetAttribute(head_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_mid_kernel<<<grid, kMidThreads, smem, st>>>(a);
    NNUE_CHECK_LAUNCH("head_mid_kernel");
    head_mid_fold_kernel<<<ceil_div(MidLayout::pLoss + 1, 256), 256, 0, st>>>(grid, partial, s.L2, s.L3, s.NC, inv_count, g_w2, g_b2,
                                                                             g_w3, g_b3, g_b1, loss);
    NNUE_CHECK_LAUNCH("head_mid_fold_kernel");
    return NNUE_OK;
}

}  // namespace nnue
