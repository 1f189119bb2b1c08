// qinfer.cuh -- launch parameters shared by the integer-inference kernels (qinfer.cu, qinfer_fast.cu).
#pragma once
#include "common.cuh"

namespace nnue {

struct QParams {
    int B, H, W, stride, oh, ow;
    int F, L1, L2, L3, NC, OC, L1p, K1, K2, K3;
    float threshold, conv_scale;
    int conv_iscale, qone, l2_iscale;
    float l1_scale, out_scale;
    const int32_t *conv_w, *conv_b;
    const int16_t *ft_w, *ft_b;
    const int32_t *w1, *b1, *w2, *b2, *wo, *bo;
    const float *images;
    float *logits, *density;
    // split form (MODE 1 / 2 of the kernel)
    int G2, CWq;              // cells of the whole feature buffer (F / OC), bitmask words per channel
    uint32_t *bits_out;       // MODE 1: [B][OC][CWq]
    const int16_t *acc_in;    // MODE 2: [B][L1]
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(lo, min(hi, v)); }

// ---- qinfer_fast.cu ---------------------------------------------------------------------------------------------
// small batches: one CTA per sample (conv words over the warps, rows of the active set over the warps, int16x2
// partial accumulators folded through shared memory, dense layers over all threads)
int launch_q_infer_cta(const QParams &q, cudaStream_t st);
// 32 x 32 images at conv stride 4 (the 8 x 8 raster of config D's grid): conv + threshold -> bitmask [B][OC][CWq] and
// density, image rows staged by bulk TMA, conv taps as constant-bank operands (taps_d [OC][28], bias in entry 27)
bool q_conv_bits32_ok(const QParams &q);
int launch_q_conv_bits32(const QParams &q, const int32_t *taps_d, cudaStream_t st);

}  // namespace nnue
