// optim.cu -- the step either side of the hot path: gradient-norm clipping and
// the SGD-momentum / Adam update of train.py:363-366, 457-471 over ONE flat fp32 buffer (parameters,
// gradients and optimizer state are views of flat buffers, see train.py FlatParamBuffer), instead of one
// small kernel per parameter tensor.  Semantics are torch's: clip_grad_norm_(norm 2, eps 1e-6),
// torch.optim.SGD (dampening 0, no nesterov, L2 weight decay), torch.optim.Adam (L2 weight decay,
// no amsgrad).  HBM-bound elementwise work: 16 B (SGD) / 20 B (Adam) moved per parameter.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

constexpr int kOptThreads = 256;

// partial[blk] = sum of g[i]^2 over the block's grid-stride slice (fixed order inside the block)
__global__ void __launch_bounds__(kOptThreads)
grad_sqnorm_partial_kernel(long long n, const float *__restrict__ g, float *__restrict__ partial) {
    __shared__ float red[kOptThreads / 32];
    float acc = 0.0f;
    for (long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
        const float v = g[i];
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.0f;
        for (int k = 0; k < kOptThreads / 32; ++k) v += red[k];
        partial[blockIdx.x] = v;
    }
}
__global__ void grad_sqnorm_fold_kernel(int nblk, const float *__restrict__ partial, float *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float v = 0.0f;
        for (int k = 0; k < nblk; ++k) v += partial[k];
        out[0] = v;
    }
}

// clip coefficient of torch.nn.utils.clip_grad_norm_: min(1, max_norm / (total_norm + 1e-6))
__device__ __forceinline__ float clip_coef(float max_norm, const float *sqnorm) {
    if (!(max_norm > 0.0f) || !sqnorm) return 1.0f;
    const float c = max_norm / (sqrtf(__ldg(sqnorm)) + 1e-6f);
    return c < 1.0f ? c : 1.0f;
}

__global__ void __launch_bounds__(kOptThreads)
sgd_step_kernel(long long n, float *__restrict__ p, float *__restrict__ g, float *__restrict__ buf, float lr,
                float momentum, float weight_decay, float max_norm, const float *__restrict__ sqnorm, int first) {
    const float coef = clip_coef(max_norm, sqnorm);
    for (long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
        const float w = p[i];
        float d = g[i] * coef;
        g[i] = d;  // clip_grad_norm_ scales the gradients in place
        d = fmaf(weight_decay, w, d);
        if (momentum != 0.0f) {
            const float b = first ? d : fmaf(momentum, buf[i], d);
            buf[i] = b;
            d = b;
        }
        p[i] = fmaf(-lr, d, w);
    }
}

__global__ void __launch_bounds__(kOptThreads)
adam_step_kernel(long long n, float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 float lr, float beta1, float beta2, float eps, float weight_decay, float bias1, float bias2_sqrt,
                 float max_norm, const float *__restrict__ sqnorm) {
    const float coef = clip_coef(max_norm, sqnorm);
    const float step_size = lr / bias1;
    for (long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
        const float w = p[i];
        float d = g[i] * coef;
        g[i] = d;
        d = fmaf(weight_decay, w, d);
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * d);        // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * d * d);    // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bias2_sqrt + eps;
        p[i] = w - step_size * (mi / denom);
    }
}

// dst[i] = src[i] * *scale: the upstream gradient of the scalar loss applied to every parameter gradient in one pass
// (train.py:352-361: loss.backward() with an arbitrary upstream factor, e.g. (loss * w).backward())
__global__ void __launch_bounds__(kOptThreads)
scale_flat_kernel(long long n, const float *__restrict__ src, const float *__restrict__ scale, float *__restrict__ dst) {
    const float g = __ldg(scale);
    for (long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) dst[i] = src[i] * g;
}

static int opt_grid(long long n) {
    long long g = (n + kOptThreads - 1) / kOptThreads;
    if (g > 8LL * kNumSMs) g = 8LL * kNumSMs;
    return g < 1 ? 1 : (int)g;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

size_t nnue_opt_workspace_bytes(long long n) { return (size_t)opt_grid(n) * 4 + 256; }

int nnue_opt_grad_sqnorm(long long n, const float *g_d, float *sqnorm_d, void *workspace_d, size_t workspace_bytes,
                         void *stream) {
    if (n < 1 || !g_d || !sqnorm_d || !workspace_d) return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < nnue_opt_workspace_bytes(n)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = opt_grid(n);
    float *partial = static_cast<float *>(workspace_d);
    grad_sqnorm_partial_kernel<<<grid, kOptThreads, 0, st>>>(n, g_d, partial);
    NNUE_CHECK_LAUNCH("grad_sqnorm_partial_kernel");
    grad_sqnorm_fold_kernel<<<1, 32, 0, st>>>(grid, partial, sqnorm_d);
    NNUE_CHECK_LAUNCH("grad_sqnorm_fold_kernel");
    return NNUE_OK;
}

int nnue_scale_flat(long long n, const float *src_d, const float *scale_d, float *dst_d, void *stream) {
    if (n < 1 || !src_d || !scale_d || !dst_d) return NNUE_ERR_INVALID_ARG;
    scale_flat_kernel<<<opt_grid(n), kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(n, src_d, scale_d, dst_d);
    NNUE_CHECK_LAUNCH("scale_flat_kernel");
    return NNUE_OK;
}

int nnue_opt_sgd_step(long long n, float *p_d, float *g_d, float *buf_d, float lr, float momentum, float weight_decay,
                      float max_norm, const float *sqnorm_d, int first_step, void *stream) {
    if (n < 1 || !p_d || !g_d || (momentum != 0.0f && !buf_d) || (max_norm > 0.0f && !sqnorm_d)) return NNUE_ERR_INVALID_ARG;
    sgd_step_kernel<<<opt_grid(n), kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        n, p_d, g_d, buf_d, lr, momentum, weight_decay, max_norm, sqnorm_d, first_step);
    NNUE_CHECK_LAUNCH("sgd_step_kernel");
    return NNUE_OK;
}

int nnue_opt_adam_step(long long n, float *p_d, float *g_d, float *m_d, float *v_d, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, float max_norm, const float *sqnorm_d, void *stream) {
    if (n < 1 || !p_d || !g_d || !m_d || !v_d || step < 1 || (max_norm > 0.0f && !sqnorm_d)) return NNUE_ERR_INVALID_ARG;
    const float bias1 = 1.0f - powf(beta1, (float)step);
    const float bias2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    adam_step_kernel<<<opt_grid(n), kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        n, p_d, g_d, m_d, v_d, lr, beta1, beta2, eps, weight_decay, bias1, bias2_sqrt, max_norm, sqnorm_d);
    NNUE_CHECK_LAUNCH("adam_step_kernel");
    return NNUE_OK;
}

}  // extern "C"
