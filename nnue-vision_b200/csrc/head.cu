// head.cu -- pairwise product + the three dense layers (forward and backward) and the fused
// mean cross-entropy.  The layers are small (L1 <= 1024 inputs, <= 1000 outputs) next to the batch,
// so they run as fp32 register-tiled FMA kernels (exact fp32 accumulation keeps the 1e-5 parity
// bar; a bf16/tf32 tensor-core path would not) with bias, ReLU, ReLU-mask and the pairwise
// transform fused into the tile loaders and epilogues.
#include "common.cuh"
#include "plan.cuh"

namespace nnue {

// C[m, n] = epilogue( sum_k A(m,k) * B(k,n) ), operands addressed through element strides so
// that one kernel serves X*W^T (forward), dY*W (input gradient) and dY^T*X (weight gradient).
struct GemmArgs {
    const float *A; long long sam, sak;
    const float *Bm; long long sbk, sbn;
    float *C; long long scm;            // C[m*scm + n]
    int M, N, K;
    const float *bias;                  // [N] added before the activation, or null
    const float *mask; long long smm;   // C *= (mask[m*smm + n] > 0): ReLU backward, or null
    int relu;                           // max(0, .) in the epilogue
    int a_pair_half;                    // > 0: A is ft[m, 2h]; A(m,k) = k < h ? ft[m,k]*ft[m,k+h] : ft[m,k-h]
    int b_pair_half;                    // > 0: same transform on B(k, n) = l0[k, n] rows of ft
    int b_ones_col;                     // >= 0: B(k, n == b_ones_col) = 1 (bias-gradient column)
    int ksplit;                         // K elements per blockIdx.z slice (split-K partials)
    long long c_split_stride;           // elements between consecutive split-K partials of C
};

__device__ __forceinline__ float gemm_load_a(const GemmArgs &g, int m, int k) {
    if (g.a_pair_half > 0) {
        const int h = g.a_pair_half;
        const float *row = g.A + (size_t)m * g.sam;
        return k < h ? __ldg(row + k) * __ldg(row + k + h) : __ldg(row + k - h);
    }
    return __ldg(g.A + (size_t)m * g.sam + (size_t)k * g.sak);
}
__device__ __forceinline__ float gemm_load_b(const GemmArgs &g, int k, int n) {
    if (n == g.b_ones_col) return 1.0f;
    if (g.b_pair_half > 0) {
        const int h = g.b_pair_half;
        const float *row = g.Bm + (size_t)k * g.sbk;
        return n < h ? __ldg(row + n) * __ldg(row + n + h) : __ldg(row + n - h);
    }
    return __ldg(g.Bm + (size_t)k * g.sbk + (size_t)n * g.sbn);
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(const GemmArgs g) {
    constexpr int BK = kGemmBK;
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_begin = blockIdx.z * g.ksplit;
    const int k_end = min(g.K, k_begin + g.ksplit);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    const bool a_k_contig = (g.sak == 1) || g.a_pair_half > 0;
    const bool b_k_contig = (g.sbk == 1) && g.b_pair_half == 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        for (int i = tid; i < BM * BK; i += NT) {
            int m, k;
            if (a_k_contig) { m = i / BK; k = i % BK; } else { k = i / BM; m = i % BM; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < g.M && gk < k_end) ? gemm_load_a(g, gm, gk) : 0.0f;
        }
        for (int i = tid; i < BN * BK; i += NT) {
            int n, k;
            if (b_k_contig) { n = i / BK; k = i % BK; } else { k = i / BN; n = i % BN; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < g.N && gk < k_end) ? gemm_load_b(g, gk, gn) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = g.C + (size_t)blockIdx.z * g.c_split_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + n);
            if (g.relu) v = fmaxf(v, 0.0f);
            if (g.mask) v = __ldg(g.mask + (size_t)m * g.smm + n) > 0.0f ? v : 0.0f;
            C[(size_t)m * g.scm + n] = v;
        }
    }
}

static int launch_gemm(GemmArgs g, int splits, cudaStream_t st) {
    g.ksplit = ceil_div(ceil_div(g.K, splits), kGemmBK) * kGemmBK;
    const int nz = ceil_div(g.K, g.ksplit);
    if (g.N <= 8) {
        gemm_kernel<128, 8, 4, 1><<<dim3(ceil_div(g.M, 128), 1, nz), 256, 0, st>>>(g);
    } else if (g.N <= 16) {
        gemm_kernel<128, 16, 4, 2><<<dim3(ceil_div(g.M, 128), 1, nz), 256, 0, st>>>(g);
    } else if (g.N <= 32) {
        // few row tiles (small batches): 32-row tiles put four times as many CTAs on the machine
        if (ceil_div(g.M, 128) * nz < kNumSMs) gemm_kernel<32, 32, 2, 2><<<dim3(ceil_div(g.M, 32), 1, nz), 256, 0, st>>>(g);
        else gemm_kernel<128, 32, 8, 2><<<dim3(ceil_div(g.M, 128), 1, nz), 256, 0, st>>>(g);
    } else {
        gemm_kernel<64, 64, 4, 4><<<dim3(ceil_div(g.M, 64), ceil_div(g.N, 64), nz), 256, 0, st>>>(g);
    }
    NNUE_CHECK_LAUNCH("gemm_kernel");
    return nz;
}

// Weight(+bias) gradient: dWb[N_out, K_in + 1] = dY^T [N_out, B] * [X | 1] [B, K_in + 1], split over
// the batch into partials, then folded in slice order into g_w [N_out, K_in] and g_b [N_out].
__global__ void fold_wgrad_kernel(int n_out, int k_in, int nparts, const float *__restrict__ partial,
                                  float *__restrict__ g_w, float *__restrict__ g_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cols = k_in + 1;
    if (i >= n_out * cols) return;
    float acc = 0.0f;
    for (int k = 0; k < nparts; ++k) acc += partial[(size_t)k * n_out * cols + i];
    const int r = i / cols, c = i % cols;
    if (c < k_in) g_w[(size_t)r * k_in + c] = acc;
    else g_b[r] = acc;
}

static int wgrad(const float *dY, int n_out, const float *X, int k_in, int pair_half, int B, float *partial,
                 float *g_w, float *g_b, cudaStream_t st) {
    GemmArgs g{};
    g.A = dY; g.sam = 1; g.sak = n_out;       // A(m = out, k = b) = dY[b, out]
    g.Bm = X; g.sbk = pair_half > 0 ? 2 * pair_half : k_in; g.sbn = 1;  // B(k = b, n) = X[b, n]
    g.b_pair_half = pair_half;
    g.b_ones_col = k_in;
    g.M = n_out; g.N = k_in + 1; g.K = B;
    g.C = partial; g.scm = k_in + 1;
    g.c_split_stride = (long long)n_out * (k_in + 1);
    const int splits = gemm_splits(n_out, k_in + 1, B, 64, 64);
    // always the 64x64 tile here: M and N are layer widths, K is the batch
    g.ksplit = ceil_div(ceil_div(g.K, splits), kGemmBK) * kGemmBK;
    const int nz = ceil_div(g.K, g.ksplit);
    gemm_kernel<64, 64, 4, 4><<<dim3(ceil_div(g.M, 64), ceil_div(g.N, 64), nz), 256, 0, st>>>(g);
    NNUE_CHECK_LAUNCH("gemm_kernel(wgrad)");
    const int n = n_out * (k_in + 1);
    fold_wgrad_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n_out, k_in, nz, partial, g_w, g_b);
    NNUE_CHECK_LAUNCH("fold_wgrad_kernel");
    return NNUE_OK;
}

// pairwise backward: g_ft[:, i] = g_l0[:, i] * ft[:, i+h] + g_l0[:, i+h];  g_ft[:, i+h] = g_l0[:, i] * ft[:, i]
__global__ void pairwise_bwd_kernel(long long n, int h, const float *__restrict__ g_l0, const float *__restrict__ ft,
                                    float *__restrict__ g_ft) {
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / h;
    const int c = (int)(i % h);
    const size_t o = (size_t)b * 2 * h + c;
    const float gp = g_l0[o], ga = g_l0[o + h];
    g_ft[o] = fmaf(gp, ft[o + h], ga);
    g_ft[o + h] = gp * ft[o];
}

// mean cross-entropy + gradient; one warp per sample
__global__ void __launch_bounds__(256)
ce_kernel(int B, int NC, const float *__restrict__ logits, const int64_t *__restrict__ labels, float inv_count,
          const float *__restrict__ g_scale, float *__restrict__ per_sample, float *__restrict__ g_logits) {
    const int lane = threadIdx.x & 31;
    const int b = (int)((1LL * blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (b >= B) return;
    const float *row = logits + (size_t)b * NC;
    float mx = -INFINITY;
    for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float se = 0.0f;
    for (int c = lane; c < NC; c += 32) se += expf(row[c] - mx);
    se = warp_sum(se);
    const int y = min(max((int)labels[b], 0), NC - 1);  // (a target outside [0, NC) must not read out of bounds)
    const float lse = mx + logf(se);
    if (per_sample && lane == 0) per_sample[b] = lse - row[y];
    if (g_logits) {
        const float g_mul = g_scale ? inv_count * __ldg(g_scale) : inv_count;
        const float inv = 1.0f / se;
        for (int c = lane; c < NC; c += 32) {
            const float p = expf(row[c] - mx) * inv;
            g_logits[(size_t)b * NC + c] = (p - (c == y ? 1.0f : 0.0f)) * g_mul;
        }
    }
}
// fixed-order tree sum of n values by one CTA
__global__ void __launch_bounds__(1024)
sum_scale_kernel(int n, const float *__restrict__ x, float scale, float *__restrict__ out) {
    __shared__ float red[32];
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * scale;
    }
}

// column sums of x [rows][cols] in two fixed-order stages (bias gradient next to the tensor-core weight gradient)
__global__ void head_colsum_partial_kernel(int rows, int cols, const float *__restrict__ x, float *__restrict__ partial) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int r0 = blockIdx.y * 256, r1 = min(rows, r0 + 256);
    float acc = 0.0f;
    int r = r0;
    for (; r + 16 <= r1; r += 16) {  // loads batched sixteen deep, same summation order
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = __ldg(x + (size_t)(r + u) * cols + c);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += t[u];
    }
    for (; r < r1; ++r) acc += __ldg(x + (size_t)r * cols + c);
    partial[(size_t)blockIdx.y * cols + c] = acc;
}
__global__ void head_fold_kernel(long long n, int nparts, long long stride, const float *__restrict__ partial, float *__restrict__ out) {
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.0f;
    int k = 0;
    for (; k + 8 <= nparts; k += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldg(partial + (size_t)(k + u) * stride + i);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += t[u];
    }
    for (; k < nparts; ++k) acc += __ldg(partial + (size_t)k * stride + i);
    out[i] = acc;
}

// layer 1 alone: act1 = relu(l0 W1^T + b1), l0 = the pairwise transform of ft_out; on the tensor cores when the shape is wide
// and the scratch (ws_head_umma_fwd bytes) is there
int head_layer1_fwd(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d, float *act1_d,
                    void *workspace_d, size_t workspace_bytes, cudaStream_t st) {
    int rc;
    if (head_umma_ok(*s) && workspace_d && workspace_bytes >= ws_head_umma_fwd(*s)) {
        // A = l0 (pairwise transform fused into the formatter), B = W1, both K-major
        const int NT = head_umma_nt(*s);
        unsigned char *at = static_cast<unsigned char *>(workspace_d);
        unsigned char *bt = at + align_up(ugemm_tile_bytes(s->B, 128, s->L1), 256);
        if ((rc = ugemm_format_rows(NT, w1_d, s->L1, s->L2, s->L1, 0, bt, st)) < 0) return rc;
        if (gemm_inline_a_ok(s->L1, s->L1 / 2)) {
            // l0 never exists in HBM: the GEMM's epilogue warps build its split-bf16 tiles from ft_out while the main loop runs
            // (the formatter wrote 100 MB and the GEMM read them back: 34 + 37 us at L1 = 1024, batch 16384)
            rc = ugemm_launch(NT, s->B, s->L2, s->L1, nullptr, bt, act1_d, s->L2, b1_d, 1, nullptr, 0, 1, 0, st, nullptr, 0, ft_out_d,
                              s->L1, s->L1 / 2);
            return rc < 0 ? rc : NNUE_OK;
        }
        if ((rc = ugemm_format_rows(128, ft_out_d, s->L1, s->B, s->L1, s->L1 / 2, at, st)) < 0) return rc;
        rc = ugemm_launch(NT, s->B, s->L2, s->L1, at, bt, act1_d, s->L2, b1_d, 1, nullptr, 0, 1, 0, st);
        return rc < 0 ? rc : NNUE_OK;
    }
    GemmArgs g{};
    g.b_ones_col = -1;
    g.A = ft_out_d; g.sam = s->L1; g.sak = 1; g.a_pair_half = s->L1 / 2;  // l0 built on the fly from ft_out
    g.Bm = w1_d; g.sbk = 1; g.sbn = s->L1;
    g.M = s->B; g.N = s->L2; g.K = s->L1;
    g.C = act1_d; g.scm = s->L2; g.bias = b1_d; g.relu = 1;
    rc = launch_gemm(g, 1, st);
    return rc < 0 ? rc : NNUE_OK;
}

// forward with optional scratch: layer 1 on the tensor cores when the shape is wide and the scratch is there
int head_fwd_ws(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d, const float *w2_d,
                const float *b2_d, const float *w3_d, const float *b3_d, float *act1_d, float *act2_d, float *logits_d,
                void *workspace_d, size_t workspace_bytes, cudaStream_t st) {
    GemmArgs g{};
    g.b_ones_col = -1;
    int rc = head_layer1_fwd(s, ft_out_d, w1_d, b1_d, act1_d, workspace_d, workspace_bytes, st);
    if (rc != NNUE_OK) return rc;
    // layer 2
    g.A = act1_d; g.sam = s->L2; g.sak = 1; g.a_pair_half = 0;
    g.Bm = w2_d; g.sbk = 1; g.sbn = s->L2;
    g.M = s->B; g.N = s->L3; g.K = s->L2;
    g.C = act2_d; g.scm = s->L3; g.bias = b2_d; g.relu = 1;
    rc = launch_gemm(g, 1, st);
    if (rc < 0) return rc;
    // output layer
    g.A = act2_d; g.sam = s->L3;
    g.Bm = w3_d; g.sbn = s->L3;
    g.N = s->NC; g.K = s->L3;
    g.C = logits_d; g.scm = s->NC; g.bias = b3_d; g.relu = 0;
    rc = launch_gemm(g, 1, st);
    return rc < 0 ? rc : NNUE_OK;
}

// scratch of nnue_head_bwd: split-K partials of the three weight(+bias) gradients, g_act2, g_act1 (= g_z1), g_l0, then
// the tensor-core scratch of layer 1 (ws_head_umma_bwd); the many-class scratch sits at the END (ws_head3_umma_bwd)
HeadBwdWs carve_head_bwd(const nnue_shape &s, void *workspace_d) {
    char *ws = static_cast<char *>(workspace_d);
    auto carve = [&](size_t bytes) { float *p = reinterpret_cast<float *>(ws); ws += align_up(bytes, 256); return p; };
    HeadBwdWs w{};
    w.p3 = carve((size_t)gemm_splits(s.NC, s.L3 + 1, s.B, 64, 64) * s.NC * (s.L3 + 1) * 4);
    w.p2 = carve((size_t)gemm_splits(s.L3, s.L2 + 1, s.B, 64, 64) * s.L3 * (s.L2 + 1) * 4);
    w.p1 = carve((size_t)gemm_splits(s.L2, s.L1 + 1, s.B, 64, 64) * s.L2 * (s.L1 + 1) * 4);
    w.g_act2 = carve((size_t)s.B * s.L3 * 4);
    w.g_act1 = carve((size_t)s.B * s.L2 * 4);
    w.g_l0 = carve((size_t)s.B * s.L1 * 4);
    w.rest = ws;
    return w;
}

// layer 1 backward from g_z1 (bw.g_act1, already masked by act1 > 0): g_w1, g_b1, g_l0 and the pairwise backward -> g_ft.
// `side` (tensor-core form only): the weight-gradient chain (two operand formatters, split-K GEMM, fold -- independent of
// the input-gradient chain and of everything downstream) runs on side->stream out of side->ws, behind an event recorded on
// `st`; bw.g_act1 must then live where the later stages of `st` do not write (the caller carves it from the side scratch).
int head_bwd_layer1(const nnue_shape *s, const HeadBwdWs &bw, const float *ft_out_d, const float *w1_d, float *g_w1_d,
                    float *g_b1_d, float *g_ft_d, cudaStream_t st, const HeadSide *side) {
    const int B = s->B, L1 = s->L1, L2 = s->L2;
    float *g_act1 = bw.g_act1, *g_l0 = bw.g_l0, *p1 = bw.p1;
    char *ws = bw.rest;
    int rc;
    GemmArgs g{};
    g.b_ones_col = -1;
    g.sak = 1; g.sbn = 1; g.M = B;
    if (head_umma_ok(*s)) {
        // both layer-1 gradients as split-bf16 tcgen05 GEMMs (scratch carved behind g_l0)
        auto carve_b = [&](size_t bytes) { unsigned char *p = reinterpret_cast<unsigned char *>(ws); ws += align_up(bytes, 256); return p; };
        unsigned char *g1_rows = carve_b(ugemm_tile_bytes(B, 128, L2)), *w1_cols = carve_b(ugemm_tile_bytes(L1, 256, L2));
        unsigned char *g1_cols = carve_b(ugemm_tile_bytes(L2, 128, B)), *l0_cols = carve_b(ugemm_tile_bytes(L1, 256, B));
        const int splits = head_umma_wgrad_splits(*s);
        float *wpart = reinterpret_cast<float *>(carve_b((size_t)splits * L2 * L1 * 4));
        float *cpart = reinterpret_cast<float *>(carve_b((size_t)ceil_div(B, 256) * L2 * 4));
        // g_w1[o, i] = sum_b g_z1[b, o] l0[b, i]: A = g_z1^T, B = l0^T (pairwise fused), K = batch, split-K partials
        cudaStream_t wst = st;
        if (side) {  // the chain's scratch comes from the side workspace, its launches go to the side stream
            unsigned char *sp = static_cast<unsigned char *>(side->ws);
            auto carve_s = [&](size_t bytes) { unsigned char *p = sp; sp += align_up(bytes, 256); return p; };
            g1_cols = carve_s(ugemm_tile_bytes(L2, 128, B)); l0_cols = carve_s(ugemm_tile_bytes(L1, 256, B));
            wpart = reinterpret_cast<float *>(carve_s((size_t)splits * L2 * L1 * 4));
            cpart = reinterpret_cast<float *>(carve_s((size_t)ceil_div(B, 256) * L2 * 4));
            NNUE_CUDA_TRY(cudaEventRecord(side->ready, st));
            NNUE_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->ready, 0));
            wst = side->stream;
        }
        if ((rc = ugemm_format_cols(128, g_act1, L2, B, L2, 0, g1_cols, wst)) < 0) return rc;
        if ((rc = ugemm_format_cols(256, ft_out_d, L1, B, L1, L1 / 2, l0_cols, wst)) < 0) return rc;
        const int nz = ugemm_launch(256, L2, L1, B, g1_cols, l0_cols, wpart, L1, nullptr, 0, nullptr, 0, splits, (long long)L2 * L1, wst);
        if (nz < 0) return nz;
        const long long nw = 1LL * L2 * L1;
        head_fold_kernel<<<(int)((nw + 255) / 256), 256, 0, wst>>>(nw, nz, nw, wpart, g_w1_d);
        NNUE_CHECK_LAUNCH("head_fold_kernel");
        if (g_b1_d) {
            const int nrc = ceil_div(B, 256);
            head_colsum_partial_kernel<<<dim3(ceil_div(L2, 128), nrc), 128, 0, wst>>>(B, L2, g_act1, cpart);
            NNUE_CHECK_LAUNCH("head_colsum_partial_kernel");
            head_fold_kernel<<<ceil_div(L2, 256), 256, 0, wst>>>(L2, nrc, L2, cpart, g_b1_d);
            NNUE_CHECK_LAUNCH("head_fold_kernel");
        }
        // g_l0 = g_z1 W1: A = g_z1, B = W1^T
        if ((rc = ugemm_format_rows(128, g_act1, L2, B, L2, 0, g1_rows, st)) < 0) return rc;
        if (head_pair_epilogue_ok(*s)) {
            // W1's columns permuted so that an N tile holds g_l0[:, i] and g_l0[:, h + i]: the epilogue writes g_ft itself
            // (no g_l0 round trip, no pairwise kernel: 37 us of the 0.96 ms step at L1 = 1024, batch 16384)
            if ((rc = ugemm_format_cols(256, w1_d, L1, L2, L1, 0, w1_cols, st, L1 / 2)) < 0) return rc;
            rc = ugemm_launch(256, B, L1, L2, g1_rows, w1_cols, g_ft_d, L1, nullptr, 0, nullptr, 0, 1, 0, st, ft_out_d, L1 / 2);
            return rc < 0 ? rc : NNUE_OK;
        }
        if ((rc = ugemm_format_cols(256, w1_d, L1, L2, L1, 0, w1_cols, st)) < 0) return rc;
        rc = ugemm_launch(256, B, L1, L2, g1_rows, w1_cols, g_l0, L1, nullptr, 0, nullptr, 0, 1, 0, st);
        if (rc < 0) return rc;
    } else {
        rc = wgrad(g_act1, L2, ft_out_d, L1, L1 / 2, B, p1, g_w1_d, g_b1_d, st);
        if (rc < 0) return rc;
        // g_l0 = g_z1 * W1, then the pairwise backward
        g.A = g_act1; g.sam = L2;
        g.Bm = w1_d; g.sbk = L1;
        g.N = L1; g.K = L2;
        g.C = g_l0; g.scm = L1; g.mask = nullptr;
        rc = launch_gemm(g, 1, st);
        if (rc < 0) return rc;
    }
    const long long n = 1LL * B * (L1 / 2);
    pairwise_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(n, L1 / 2, g_l0, ft_out_d, g_ft_d);
    NNUE_CHECK_LAUNCH("pairwise_bwd_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_head_fwd(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d, const float *w2_d,
                  const float *b2_d, const float *w3_d, const float *b3_d, float *act1_d, float *act2_d,
                  float *logits_d, void *stream) {
    if (!s || !ft_out_d || !w1_d || !b1_d || !w2_d || !b2_d || !w3_d || !b3_d || !act1_d || !act2_d || !logits_d)
        return NNUE_ERR_INVALID_ARG;
    return head_fwd_ws(s, ft_out_d, w1_d, b1_d, w2_d, b2_d, w3_d, b3_d, act1_d, act2_d, logits_d, nullptr, 0,
                       static_cast<cudaStream_t>(stream));
}

int nnue_ce_fwd_bwd(int B, int NC, const float *logits_d, const int64_t *labels_d, float inv_count,
                    const float *g_scale_d, float *loss_d, float *per_sample_d, float *g_logits_d, void *workspace_d,
                    size_t workspace_bytes, void *stream) {
    if (B < 1 || NC < 1 || !logits_d || !labels_d || (!loss_d && !g_logits_d)) return NNUE_ERR_INVALID_ARG;
    float *per = per_sample_d;
    if (!per && loss_d) {
        if (!workspace_d || workspace_bytes < ws_ce(B)) return NNUE_ERR_WORKSPACE;
        per = static_cast<float *>(workspace_d);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ce_kernel<<<ceil_div(B, 8), 256, 0, st>>>(B, NC, logits_d, labels_d, inv_count, g_scale_d, per, g_logits_d);
    NNUE_CHECK_LAUNCH("ce_kernel");
    if (loss_d) {
        sum_scale_kernel<<<1, 1024, 0, st>>>(B, per, inv_count, loss_d);
        NNUE_CHECK_LAUNCH("sum_scale_kernel");
    }
    return NNUE_OK;
}

int nnue_head_bwd(const nnue_shape *s, const float *g_logits_d, const float *ft_out_d, const float *act1_d,
                  const float *act2_d, const float *w1_d, const float *w2_d, const float *w3_d, float *g_w1_d,
                  float *g_b1_d, float *g_w2_d, float *g_b2_d, float *g_w3_d, float *g_b3_d, float *g_ft_d,
                  void *workspace_d, size_t workspace_bytes, void *stream) {
    if (!s || !g_logits_d || !ft_out_d || !act1_d || !act2_d || !w1_d || !w2_d || !w3_d || !g_w1_d || !g_b1_d ||
        !g_w2_d || !g_b2_d || !g_w3_d || !g_b3_d || !g_ft_d || !workspace_d)
        return NNUE_ERR_INVALID_ARG;
    if (workspace_bytes < ws_head_bwd(*s)) return NNUE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = s->B, L2 = s->L2, L3 = s->L3, NC = s->NC;
    const HeadBwdWs bw = carve_head_bwd(*s, workspace_d);
    float *p3 = bw.p3, *p2 = bw.p2, *g_act2 = bw.g_act2, *g_act1 = bw.g_act1;
    int rc;
    GemmArgs g{};
    g.b_ones_col = -1;
    g.sak = 1; g.sbn = 1; g.M = B;  // (common to the layer gradients below)
    if (head3_umma_ok(*s)) {
        // many classes: both output-layer gradients as split-bf16 tcgen05 GEMMs (scratch at the END of the workspace, behind
        // everything the other layers carve)
        unsigned char *w3s = reinterpret_cast<unsigned char *>(static_cast<char *>(workspace_d) + ws_head_bwd(*s) - ws_head3_umma_bwd(*s));
        auto carve3 = [&](size_t bytes) { unsigned char *p = w3s; w3s += align_up(bytes, 256); return p; };
        unsigned char *gl_rows = carve3(ugemm_tile_bytes(B, 128, NC)), *w3_cols = carve3(ugemm_tile_bytes(L3, 128, NC));
        unsigned char *gl_cols = carve3(ugemm_tile_bytes(NC, 128, B)), *a2_cols = carve3(ugemm_tile_bytes(L3, 128, B));
        const int splits = head3_umma_wgrad_splits(*s);
        float *wpart = reinterpret_cast<float *>(carve3((size_t)splits * NC * L3 * 4));
        float *cpart = reinterpret_cast<float *>(carve3((size_t)ceil_div(B, 256) * NC * 4));
        // g_w3[o, i] = sum_b g_logits[b, o] act2[b, i]: A = g_logits^T, B = act2^T, K = batch, split-K partials
        if ((rc = ugemm_format_cols(128, g_logits_d, NC, B, NC, 0, gl_cols, st)) < 0) return rc;
        if ((rc = ugemm_format_cols(128, act2_d, L3, B, L3, 0, a2_cols, st)) < 0) return rc;
        const int nz = ugemm_launch(128, NC, L3, B, gl_cols, a2_cols, wpart, L3, nullptr, 0, nullptr, 0, splits, (long long)NC * L3, st);
        if (nz < 0) return nz;
        const long long nw = 1LL * NC * L3;
        head_fold_kernel<<<(int)((nw + 255) / 256), 256, 0, st>>>(nw, nz, nw, wpart, g_w3_d);
        NNUE_CHECK_LAUNCH("head_fold_kernel");
        const int nrc = ceil_div(B, 256);
        head_colsum_partial_kernel<<<dim3(ceil_div(NC, 128), nrc), 128, 0, st>>>(B, NC, g_logits_d, cpart);
        NNUE_CHECK_LAUNCH("head_colsum_partial_kernel");
        head_fold_kernel<<<ceil_div(NC, 256), 256, 0, st>>>(NC, nrc, NC, cpart, g_b3_d);
        NNUE_CHECK_LAUNCH("head_fold_kernel");
        // g_z2 = (g_logits W3) masked by act2 > 0: A = g_logits, B = W3^T (rows = the L3 inputs, K = classes)
        if ((rc = ugemm_format_rows(128, g_logits_d, NC, B, NC, 0, gl_rows, st)) < 0) return rc;
        if ((rc = ugemm_format_cols(128, w3_d, L3, NC, L3, 0, w3_cols, st)) < 0) return rc;
        rc = ugemm_launch(128, B, L3, NC, gl_rows, w3_cols, g_act2, L3, nullptr, 0, act2_d, L3, 1, 0, st);
        if (rc < 0) return rc;
    } else {
        rc = wgrad(g_logits_d, NC, act2_d, L3, 0, B, p3, g_w3_d, g_b3_d, st);
        if (rc < 0) return rc;
        // g_z2 = (g_logits * W3) masked by act2 > 0
        g.A = g_logits_d; g.sam = NC; g.sak = 1;
        g.Bm = w3_d; g.sbk = L3; g.sbn = 1;
        g.M = B; g.N = L3; g.K = NC;
        g.C = g_act2; g.scm = L3; g.mask = act2_d; g.smm = L3;
        rc = launch_gemm(g, 1, st);
        if (rc < 0) return rc;
    }
    rc = wgrad(g_act2, L3, act1_d, L2, 0, B, p2, g_w2_d, g_b2_d, st);
    if (rc < 0) return rc;
    // g_z1 = (g_z2 * W2) masked by act1 > 0
    g.A = g_act2; g.sam = L3;
    g.Bm = w2_d; g.sbk = L2;
    g.N = L2; g.K = L3;
    g.C = g_act1; g.scm = L2; g.mask = act1_d; g.smm = L2;
    rc = launch_gemm(g, 1, st);
    if (rc < 0) return rc;
    return head_bwd_layer1(s, bw, ft_out_d, w1_d, g_w1_d, g_b1_d, g_ft_d, st, nullptr);
}

}  // extern "C"
