// gemm_umma.cu -- fp32-accurate dense GEMM on the 5th-generation tensor cores, for the layer-stack contractions
// that ARE dense at the benchmark batch (the 1024 -> 128 layer of the reference's "real" config: 13 GFLOP per
// training step, 2/3 of that step on the fp32 FMA kernels of head.cu).
//
//   C[m, n] = epilogue( sum_k A(m, k) * B(n, k) )
//
// Both operands are split exactly into three bf16 terms (umma.cuh) and the six term pairs (i, j), i + j <= 4, are
// issued per k-step, so products are exact and accumulation is fp32 in tensor memory: the same arithmetic contract
// as the feature-transformer value gradient (ft_umma.cu), held to the same 1e-5 parity bar.
//
// Operands are pre-formatted into UMMA tile order (canonical K-major, no swizzle; tiles [row tile][k-step][term] of
// [RT x 16] bf16) by two small kernels -- "rows" when k runs along the source rows' columns, "cols" when k runs
// down the source's rows (a transposed operand) -- with the pairwise transform of nnue.py:660-666 fused into the
// load where the operand is l0.  The main kernel is TMA-fed: warp 0 streams A and B tiles through a ring of
// stages, one thread of warp 1 issues the UMMAs (128 x NT x 16), warps 2-5 run the epilogue (bias, ReLU, ReLU
// mask, split-K partials).  Two CTAs per SM.
#include "common.cuh"
#include "plan.cuh"
#include "umma.cuh"

namespace nnue {

constexpr int kGuThreads = 192;
constexpr int kGuM = 128;

struct UGemmArgs {
    int M, N, n_ks;                      // n_ks = k-steps of 16 (K padded with zeros by the formatters)
    const unsigned char *at, *bt;        // A tiles [ceil(M/128)][n_ks][3][128 x 16], B tiles [ceil(N/NT)][n_ks][3][NT x 16]
    float *C; long long ldc;             // C[m * ldc + n]
    const float *bias;                   // [N] added before the activation, or null
    int relu;                            // max(0, .)
    const float *mask; long long ldm;    // C *= (mask[m * ldm + n] > 0), or null
    int ks_per_split;                    // k-steps per blockIdx.z slice
    long long c_split_stride;            // elements between consecutive split-K partials of C
    // pairwise-backward epilogue (NT = 256, B operand formatted with pair_perm = h): the tile's first 128 columns are
    // g_l0[:, i], the last 128 g_l0[:, h + i] for i = 128 nt ..; C is g_ft [M][2 h] and is written as
    //   g_ft[:, i] = g_l0[:, i] * ft[:, h + i] + g_l0[:, h + i],   g_ft[:, h + i] = g_l0[:, i] * ft[:, i]     (nnue.py:660-666 backward)
    const float *pair_ft;                // ft_out [M][2 h] (ldc), or null
    int pair_h;
    // A operand produced in the kernel from a row-major fp32 source (umma.cuh: row_chunk_*), or null = pre-formatted tiles
    const float *a_src; long long a_ld;  // A(m, k) = a_src[m * a_ld + k] through the pairwise transform when a_pair_half > 0
    int a_pair_half;
};

// value of operand element (r, k) read from a row-major fp32 source; pair_half > 0 applies the pairwise transform
// along the K/column axis of the SOURCE ROW: l0[k] = k < h ? x[k] * x[k + h] : x[k - h]
__device__ __forceinline__ float pair_load(const float *row, int k, int h) {
    if (h > 0) return k < h ? __ldg(row + k) * __ldg(row + k + h) : __ldg(row + k - h);
    return __ldg(row + k);
}

// "rows": operand row r = source row r, k = source column.  One thread = 8 consecutive k of one row.
template <int RT>
__global__ void ugemm_format_rows_kernel(const float *__restrict__ src, long long ld, int nrows, int K, int pair_half,
                                         int n_rt, int n_ks, unsigned char *__restrict__ out) {
    // neighbouring threads take neighbouring ROWS of the same 8-wide k chunk: their 16-byte outputs are contiguous
    // in the tile (512 B per warp) and each reads one full 32-byte sector of its row
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    const int NR = n_rt * RT;
    if (i >= 1LL * NR * n_ks * 2) return;
    const int r = (int)(i % NR), kc = (int)(i / NR);
    float v[8];
    if (r < nrows && pair_half == 0 && kc * 8 + 8 <= K && (ld & 3) == 0) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(src + (size_t)r * ld + kc * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(src + (size_t)r * ld + kc * 8) + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else if (r < nrows && pair_half > 0 && (pair_half & 7) == 0 && kc * 8 + 8 <= K && (ld & 3) == 0) {
        // l0 chunk: product of the two halves of the row, or a copy of the first half
        const int k = kc * 8;
        const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)r * ld + (k < pair_half ? k : k - pair_half));
        const float4 lo = __ldg(p), hi = __ldg(p + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        if (k < pair_half) {
            const float4 *q = reinterpret_cast<const float4 *>(src + (size_t)r * ld + k + pair_half);
            const float4 lo2 = __ldg(q), hi2 = __ldg(q + 1);
            v[0] *= lo2.x; v[1] *= lo2.y; v[2] *= lo2.z; v[3] *= lo2.w; v[4] *= hi2.x; v[5] *= hi2.y; v[6] *= hi2.z; v[7] *= hi2.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 8 + j;
            v[j] = (r < nrows && k < K) ? pair_load(src + (size_t)r * ld, k, pair_half) : 0.0f;
        }
    }
    uint4 o[3];
    split3x8(v, o);
    const int rt = r / RT, rr = r % RT, ks = kc >> 1;
    unsigned char *tile = out + ((size_t)rt * n_ks + ks) * 3 * (RT * 32) + (uint32_t)(kc & 1) * (RT * 16) + (uint32_t)(rr >> 3) * 128 +
                          (uint32_t)(rr & 7) * 16;
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(tile + (uint32_t)t * (RT * 32)) = o[t];
}
// "cols": operand row n = source column n, k = source row (K = number of source rows).  One thread = 8 consecutive
// k (source rows) of one column; neighbouring threads take neighbouring columns (coalesced reads).
// pair_half > 0: the source row is x[2h] and the operand column n is l0[n].
// pair_perm = h > 0 (RT = 256, ncols = 2 h, h a multiple of 128): operand row n of tile t holds source column
// 128 t + n % 256 for the first 128 rows of the tile and h + 128 t + n % 256 - 128 for the last 128, so that one N tile
// carries both members of every pairwise pair (the GEMM's pairwise-backward epilogue).
template <int RT>
__global__ void ugemm_format_cols_kernel(const float *__restrict__ src, long long ld, int K, int ncols, int pair_half,
                                         int pair_perm, int n_rt, int n_ks, unsigned char *__restrict__ out) {
    const long long i = 1LL * blockIdx.x * blockDim.x + threadIdx.x;
    const int NR = n_rt * RT;
    if (i >= 1LL * NR * n_ks * 2) return;
    const int n = (int)(i % NR), kc = (int)(i / NR);
    int col = n;
    if (pair_perm > 0) {
        const int t = n / 256, r = n % 256;
        col = r < 128 ? 128 * t + r : pair_perm + 128 * t + r - 128;
    }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kc * 8 + j;
        v[j] = (n < ncols && k < K) ? pair_load(src + (size_t)k * ld, col, pair_half) : 0.0f;
    }
    uint4 o[3];
    split3x8(v, o);
    const int rt = n / RT, rr = n % RT, ks = kc >> 1;
    unsigned char *tile = out + ((size_t)rt * n_ks + ks) * 3 * (RT * 32) + (uint32_t)(kc & 1) * (RT * 16) + (uint32_t)(rr >> 3) * 128 +
                          (uint32_t)(rr & 7) * 16;
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4 *>(tile + (uint32_t)t * (RT * 32)) = o[t];
}

template <int NT>
struct UGemmCfg {
    static constexpr uint32_t kATile = kGuM * 32, kBTile = NT * 32;
    static constexpr uint32_t kABytes = 3 * kATile, kBBytes = 3 * kBTile;
    static constexpr int kStages = NT >= 256 ? 3 : 4;
    static constexpr uint32_t kTmemCols = NT < 32 ? 32 : NT;
    static constexpr size_t kSmem = 1024 + (size_t)kStages * (kABytes + kBBytes);
};

// grid = (N tiles, M tiles, K splits)
template <int NT>
__global__ void __launch_bounds__(kGuThreads, 2)
ugemm_kernel(const UGemmArgs g) {
    using Cfg = UGemmCfg<NT>;
    constexpr int ST = Cfg::kStages;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + ST;
    uint64_t *done = empty + ST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    unsigned char *sa = smem_raw + 1024;
    unsigned char *sb = sa + ST * Cfg::kABytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nt = blockIdx.x, mt = blockIdx.y, split = blockIdx.z;
    const int ks0 = split * g.ks_per_split, ks1 = min(g.n_ks, ks0 + g.ks_per_split), n_stage = ks1 - ks0;

    const bool a_inline = g.a_src != nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) {
            mbar_init(&full[i], a_inline ? 5 : 1);  // the TMA thread (+ the four producer warps of an in-kernel A operand)
            mbar_init(&empty[i], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const unsigned char *asrc = g.at + ((size_t)mt * g.n_ks + ks0) * Cfg::kABytes;
            const unsigned char *bsrc = g.bt + ((size_t)nt * g.n_ks + ks0) * Cfg::kBBytes;
            for (int j = 0; j < n_stage; ++j) {
                const int st = j % ST;
                if (j >= ST) mbar_wait(&empty[st], ((j / ST) - 1) & 1);
                mbar_arrive_expect_tx(&full[st], a_inline ? Cfg::kBBytes : Cfg::kABytes + Cfg::kBBytes);
                if (!a_inline)
                    tma_bulk_g2s(sa + (uint32_t)st * Cfg::kABytes, asrc + (size_t)j * Cfg::kABytes, Cfg::kABytes, &full[st]);
                tma_bulk_g2s(sb + (uint32_t)st * Cfg::kBBytes, bsrc + (size_t)j * Cfg::kBBytes, Cfg::kBBytes, &full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kGuM, NT);
            for (int j = 0; j < n_stage; ++j) {
                const int st = j % ST;
                mbar_wait(&full[st], (j / ST) & 1);
                tcgen05_fence_after();
                const uint32_t a_base = smem_u32(sa + (uint32_t)st * Cfg::kABytes), b_base = smem_u32(sb + (uint32_t)st * Cfg::kBBytes);
#pragma unroll
                for (int ta = 0; ta < 3; ++ta)
#pragma unroll
                    for (int tb = 0; ta + tb < 3; ++tb)
                        umma_bf16(tmem_acc, umma_smem_desc(a_base + (uint32_t)ta * Cfg::kATile, kGuM * 16, 128),
                                  umma_smem_desc(b_base + (uint32_t)tb * Cfg::kBTile, NT * 16, 128), idesc, (j | ta | tb) ? 1u : 0u);
                umma_commit(&empty[st]);
            }
            umma_commit(done);
        }
    } else {
        const int q = warp & 3, m = mt * kGuM + q * 32 + lane;
        if (a_inline) {
            // ---- A producers: the four epilogue warps build the stage's A tiles from the fp32 rows while the main loop
            // runs (thread = tile row; the loads of k-step j + 1 are in flight while k-step j is split and stored) ----
            const bool live = m < g.M;
            const float *row = g.a_src + (size_t)min(m, g.M - 1) * g.a_ld;
            const int h = g.a_pair_half;
            unsigned char *dst0 = sa + (uint32_t)(q * 32 + lane) * 16;
            RowChunk nxt;
            row_chunk_load(nxt, row, ks0 * 16, h, live);
            for (int j = 0; j < n_stage; ++j) {
                const int st = j % ST, k0 = (ks0 + j) * 16;
                const RowChunk cur = nxt;
                if (j + 1 < n_stage) row_chunk_load(nxt, row, k0 + 16, h, live);
                if (j >= ST) mbar_wait(&empty[st], ((j / ST) - 1) & 1);
                row_chunk_store(cur, h > 0 && k0 < h, dst0 + (uint32_t)st * Cfg::kABytes, Cfg::kATile);
                fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st]);
            }
        }
        mbar_wait(done, 0);
        tcgen05_fence_after();
        const uint32_t tbase = tmem_acc + ((uint32_t)(q * 32) << 16);
        float *crow = g.C + (size_t)split * g.c_split_stride + (size_t)min(m, g.M - 1) * g.ldc;
        const float *mrow = g.mask ? g.mask + (size_t)min(m, g.M - 1) * g.ldm : nullptr;
        if (NT == 256 && g.pair_ft) {
            // Pairwise backward on the accumulator rows: columns c (product slot) and 128 + c (pass-through slot) of this tile.
            // A thread owns one ROW of the accumulator, but global memory wants a warp on one row at a time (the first
            // version read ft and wrote g_ft 64 bytes per thread and row: 8 KB in flight per SM, 79 us for 134 MB).  So each
            // warp transposes its 32 rows through the (by now idle) operand ring, 64 pair columns at a time, and then walks
            // them row by row: 128-byte loads of ft / stores of g_ft per instruction, eight rows in flight.
            constexpr int kPitch = 129;  // odd: lane r writing row r and lane c reading column c are both conflict-free
            float *S = reinterpret_cast<float *>(sa) + (size_t)q * 32 * kPitch;
            const int h = g.pair_h;
            const int m0 = mt * kGuM + q * 32;
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float gp[16], ga[16];
                    tmem_ld16(tbase + (uint32_t)(cb * 64 + c0), gp);
                    tmem_ld16(tbase + (uint32_t)(128 + cb * 64 + c0), ga);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        S[lane * kPitch + c0 + u] = gp[u];
                        S[lane * kPitch + 64 + c0 + u] = ga[u];
                    }
                }
                __syncwarp();
                const int i0 = nt * 128 + cb * 64;
#pragma unroll 1
                for (int r0 = 0; r0 < 32; r0 += 8) {
                    // eight rows' loads first (row index clamped, not branched on: a branch per row keeps the compiler from
                    // batching them and the loop then pays one DRAM latency per row), then the arithmetic and the stores
                    float fa[8][2], fb[8][2];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float *frow = g.pair_ft + (size_t)min(m0 + r0 + u, g.M - 1) * g.ldc + i0;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            fa[u][hh] = __ldg(frow + lane + 32 * hh);
                            fb[u][hh] = __ldg(frow + h + lane + 32 * hh);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int rr = r0 + u;
                        if (m0 + rr < g.M) {  // warp-uniform
                            float *orow = g.C + (size_t)(m0 + rr) * g.ldc + i0;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int c = lane + 32 * hh;
                                const float p_ = S[rr * kPitch + c], a_ = S[rr * kPitch + 64 + c];
                                orow[c] = fmaf(p_, fb[u][hh], a_);
                                orow[h + c] = p_ * fa[u][hh];
                            }
                        }
                    }
                }
                __syncwarp();
            }
        } else
#pragma unroll 2
        for (int c0 = 0; c0 < NT; c0 += 16) {
            const int n0 = nt * NT + c0;
            float v[16];
            tmem_ld16(tbase + (uint32_t)c0, v);
            tmem_ld_wait();
            if (m < g.M && n0 < g.N) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int n = n0 + u;
                    if (n < g.N) {
                        float x = v[u];
                        if (g.bias) x += __ldg(g.bias + n);
                        if (g.relu) x = fmaxf(x, 0.0f);
                        if (mrow) x = __ldg(mrow + n) > 0.0f ? x : 0.0f;
                        v[u] = x;
                    }
                }
                if (n0 + 16 <= g.N && (g.ldc & 3) == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        reinterpret_cast<float4 *>(crow + n0)[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                } else {
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (n0 + u < g.N) crow[n0 + u] = v[u];
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_acc);
}

// ---- host side --------------------------------------------------------------------------------------------------
int ugemm_format_rows(int RT, const float *src, long long ld, int nrows, int K, int pair_half, unsigned char *out,
                      cudaStream_t st) {
    const int n_rt = ceil_div(nrows, RT), n_ks = ceil_div(K, 16);
    const long long n = 1LL * n_rt * RT * n_ks * 2;
    const int grid = (int)((n + 255) / 256);
    if (RT == 128) ugemm_format_rows_kernel<128><<<grid, 256, 0, st>>>(src, ld, nrows, K, pair_half, n_rt, n_ks, out);
    else ugemm_format_rows_kernel<256><<<grid, 256, 0, st>>>(src, ld, nrows, K, pair_half, n_rt, n_ks, out);
    NNUE_CHECK_LAUNCH("ugemm_format_rows_kernel");
    return NNUE_OK;
}
int ugemm_format_cols(int RT, const float *src, long long ld, int K, int ncols, int pair_half, unsigned char *out,
                      cudaStream_t st, int pair_perm) {
    const int n_rt = ceil_div(ncols, RT), n_ks = ceil_div(K, 16);
    const long long n = 1LL * n_rt * RT * n_ks * 2;
    const int grid = (int)((n + 255) / 256);
    if (pair_perm > 0 && (RT != 256 || pair_perm % 128 || ncols != 2 * pair_perm)) return NNUE_ERR_INVALID_ARG;
    if (RT == 128) ugemm_format_cols_kernel<128><<<grid, 256, 0, st>>>(src, ld, K, ncols, pair_half, 0, n_rt, n_ks, out);
    else ugemm_format_cols_kernel<256><<<grid, 256, 0, st>>>(src, ld, K, ncols, pair_half, pair_perm, n_rt, n_ks, out);
    NNUE_CHECK_LAUNCH("ugemm_format_cols_kernel");
    return NNUE_OK;
}

// C = epilogue(A B^T) from formatted tiles.  NT = 128 or 256 (rows per B tile); splits >= 1 (split-K partials at
// C + z * c_split_stride).  Returns the number of splits actually used (every split non-empty), or < 0.
int ugemm_launch(int NT, int M, int N, int K, const unsigned char *at, const unsigned char *bt, float *C, long long ldc,
                 const float *bias, int relu, const float *mask, long long ldm, int splits, long long c_split_stride,
                 cudaStream_t st, const float *pair_ft, int pair_h, const float *a_src, long long a_ld, int a_pair_half) {
    UGemmArgs g{};
    if (a_src && (K % 16 || (a_ld & 3) || a_pair_half % 16 || (reinterpret_cast<uintptr_t>(a_src) & 15))) return NNUE_ERR_INVALID_ARG;
    g.a_src = a_src; g.a_ld = a_ld; g.a_pair_half = a_pair_half;
    if (pair_ft && (NT != 256 || splits > 1 || N != 2 * pair_h || pair_h % 128 || (ldc & 3))) return NNUE_ERR_INVALID_ARG;
    g.pair_ft = pair_ft; g.pair_h = pair_h;
    g.M = M; g.N = N; g.n_ks = ceil_div(K, 16);
    g.at = at; g.bt = bt; g.C = C; g.ldc = ldc; g.bias = bias; g.relu = relu; g.mask = mask; g.ldm = ldm;
    if (splits < 1) splits = 1;
    g.ks_per_split = ceil_div(g.n_ks, splits);
    const int nz = ceil_div(g.n_ks, g.ks_per_split);
    g.c_split_stride = c_split_stride;
    const dim3 grid(ceil_div(N, NT), ceil_div(M, kGuM), nz);
    if (NT == 128) {
        NNUE_CUDA_TRY(cudaFuncSetAttribute(ugemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UGemmCfg<128>::kSmem));
        ugemm_kernel<128><<<grid, kGuThreads, UGemmCfg<128>::kSmem, st>>>(g);
    } else {
        NNUE_CUDA_TRY(cudaFuncSetAttribute(ugemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UGemmCfg<256>::kSmem));
        ugemm_kernel<256><<<grid, kGuThreads, UGemmCfg<256>::kSmem, st>>>(g);
    }
    NNUE_CHECK_LAUNCH("ugemm_kernel");
    return nz;
}

}  // namespace nnue
