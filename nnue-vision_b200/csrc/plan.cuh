// plan.cuh -- launch planning shared by the kernels' host wrappers and nnue_workspace_bytes.
#pragma once
#include "common.cuh"

namespace nnue {

// ---- feature-transformer column geometry -------------------------------------------------
// A table row of L1 floats is walked as float4s.  LPR lanes cooperate on one row chunk of
// CC = 4*LPR columns; with LPR < 32 a warp processes NG = 32/LPR rows at once.
struct ColPlan {
    int LPR;      // lanes per row chunk: 4, 8, 16 or 32;  0 = no vector path (generic kernel)
    int CC;       // columns per chunk (4*LPR)
    int nchunks;  // column chunks per row (L1 / CC)
};
inline ColPlan col_plan(int L1) {
    ColPlan p{0, 0, 0};
    if (L1 % 4) return p;
    const int vl = L1 / 4;
    if (vl >= 32) {
        if (vl % 32) return p;
        p.LPR = 32;
    } else if (vl == 4 || vl == 8 || vl == 16) {
        p.LPR = vl;
    } else {
        return p;
    }
    p.CC = 4 * p.LPR;
    p.nchunks = L1 / p.CC;
    return p;
}

// ---- FT weight-gradient (transposed-bitmask segment reduction) ---------------------------
constexpr int kDwWarps = 8;  // warps per CTA
constexpr int kDwPPW = 8;    // padded positions owned by one warp
constexpr int kDwPPC = kDwWarps * kDwPPW;
struct DwPlan {
    ColPlan col;
    int TS;        // samples per staged g_ft tile (multiple of 32)
    int ntiles;    // ceil(B / TS)
    int tpg;       // tiles per tile-group (one CTA walks them sequentially)
    int ngroups;   // tile groups -> partial buffers
    int pchunks;   // ceil(PP / kDwPPC)
    bool direct;   // single group and no clamp aliasing: write g_w rows directly, no fold
    size_t smem;   // dynamic shared bytes
};
inline DwPlan plan_ft_bwd_dw(const nnue_shape &s) {
    DwPlan d{};
    d.col = col_plan(s.L1);
    const int CC = d.col.LPR ? d.col.CC : 0;
    if (!CC) return d;
    int TS = 256;
    while (TS > 32 && (size_t)TS * CC * 4 > 32 * 1024) TS >>= 1;
    while (TS > 32 && TS / 2 >= s.B) TS >>= 1;
    d.TS = TS;
    d.ntiles = ceil_div(s.B, TS);
    d.pchunks = ceil_div(s.PP, kDwPPC);
    const long long ctas_per_group = 1LL * d.pchunks * d.col.nchunks;
    int want_groups = (int)((4LL * kNumSMs + ctas_per_group - 1) / ctas_per_group);
    if (want_groups < 1) want_groups = 1;
    // partial buffers cost ngroups * P * L1 floats: cap them at 64 MiB
    const long long per_group = 1LL * s.P * s.L1 * 4;
    long long cap = (64LL << 20) / (per_group > 0 ? per_group : 1);
    if (cap < 1) cap = 1;
    if (want_groups > cap) want_groups = (int)cap;
    if (want_groups > d.ntiles) want_groups = d.ntiles;
    d.tpg = ceil_div(d.ntiles, want_groups);
    d.ngroups = ceil_div(d.ntiles, d.tpg);
    d.direct = (d.ngroups == 1 && s.P <= s.F);
    d.smem = (size_t)TS * CC * 4 + 64;
    return d;
}
constexpr int kColsumRows = 256;  // rows per column-sum partial
inline size_t ws_ft_bwd_dw(const nnue_shape &s) {
    const DwPlan d = plan_ft_bwd_dw(s);
    size_t bytes = align_up((size_t)ceil_div(s.B, kColsumRows) * s.L1 * 4, 256);
    if (d.col.LPR && !d.direct)  // per-group partials + the rows parked for the aliased last row
        bytes += ((size_t)d.ngroups * s.P + (size_t)(s.P > s.F - 1 ? s.P - (s.F - 1) : 0)) * s.L1 * 4;
    return bytes;
}

// ---- FT value gradient: per-CTA threshold-gradient partials [grid][C] ----------------------
inline int dval_grid(const nnue_shape &s) {
    long long grid = (s.B + 7) / 8;  // 8 warps (samples) per CTA
    if (grid > 64LL * kNumSMs) grid = 64LL * kNumSMs;
    if (grid < kNumSMs) grid = kNumSMs;  // the staged variant always runs kNumSMs CTAs
    return (int)grid;
}
inline size_t ws_ft_bwd_dval(const nnue_shape &s) { return (size_t)dval_grid(s) * s.C * 4; }

// ---- extraction backward -----------------------------------------------------------------
constexpr int kExbCCH = 2;      // channels per warp (register accumulators: 2 x 27 + 2)
constexpr int kExbThreads = 256;
constexpr int kExbWarps = kExbThreads / 32;
// channel chunks handled side by side by the warps of one CTA (they share the image reads through L1)
inline int exb_cpc(int C) {
    const int nchunks = ceil_div(C, kExbCCH);
    int cpc = 1;
    while (cpc * 2 <= nchunks && cpc * 2 <= kExbWarps) cpc *= 2;
    return cpc;
}
inline int exb_grid_y(int C) { return ceil_div(ceil_div(C, kExbCCH), exb_cpc(C)); }
inline int exb_grid_x(const nnue_shape &s) {
    const long long units = 1LL * s.B * s.CW;
    const int streams = kExbWarps / exb_cpc(s.C);
    const int gy = exb_grid_y(s.C);
    long long gx = (8LL * kNumSMs + gy - 1) / gy;
    const long long max_gx = (units + streams - 1) / streams;
    if (gx > max_gx) gx = max_gx;
    return gx < 1 ? 1 : (int)gx;
}
inline size_t ws_extract_bwd(const nnue_shape &s) { return (size_t)exb_grid_x(s) * s.C * 28 * 4; }

// ---- head (dense layers) -----------------------------------------------------------------
constexpr int kGemmBK = 16;
inline int gemm_splits(int M, int N, int K, int BM, int BN) {
    const long long tiles = 1LL * ceil_div(M, BM) * ceil_div(N, BN);
    long long want = (2LL * kNumSMs + tiles - 1) / tiles;
    const long long max_by_k = ceil_div(K, 8 * kGemmBK);
    if (want > max_by_k) want = max_by_k;
    return want < 1 ? 1 : (int)want;
}
// split-K partials of the three weight(+bias) gradients, then g_act2, g_act1, g_l0
inline size_t ws_head_bwd(const nnue_shape &s) {
    const size_t B = s.B;
    size_t bytes = 0;
    bytes += align_up((size_t)gemm_splits(s.NC, s.L3 + 1, s.B, 64, 64) * s.NC * (s.L3 + 1) * 4, 256);
    bytes += align_up((size_t)gemm_splits(s.L3, s.L2 + 1, s.B, 64, 64) * s.L3 * (s.L2 + 1) * 4, 256);
    bytes += align_up((size_t)gemm_splits(s.L2, s.L1 + 1, s.B, 64, 64) * s.L2 * (s.L1 + 1) * 4, 256);
    bytes += align_up(B * s.L3 * 4, 256) + align_up(B * s.L2 * 4, 256) + align_up(B * s.L1 * 4, 256);
    return bytes;
}
inline size_t ws_ce(int B) { return align_up((size_t)B * 4, 256); }

}  // namespace nnue
