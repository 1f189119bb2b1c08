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
// ---- dense register-stationary backward kernels (small tables), see ft.cu ------------------------
constexpr int kTileTS = 32;      // samples per staged g_ft tile (= one transposed bitmask word)
constexpr int kTileStages = 4;
constexpr int kOwnWarps = 16;    // row-owner warps per CTA
struct OwnPlan {
    bool ok;
    int NH;            // CTAs ("roles") needed to cover the NW bitmask words
    int nq;            // sample streams; grid = nq * NH
    int grid, ntiles;
    size_t smem;
};
inline bool dense_shape_ok(const nnue_shape &s) { return s.L1 == 64 || s.L1 == 32; }
// ---- tcgen05 / TMEM contractions (ft_umma.cu): any L1 that is a multiple of 64 -----------------------------
constexpr int kUmmaNCols = 64;    // forward / weight gradient: table columns per CTA (x 3 terms = UMMA N 192)
constexpr int kUmmaGbinN = 256;   // value gradient: padded positions per CTA (UMMA N)
// ---- which formulation of the feature transformer serves a shape ----------------------------------------------------
// dense : the three contractions as bitmask GEMMs on the tensor cores (ft_umma.cu).  Work is independent of how many
//         positions are active: 2 B PP L1 flops x (3 + 3 + 6) exact split-bf16 term products per step, in M tiles of 128
//         samples, plus re-formatting the table into operand tiles (two passes over it) and a K loop of NW stages whose
//         latency shows at small batches.  The table is re-read once per wave of M tiles, not once per sample.
// gather: index-driven (nnue.py:686-710; ft_gather.cu, ft.cu): every active (sample, position) pair moves one
//         table row (forward, value gradient) or one g_ft row (weight gradient): 3 x nnz x L1 x 4 bytes per step.
// Measured on B200 at SURVEY config I (F = 65536, L1 = 1024, 0.43 of the positions active; profiles/r2_ft_compare_*):
//   batch 512: dense 0.62 / 0.46 / 0.68 ms (forward / weight gradient / value gradient), gather 9.5 / 2.7 / 11.8 ms;
//   batch   1: dense 0.63 /  -   / 0.22 ms,                                              gather 0.036 / - / 0.044 ms;
// the crossover lies near 16 samples.  The model below reproduces that with the measured rates: the tcgen05 kernels
// sustain ~0.55 of the measured sustained bf16 peak over the three contractions, a K-loop stage costs ~0.22 us, the
// formatting passes run at ~0.8 of the HBM copy bandwidth, and the gather kernels move ~0.65 of it.  A row of L1 floats
// gathered per ACTIVE position costs more than 12 flops per position and column on a 1.4 PFLOP/s pipe as soon as more
// than ~2 % of the positions are active, so at the density of the reference model (0.33 - 0.43) the dense
// form wins for every batch that fills a few M tiles; the gather form serves single-digit batches of large tables and
// models whose thresholds have trained the density down.  `ft_density_permille` (default 400) tells the model what to
// expect; `ft_form` forces a formulation (1 dense, 2 gather).  Tables small enough to sit in shared memory / L1 (the
// CIFAR configs) always take the dense form: there the choice was measured per kernel in round 1.
constexpr double kTensorPeakFlops = 1384.6e12, kHbmPeakBytes = 6464.9e9;
struct FtCost {
    double dense_s, gather_s;
};
inline FtCost ft_cost(const nnue_shape &s) {
    FtCost c{};
    const double B128 = ceil_div(s.B, 128) * 128.0, L1 = s.L1, table = (double)s.F * s.L1 * 4.0;
    const double density = get_option(kOptFtDensity) / 1000.0;
    c.dense_s = 12.0 * 2.0 * B128 * (double)s.PP * L1 / (0.55 * kTensorPeakFlops)   // tensor work, whole M tiles
                + 2.0 * (table + 1.5 * table) / (0.8 * kHbmPeakBytes)                // two formatting passes: read fp32, write 3 x bf16
                + 0.22e-6 * s.NW;                                                    // forward K loop: one stage per bitmask word
    c.gather_s = 3.0 * density * (double)s.B * (double)s.P * L1 * 4.0 / (0.65 * kHbmPeakBytes);
    return c;
}
inline bool ft_prefers_dense(const nnue_shape &s) {
    const int form = get_option(kOptFtForm);
    if (form == 1) return true;
    if (form == 2) return false;
    if ((double)s.F * s.L1 * 4.0 <= 1048576.0) return true;  // small tables: on-chip either way, dense measured faster
    const FtCost c = ft_cost(s);
    return c.dense_s <= c.gather_s;
}
inline bool ft_umma_ok(const nnue_shape &s) { return get_option(kOptFtUmma) && s.L1 >= 64 && s.L1 % 64 == 0 && ft_prefers_dense(s); }
// split-bf16 tiles of a [K rows][L1] operand, K a multiple of 32: 3 terms x 2 bytes per element
inline size_t umma_kt_bytes(size_t K, const nnue_shape &s) { return K * s.L1 * 6; }
struct UmmaDwPlan {
    int chunk_blocks, n_chunks;  // K-chunks of `chunk_blocks` 32-sample blocks, one partial buffer each
};
inline UmmaDwPlan plan_ft_dw_umma(const nnue_shape &s) {
    UmmaDwPlan p{};
    const long long tiles = 1LL * ceil_div(s.PP, 128) * (s.L1 / kUmmaNCols);
    long long want = (2LL * kNumSMs + tiles - 1) / tiles;  // two resident CTAs per SM, one wave
    const long long per_chunk = 1LL * (s.P + 1) * s.L1 * 4;
    long long cap = (256LL << 20) / (per_chunk > 0 ? per_chunk : 1);  // partial buffers capped at 256 MiB
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want > s.BW) want = s.BW;
    if (want < 1) want = 1;
    p.chunk_blocks = ceil_div(s.BW, (int)want);
    p.n_chunks = ceil_div(s.BW, p.chunk_blocks);
    return p;
}
// the padded position whose A row is all ones in the weight-gradient GEMM (its output row = the bias gradient):
// the first padding cell of channel 0, or one row past the last M tile when the cell words have no padding
__host__ __device__ inline int umma_bias_pp(const nnue_shape &s) { return (s.Gh * s.Gw) % 32 ? s.Gh * s.Gw : s.PP; }
// value gradient scratch: split g_ft tiles [ceil(B/128)][L1/16][3][128 x 16] | split table tiles [ceil(PP/256)][L1/16][3][256 x 16]
inline size_t ws_ft_gbin_umma(const nnue_shape &s) {
    if (!ft_umma_ok(s)) return 0;
    return align_up((size_t)ceil_div(s.B, 128) * 128 * s.L1 * 6, 256) + align_up((size_t)ceil_div(s.PP, kUmmaGbinN) * kUmmaGbinN * s.L1 * 6, 256);
}
// weight gradient scratch: split g_ft^T tiles | partial[chunk][P + 1][L1] | rows parked for the aliased last row
inline size_t ws_ft_dw_umma(const nnue_shape &s) {
    if (!ft_umma_ok(s)) return 0;
    const size_t alias_rows = (size_t)(s.P > s.F - 1 ? s.P - (s.F - 1) : 0);
    return align_up(umma_kt_bytes((size_t)s.BW * 32, s), 256) + ((size_t)plan_ft_dw_umma(s).n_chunks * (s.P + 1) + alias_rows) * s.L1 * 4;
}
// The two gradients may run concurrently on two streams: disjoint regions.  The FIRST region holds the value gradient's
// operands and, after it, the conv gradient's per-CTA partials ([grid <= kNumSMs][C][28] floats, written at offset 0 by
// nnue_conv_bwd while the weight gradient may still be running on the side stream): it is sized for the larger of the
// two, and the weight gradient's region starts behind it.
inline size_t ws_conv_partials_bound(const nnue_shape &s) { return align_up((size_t)kNumSMs * s.C * 28 * 4, 256); }
size_t ws_input_bwd_fused(const nnue_shape &s);  // scratch of the one-kernel input gradient (input_bwd_fused.cu), 0 if it does not apply
inline size_t ws_bwd_front_umma(const nnue_shape &s) {
    size_t a = ws_ft_gbin_umma(s);
    const size_t b = ws_conv_partials_bound(s), c = align_up(ws_input_bwd_fused(s), 256);
    if (b > a) a = b;
    return c > a ? c : a;
}
inline size_t ws_ft_bwd_umma(const nnue_shape &s) { return ws_bwd_front_umma(s) + align_up(ws_ft_dw_umma(s), 256); }
int launch_ft_fwd_umma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *bias, float *out,
                       void *workspace, cudaStream_t st);
int launch_ft_bwd_dw_umma(const nnue_shape &s, const uint32_t *bits_s, const float *g_ft, void *ws, float *g_w, float *g_b,
                          cudaStream_t st);
// integer accumulate (qinfer.cu): which bitmask words of a channel are walked (stage jj = word (jj / cw_used) * cw_all +
// jj % cw_used, n_words_used stages) and, when `logits` is set, the layer stack run in the kernel's epilogue
struct QAccArgs {
    int cw_all, cw_used, n_words_used;
    const int32_t *w1, *b1, *w2, *b2, *wo, *bo;  // dp4a words [K][n_out] and biases of the three dense layers
    int L2, L3, NC, K2, K3, qone, l2_iscale;
    float l1_scale, out_scale;
    float *logits;                               // [B][NC]; null: the int16 accumulators are written instead
};
// the stack fits the epilogue's per-row scratch: L1 = 64 (one N tile), at most 32 / 32 / 64 outputs per layer
inline bool q_stack_fused_ok(int L1, int L2, int L3, int NC) { return L1 == 64 && L2 <= 32 && L3 <= 32 && NC <= 64; }
int launch_q_accumulate_umma(int B, int NW, int L1, const uint32_t *bits, const unsigned char *tiles, const int32_t *bias,
                             int16_t *acc16, const QAccArgs &qa, cudaStream_t st);
int launch_ft_bwd_gbin_umma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *g_ft, void *workspace,
                            float *gbin, cudaStream_t st, const void *table_tiles = nullptr);
// both tile orders of the table in one buffer (forward tiles | value-gradient tiles); which: bit 0 forward, bit 1 value gradient
inline size_t ft_tables_bytes(const nnue_shape &s) {
    if (!ft_umma_ok(s)) return 0;
    return align_up(umma_kt_bytes((size_t)s.PP, s), 256) + align_up((size_t)ceil_div(s.PP, kUmmaGbinN) * kUmmaGbinN * s.L1 * 6, 256);
}
int launch_ft_format_tables(const nnue_shape &s, const float *w, void *tables, int which, cudaStream_t st);
int launch_ft_fwd_umma_tiles(const nnue_shape &s, const uint32_t *bits_s, const void *tables, const float *bias, float *out,
                             cudaStream_t st);

// ---- TMA-staged row gather (ft_gather.cu): the index-driven forward / value gradient for tables in HBM or L2 ----
bool ft_gather_ok(const nnue_shape &s);       // forward: L1 a multiple of 128
bool ft_gather_dval_ok(const nnue_shape &s);  // value gradient: L1 in {128, 256, 512, 1024}, whole cell words per channel
int ft_gather_dval_grid(const nnue_shape &s);
size_t ws_ft_gather_fwd(const nnue_shape &s);
int launch_ft_gather_fwd(const nnue_shape &s, const uint32_t *bits, const float *w, const float *bias, float *out, void *workspace,
                         size_t workspace_bytes, cudaStream_t st);
int launch_ft_gather_dval(const nnue_shape &s, const uint32_t *bits, const float *w, const float *g_ft, const float *xpad,
                          const float *thr, float *dval, float *thr_partial, cudaStream_t st);

inline OwnPlan plan_ft_bwd_dw_owner(const nnue_shape &s) {
    OwnPlan o{};
    if (!get_option(kOptDwOwner) || !dense_shape_ok(s)) return o;
    o.NH = ceil_div(s.NW, kOwnWarps);
    if (o.NH > kNumSMs) return o;
    o.ntiles = ceil_div(s.B, kTileTS);
    o.nq = kNumSMs / o.NH;
    if (o.nq > o.ntiles) o.nq = o.ntiles;
    if (o.nq < 1) o.nq = 1;
    if ((size_t)o.nq * s.P * s.L1 * 4 > ((size_t)256 << 20)) return o;  // partial buffers capped at 256 MiB
    o.grid = o.nq * o.NH;
    o.smem = 128 + (size_t)kTileStages * kTileTS * s.L1 * 4;
    o.ok = true;
    return o;
}
constexpr int kDvWarps = 8;      // 8 warps x up to 255 registers: kDvCH table rows of L1 floats per lane
constexpr int kDvCH = 2;
struct DvPlan {
    bool ok;
    int NH, nq, grid, ntiles;
    size_t smem;
};
inline DvPlan plan_ft_bwd_dval_dense(const nnue_shape &s) {
    DvPlan d{};
    if (!dense_shape_ok(s)) return d;
    d.NH = ceil_div(ceil_div(s.C, kDvCH) * s.CW, kDvWarps);
    if (d.NH > kNumSMs) return d;
    d.ntiles = ceil_div(s.B, kTileTS);
    d.nq = kNumSMs / d.NH;
    if (d.nq > d.ntiles) d.nq = d.ntiles;
    if (d.nq < 1) d.nq = 1;
    d.grid = d.nq * d.NH;
    d.smem = 128 + (size_t)kTileStages * kTileTS * s.L1 * 4;
    d.ok = true;
    return d;
}
constexpr int kFbWarps = 8;      // both gradients at once: table row + its gradient per lane (2 x L1 registers)
struct FbPlan {
    bool ok;
    int NH, nq, grid, ntiles;
    size_t smem;
};
inline FbPlan plan_ft_bwd_both(const nnue_shape &s) {
    FbPlan f{};
    if (!get_option(kOptFtBwdBoth) || !get_option(kOptDwOwner) || !get_option(kOptInputFused) || !dense_shape_ok(s))
        return f;
    f.NH = ceil_div(s.NW, kFbWarps);
    if (f.NH > kNumSMs) return f;
    f.ntiles = ceil_div(s.B, kTileTS);
    f.nq = kNumSMs / f.NH;
    if (f.nq > f.ntiles) f.nq = f.ntiles;
    if (f.nq < 1) f.nq = 1;
    if ((size_t)f.nq * s.P * s.L1 * 4 > ((size_t)256 << 20)) return f;
    f.grid = f.nq * f.NH;
    f.smem = 128 + (size_t)kTileStages * kTileTS * s.L1 * 4;
    f.ok = true;
    return f;
}
int launch_ft_bwd_dval_dense(const nnue_shape &s, const uint32_t *bits_s, const float *ft_w, const float *g_ft,
                             float *dval, cudaStream_t st);

constexpr int kColsumRows = 256;  // rows per column-sum partial
inline size_t ws_ft_bwd_both(const nnue_shape &s) {
    const FbPlan f = plan_ft_bwd_both(s);
    if (!f.ok) return 0;
    const size_t alias_rows = (size_t)(s.P > s.F - 1 ? s.P - (s.F - 1) : 0);
    return align_up((size_t)f.nq * s.L1 * 4, 256) + ((size_t)f.nq * s.P + alias_rows) * s.L1 * 4;
}
inline size_t ws_ft_bwd_dw(const nnue_shape &s) {
    const size_t alias_rows = (size_t)(s.P > s.F - 1 ? s.P - (s.F - 1) : 0);
    const OwnPlan o = plan_ft_bwd_dw_owner(s);
    if (o.ok)  // bias partials [nq][L1] | row partials [nq][P][L1] | rows parked for the aliased last row
        return align_up((size_t)o.nq * s.L1 * 4, 256) + ((size_t)o.nq * s.P + alias_rows) * s.L1 * 4;
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
// tcgen05 form of the general value gradient: dense masked g_bin by ft_gbin_umma_kernel, then a threshold-gradient
// reduction over (g_bin, activations) in `thr_chunks` sample chunks per channel
inline int thr_chunks(const nnue_shape &s) {
    int n = ceil_div(4 * kNumSMs, s.C);
    if (n > s.B) n = s.B;
    return n < 1 ? 1 : n;
}
inline size_t ws_ft_bwd_dval(const nnue_shape &s) {
    const size_t a = (size_t)dval_grid(s) * s.C * 4;
    const size_t b = ft_umma_ok(s) ? ws_ft_gbin_umma(s) + (size_t)thr_chunks(s) * s.C * 4 : 0;
    return a > b ? a : b;
}

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

// ---- input gradient: dense value gradient (ft.cu) + conv gradient from TMA-staged images (input_bwd.cu) ----
// conv-gradient variant 0: 16 warps x 2 channels of a cell word per warp; 1: 8 warps x 4 channels
constexpr int kInMaxStages = 16;  // ring depth: HBM latency under load is ~1 us, a sample is consumed in ~0.3 us
constexpr int kInHeader = 256;    // bytes of mbarriers in front of the ring (2 x kInMaxStages x 8)
constexpr int kInLag = 2;        // the producer refills a stage this many samples after its release
constexpr size_t kMaxSmemOptin = 227 * 1024;  // sm_100: opt-in dynamic shared memory per CTA
struct InPlan {
    bool fused;
    int CH, WARPS;     // channels of one cell word owned by a warp; warps per CTA (lane 0 of warp 0 is the TMA producer)
    int NH, nq, grid, ST;
    int stage_floats;  // 3 x (H*W + 4 pad) image planes | g_bin row [PP] | activation row [PP], padded to 128 B
    int stage_off;     // byte offset of stage 0 in dynamic shared memory
    size_t smem;
};
inline InPlan plan_input_bwd(const nnue_shape &s) {
    InPlan p{};
    const long long HW = 1LL * s.H * s.W;
    if (!get_option(kOptInputFused) || !(dense_shape_ok(s) || ft_umma_ok(s)) || HW % 4 || HW > 16384) return p;
    p.CH = get_option(kOptInputVariant) == 1 ? 4 : 2;
    p.WARPS = 32 / p.CH;
    const int units = ceil_div(s.C, p.CH) * s.CW;
    p.NH = ceil_div(units, p.WARPS);
    if (p.NH > 8) return p;
    p.stage_floats = (int)align_up((size_t)(3 * (HW + 4) + 2 * s.PP), 32);  // g_bin row + optional activation row
    p.stage_off = (int)align_up(kInHeader + ((size_t)s.C * 28 + align_up((size_t)s.C, 4) + 32 * 28) * 4, 128);
    const size_t room = kMaxSmemOptin - (size_t)p.stage_off;
    int ST = (int)(room / ((size_t)p.stage_floats * 4));
    if (ST > kInMaxStages) ST = kInMaxStages;
    if (ST < 2) return p;
    p.ST = ST;
    p.nq = kNumSMs / p.NH;
    if (p.nq > s.B) p.nq = s.B;
    p.grid = p.nq * p.NH;
    p.smem = (size_t)p.stage_off + (size_t)ST * p.stage_floats * 4;
    p.fused = true;
    return p;
}

// ---- conv / threshold gradient from row-staged images (input_bwd_rows.cu): images beyond the whole-image staging ----
constexpr int kRowsCH = 4;     // channels per warp (4 x 27 accumulators in registers; 2 x 16 warps measured 10 % slower)
constexpr int kRowsWarps = 8;  // warps per CTA: 32 channels
struct RowsPlan {
    bool ok;
    int NCB, nq, grid;   // channel blocks of 32, unit streams per channel block, CTAs
    int NR, ST;          // raster rows a cell word can span (one TMA box each), ring depth
    int WB, tile;        // box width in floats (= W), floats per box (9 WB rounded up to 32)
    int stage_floats, stage_off;
    int dx_off;          // floats from a stage's start to its g_bin [32][32] | activations [32][32] | bitmask words [32]
    size_t smem;
};
inline RowsPlan plan_conv_bwd_rows(const nnue_shape &s) {
    RowsPlan p{};
    if (!get_option(kOptInputRows) || s.W % 4 || s.CW > 256) return p;  // tensor-map rows need a 16-byte pitch
    p.WB = s.W;
    if (p.WB > 256) return p;  // box dimension limit of the TMA unit
    const int cells = s.Gh * s.Gw;
    int max_nr = 1;
    for (int j = 0; j < s.CW; ++j) {
        const int first = (j * 32) / s.Gw, last = (j * 32 + 31 < cells ? j * 32 + 31 : cells - 1) / s.Gw;
        if (last - first + 1 > max_nr) max_nr = last - first + 1;
    }
    p.NR = max_nr;
    if (p.NR > kRowsWarps) return p;  // warp r issues the box of raster row r
    p.tile = (int)align_up((size_t)9 * p.WB, 32);
    p.dx_off = p.NR * p.tile;
    p.stage_floats = p.dx_off + 2048 + 32;
    p.stage_off = (int)align_up((size_t)kInHeader + (size_t)kRowsWarps * kRowsCH * 28 * 4 + (size_t)s.CW * 132, 1024);
    int ST = (int)((kMaxSmemOptin - (size_t)p.stage_off) / ((size_t)p.stage_floats * 4));
    if (ST > kInMaxStages) ST = kInMaxStages;
    if (ST < 3) return p;
    p.ST = ST;
    p.NCB = ceil_div(s.C, kRowsWarps * kRowsCH);
    if (p.NCB > kNumSMs) return p;
    p.nq = kNumSMs / p.NCB;
    const long long units = 1LL * s.B * s.CW;
    if (p.nq > units) p.nq = (int)units;
    p.grid = p.nq * p.NCB;
    p.smem = (size_t)p.stage_off + (size_t)ST * p.stage_floats * 4;
    p.ok = true;
    return p;
}
int launch_conv_bwd_rows(const nnue_shape &s, const RowsPlan &pl, const float *images, const uint32_t *bits_s, const float *dval,
                         const float *xpad, const float *thr, float *partial, cudaStream_t st);

// ---- value + conv + threshold gradients in one kernel, g_bin kept in tensor memory (input_bwd_fused.cu) ----
struct FusedPlan {
    bool ok;
    int nq, rounds, ST, n_ks;            // sample streams (= CTAs), 32-sample rounds per stream, ring depth, k steps
    uint32_t a_bytes, b_bytes;           // one channel's table tiles, one round's g_ft tiles
    uint32_t a_off, b_off, stage_off;    // shared-memory layout
    size_t smem;
    size_t ws_wtiles, ws_gtiles, ws_bytes;  // scratch: table tiles | g_ft tiles | per-CTA partials
};
FusedPlan plan_input_bwd_fused(const nnue_shape &s);
int launch_input_bwd_fused(const nnue_shape &s, const FusedPlan &p, const float *images, const uint32_t *bits_s, const float *xpad,
                           const float *ft_w, const float *g_ft, const float *thr, void *workspace, float **partial_out,
                           cudaStream_t st);

// ---- extraction forward, TMA-staged form (extract.cu) -------------------------------------------------
constexpr int kExtWarps = 8;   // warps per CTA (lane 0 of warp 0 is the TMA producer)
constexpr int kExtCH = 4;      // channels of one cell word owned by a warp (their taps live in registers)
struct ExtPlan {
    bool ok;
    int NH, nq, grid, ST;
    int stage_floats;  // 3 x (H*W + 4 pad) image planes, padded to 128 B
    size_t smem;
};
inline ExtPlan plan_extract_tma(const nnue_shape &s) {
    ExtPlan p{};
    const long long HW = 1LL * s.H * s.W;
    if (!get_option(kOptExtractTma) || HW % 4 || HW > 16384) return p;
    const int units = ceil_div(s.C, kExtCH) * s.CW;
    p.NH = ceil_div(units, kExtWarps);
    if (p.NH > 8) return p;
    p.stage_floats = (int)align_up((size_t)(3 * (HW + 4)), 32);
    int ST = (int)((kMaxSmemOptin - kInHeader) / ((size_t)p.stage_floats * 4));
    if (ST > kInMaxStages) ST = kInMaxStages;
    if (ST < 2) return p;
    p.ST = ST;
    p.nq = kNumSMs / p.NH;
    if (p.nq > s.B) p.nq = s.B;
    p.grid = p.nq * p.NH;
    p.smem = kInHeader + (size_t)ST * p.stage_floats * 4;
    p.ok = true;
    return p;
}

// ---- head (dense layers) -----------------------------------------------------------------
constexpr int kGemmBK = 16;
inline int gemm_splits(int M, int N, int K, int BM, int BN) {
    const long long tiles = 1LL * ceil_div(M, BM) * ceil_div(N, BN);
    long long want = (2LL * kNumSMs + tiles - 1) / tiles;
    const long long max_by_k = ceil_div(K, 8 * kGemmBK);
    if (want > max_by_k) want = max_by_k;
    return want < 1 ? 1 : (int)want;
}
// ---- layer 1 of wide stacks on the tensor cores (gemm_umma.cu) ---------------------------------------------------
// The 1024 -> 128 layer of the reference's "real" config is a dense contraction at the benchmark batch (13 GFLOP per
// step); the config-D stack (64 -> 32 -> 8 -> 10) is not and keeps the fused FMA kernel.
inline bool head_umma_ok(const nnue_shape &s) {
    return get_option(kOptHeadUmma) && s.L1 >= 256 && s.L1 % 16 == 0 && s.L2 >= 16 && s.L2 <= 256 && s.B >= 256;
}
inline int head_umma_nt(const nnue_shape &s) { return s.L2 <= 128 ? 128 : 256; }  // B-tile rows of the forward GEMM
inline size_t ugemm_tile_bytes(int rows, int RT, int K) { return (size_t)ceil_div(rows, RT) * RT * ceil_div(K, 16) * 16 * 6; }
inline int head_umma_wgrad_splits(const nnue_shape &s) {
    const int tiles = ceil_div(s.L1, 256) * ceil_div(s.L2, 128);
    int want = ceil_div(kNumSMs, tiles);
    const int n_ks = ceil_div(s.B, 16);
    if (want > n_ks / 8) want = n_ks / 8;  // at least 8 k-steps per split
    return want < 1 ? 1 : want;
}
// ---- the output layer of many-class heads (1000 classes at the ImageNet-shaped configs) ----------------------------
// Its two backward contractions -- g_w3 = g_logits^T act2 (K = batch) and g_z2 = g_logits W3 (K = classes) -- are skinny
// (32 columns) and ran at 4 - 8 TFLOP/s on the fp32 FMA GEMM kernel: 252 + 139 us of the 2.4 ms I-s step.  As split-bf16
// tcgen05 GEMMs (gemm_umma.cu, exact products, fp32 accumulation) they are bound by formatting g_logits twice.
inline bool head3_umma_ok(const nnue_shape &s) { return get_option(kOptHeadUmma) && s.NC >= 128 && s.B >= 256 && s.L3 <= 128; }
inline int head3_umma_wgrad_splits(const nnue_shape &s) {
    int want = ceil_div(2 * kNumSMs, ceil_div(s.NC, 128));
    const int n_ks = ceil_div(s.B, 16);
    if (want > n_ks / 8) want = n_ks / 8;  // at least 8 k-steps per split
    return want < 1 ? 1 : want;
}
// g_logits rows | W3 cols | g_logits cols | act2 cols | split-K partials [splits][NC][L3] | colsum partials
inline size_t ws_head3_umma_bwd(const nnue_shape &s) {
    if (!head3_umma_ok(s)) return 0;
    return align_up(ugemm_tile_bytes(s.B, 128, s.NC), 256) + align_up(ugemm_tile_bytes(s.L3, 128, s.NC), 256) +
           align_up(ugemm_tile_bytes(s.NC, 128, s.B), 256) + align_up(ugemm_tile_bytes(s.L3, 128, s.B), 256) +
           align_up((size_t)head3_umma_wgrad_splits(s) * s.NC * s.L3 * 4, 256) + align_up((size_t)ceil_div(s.B, 256) * s.NC * 4, 256);
}
// forward scratch: l0 rows tiles | W1 rows tiles
inline size_t ws_head_umma_fwd(const nnue_shape &s) {
    if (!head_umma_ok(s)) return 0;
    return align_up(ugemm_tile_bytes(s.B, 128, s.L1), 256) + align_up(ugemm_tile_bytes(s.L2, head_umma_nt(s), s.L1), 256);
}
// backward scratch: g_z1 rows | W1 cols | g_z1 cols | l0 cols | split-K partials [splits][L2][L1] | colsum partials
inline size_t ws_head_umma_bwd(const nnue_shape &s) {
    if (!head_umma_ok(s)) return 0;
    return align_up(ugemm_tile_bytes(s.B, 128, s.L2), 256) + align_up(ugemm_tile_bytes(s.L1, 256, s.L2), 256) +
           align_up(ugemm_tile_bytes(s.L2, 128, s.B), 256) + align_up(ugemm_tile_bytes(s.L1, 256, s.B), 256) +
           align_up((size_t)head_umma_wgrad_splits(s) * s.L2 * s.L1 * 4, 256) + align_up((size_t)ceil_div(s.B, 256) * s.L2 * 4, 256);
}
int ugemm_format_rows(int RT, const float *src, long long ld, int nrows, int K, int pair_half, unsigned char *out,
                      cudaStream_t st);
int ugemm_format_cols(int RT, const float *src, long long ld, int K, int ncols, int pair_half, unsigned char *out,
                      cudaStream_t st, int pair_perm = 0);
// pair_ft != null (NT = 256, B operand formatted with pair_perm = pair_h): the pairwise backward runs in the epilogue and C is g_ft
int ugemm_launch(int NT, int M, int N, int K, const unsigned char *at, const unsigned char *bt, float *C, long long ldc,
                 const float *bias, int relu, const float *mask, long long ldm, int splits, long long c_split_stride,
                 cudaStream_t st, const float *pair_ft = nullptr, int pair_h = 0, const float *a_src = nullptr, long long a_ld = 0,
                 int a_pair_half = 0);
// a_src != null: the A operand is built inside the kernel from the row-major fp32 source (K % 16 == 0; `at` is unused)
inline bool gemm_inline_a_ok(int K, int pair_half) { return get_option(kOptInlineA) && K % 16 == 0 && K >= 256 && pair_half % 16 == 0; }
// the input gradient of layer 1 with the pairwise backward in the GEMM's epilogue: both members of a pair in one N tile
inline bool head_pair_epilogue_ok(const nnue_shape &s) { return get_option(kOptHeadPairEpi) && head_umma_ok(s) && (s.L1 / 2) % 128 == 0; }
// forward of the stack with optional scratch (head.cu): with ws_head_umma_fwd bytes layer 1 runs on the tensor cores
int head_fwd_ws(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d, const float *w2_d,
                const float *b2_d, const float *w3_d, const float *b3_d, float *act1_d, float *act2_d, float *logits_d,
                void *workspace_d, size_t workspace_bytes, cudaStream_t st);

// split-K partials of the three weight(+bias) gradients, then g_act2, g_act1, g_l0 (+ the tensor-core scratch)
inline size_t ws_head_bwd(const nnue_shape &s) {
    const size_t B = s.B;
    size_t bytes = 0;
    bytes += align_up((size_t)gemm_splits(s.NC, s.L3 + 1, s.B, 64, 64) * s.NC * (s.L3 + 1) * 4, 256);
    bytes += align_up((size_t)gemm_splits(s.L3, s.L2 + 1, s.B, 64, 64) * s.L3 * (s.L2 + 1) * 4, 256);
    bytes += align_up((size_t)gemm_splits(s.L2, s.L1 + 1, s.B, 64, 64) * s.L2 * (s.L1 + 1) * 4, 256);
    bytes += align_up(B * s.L3 * 4, 256) + align_up(B * s.L2 * 4, 256) + align_up(B * s.L1 * 4, 256);
    return bytes + ws_head_umma_bwd(s) + ws_head3_umma_bwd(s);
}
inline size_t ws_input_bwd(const nnue_shape &s) {
    const InPlan p = plan_input_bwd(s);
    if (p.fused) {  // gbin plane | max(conv-gradient partials, value-gradient operand scratch)
        size_t rest = (size_t)p.grid * s.C * 28 * 4;
        if (ws_ft_gbin_umma(s) > rest) rest = ws_ft_gbin_umma(s);
        return align_up((size_t)s.B * s.PP * 4, 256) + rest;
    }
    size_t a = ws_ft_bwd_dval(s);
    const size_t b = ws_extract_bwd(s);
    if (b > a) a = b;
    const RowsPlan rp = plan_conv_bwd_rows(s);
    const size_t rows_ws = align_up((size_t)rp.grid * s.C * 28 * 4, 256) + ws_ft_gbin_umma(s);
    if (rp.ok && rows_ws > a) a = rows_ws;
    return 2 * align_up((size_t)s.B * s.PP * 4, 256) + a;
}
// pre-threshold conv activations in padded-position layout (extract.cu); used by the general input-gradient path
int extract_xpad(const nnue_shape &s, const float *images, const float *conv_w, const float *thr, float *xpad,
                 cudaStream_t st);
// ---- feature-transformer contractions on the tensor cores (ft_mma.cu) ------------------------------
constexpr int kMmaGP = 2;              // column-block pairs per warp: a warp's tile is 32 columns wide
constexpr int kMmaFwdThreads = 512;   // forward: 16 warps per persistent CTA
constexpr int kMmaGbinThreads = 384;  // value gradient: 12 warps per persistent CTA (up to 168 registers each)
constexpr int kMmaDwWarps = 8;        // weight gradient: words (warps) per CTA, two CTAs per SM
constexpr int kGbinSplit = 8;         // value gradient: position slices (one persistent CTA group each)
struct MmaPlan {
    bool ok;
    int n_chunks, chunk_blocks;  // weight gradient: K-chunks of `chunk_blocks` 32-sample blocks
};
inline size_t mma_fwd_smem(const nnue_shape &s) { return 128 + (size_t)(s.PP / 16) * 3 * kMmaGP * 512; }
inline size_t mma_gbin_smem(const nnue_shape &s) {
    return 128 + (size_t)ceil_div(s.NW, kGbinSplit) * 4 * 3 * (s.L1 / 32) * 512;
}
inline MmaPlan plan_ft_mma(const nnue_shape &s) {
    MmaPlan m{};
    if (!get_option(kOptFtMma) || !dense_shape_ok(s) || get_option(kOptFtForm) == 2) return m;
    if (mma_fwd_smem(s) > kMaxSmemOptin || mma_gbin_smem(s) > kMaxSmemOptin) return m;  // table slice must fit
    // two resident CTAs per SM, two waves
    const int ctas_per_chunk = ceil_div(s.NW, kMmaDwWarps) * (s.L1 / (16 * kMmaGP));
    int want = ceil_div(2 * kNumSMs, ctas_per_chunk);  // two resident CTAs per SM, one wave
    if (want < 1) want = 1;
    m.chunk_blocks = ceil_div(s.BW, want);
    if (m.chunk_blocks > 16) m.chunk_blocks = 16;  // 96 KB of fragments per CTA at most (two CTAs per SM)
    m.n_chunks = ceil_div(s.BW, m.chunk_blocks);
    if ((size_t)m.n_chunks * s.P * s.L1 * 4 > ((size_t)1 << 30)) return m;
    m.ok = true;
    return m;
}
inline size_t mma_wfrag_bytes(const nnue_shape &s) { return (size_t)(s.PP / 16) * 3 * (s.L1 / 16) * 32 * 16; }
inline size_t mma_gfrag_bytes(const nnue_shape &s) { return (size_t)(s.BW * 2) * 3 * (s.L1 / 16) * 32 * 16; }
inline size_t ws_ft_fwd(const nnue_shape &s) {
    const size_t a = plan_ft_mma(s).ok ? mma_wfrag_bytes(s) : 0, b = ft_umma_ok(s) ? umma_kt_bytes((size_t)s.PP, s) : 0;
    return a > b ? a : b;
}
// first region of the warp-level MMA backward: the table fragments of the value gradient, then (reused) the conv
// gradient's per-CTA partials -- see ws_bwd_front_umma
inline size_t ws_bwd_front_mma(const nnue_shape &s) {
    const size_t a = align_up(mma_wfrag_bytes(s), 256), b = ws_conv_partials_bound(s);
    return a > b ? a : b;
}
inline size_t ws_ft_bwd_mma(const nnue_shape &s) {
    const MmaPlan m = plan_ft_mma(s);
    if (!m.ok) return 0;
    const size_t alias_rows = (size_t)(s.P > s.F - 1 ? s.P - (s.F - 1) : 0);
    return ws_bwd_front_mma(s) + align_up(mma_gfrag_bytes(s), 256) +
           align_up((size_t)m.n_chunks * s.L1 * 4, 256) + ((size_t)m.n_chunks * s.P + alias_rows) * s.L1 * 4;
}
int launch_ft_fwd_mma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *bias, float *out,
                      void *workspace, cudaStream_t st);
int launch_ft_bwd_dw_mma(const nnue_shape &s, const uint32_t *bits_s, const float *g_ft, uint4 *gfrag, float *partial,
                         float *bias_partial, int *n_chunks, cudaStream_t st);
int launch_ft_bwd_gbin_mma(const nnue_shape &s, const uint32_t *bits_s, const float *w, const float *g_ft, uint4 *wfrag,
                           float *gbin, cudaStream_t st);

// ---- fused head training step (head_fused.cu): small stacks only ---------------------------------
constexpr int kHeadTile = 128;       // samples (= threads) per CTA tile
constexpr int kHeadPartial = 2492;   // floats per per-CTA gradient block (HeadLayout<64,32,8,16>::pTotal)
inline bool head_train_fused_ok(const nnue_shape &s) {
    return get_option(kOptHeadFused) && s.L1 % 2 == 0 && s.L1 <= 64 && s.L2 <= 32 && s.L3 <= 8 && s.NC <= 16;
}
inline int head_train_grid(const nnue_shape &s) {
    const int ntiles = ceil_div(s.B, kHeadTile);
    return ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs;
}
// ---- layers 2 - 3 + cross-entropy + their backward in one kernel (head_mid.cu) for wide stacks with a small tail ----
constexpr int kHeadMidPartial = 4788;  // floats per per-CTA gradient block (MidLayout::pTotal)
inline bool head_mid_ok(const nnue_shape &s) {
    return get_option(kOptHeadMid) && !head_train_fused_ok(s) && s.L2 % 4 == 0 && s.L2 <= 128 && s.L3 <= 32 && s.NC <= 16;
}
inline int head_mid_grid(const nnue_shape &s) {  // persistent CTAs over 128-sample tiles (option head_mid > 1 caps the grid: tests)
    const int t = ceil_div(s.B, 128), cap = get_option(kOptHeadMid) > 1 ? get_option(kOptHeadMid) : kNumSMs;
    return t < cap ? t : (cap < kNumSMs ? cap : kNumSMs);
}
int launch_head_mid(const nnue_shape &s, const float *act1, const int64_t *labels, float inv_count, const float *w2, const float *b2,
                    const float *w3, const float *b3, float *loss, float *g_z1, float *g_w2, float *g_b2, float *g_w3, float *g_b3,
                    float *g_b1, float *partial, cudaStream_t st);
// pieces of head.cu shared with nnue_head_train
struct HeadBwdWs { float *p3, *p2, *p1, *g_act2, *g_act1, *g_l0; char *rest; };
HeadBwdWs carve_head_bwd(const nnue_shape &s, void *workspace_d);
int head_layer1_fwd(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d, float *act1_d,
                    void *workspace_d, size_t workspace_bytes, cudaStream_t st);
// a second stream with its own scratch for the layer-1 weight-gradient chain (nnue_head_train_overlapped)
struct HeadSide { cudaStream_t stream; void *ws; size_t ws_bytes; cudaEvent_t ready; };
// side scratch: g_z1 [B][L2] | g_z1^T tiles | l0^T tiles | split-K partials | column-sum partials
inline size_t ws_head_side(const nnue_shape &s) {
    if (!head_umma_ok(s) || !head_mid_ok(s)) return 0;
    return align_up((size_t)s.B * s.L2 * 4, 256) + align_up(ugemm_tile_bytes(s.L2, 128, s.B), 256) +
           align_up(ugemm_tile_bytes(s.L1, 256, s.B), 256) + align_up((size_t)head_umma_wgrad_splits(s) * s.L2 * s.L1 * 4, 256) +
           align_up((size_t)ceil_div(s.B, 256) * s.L2 * 4, 256);
}
// (g_b1_d null: the bias gradient has been produced already, by head_mid)
int head_bwd_layer1(const nnue_shape *s, const HeadBwdWs &bw, const float *ft_out_d, const float *w1_d, float *g_w1_d,
                    float *g_b1_d, float *g_ft_d, cudaStream_t st, const HeadSide *side = nullptr);

inline size_t ws_head_train(const nnue_shape &s) {
    if (head_train_fused_ok(s)) return (size_t)head_train_grid(s) * kHeadPartial * 4;
    const size_t B = s.B;
    return align_up(B * s.L2 * 4, 256) + align_up(B * s.L3 * 4, 256) + 2 * align_up(B * s.NC * 4, 256) +
           align_up(B * 4, 256) + ws_head_bwd(s) + ws_head_umma_fwd(s) + align_up((size_t)kNumSMs * kHeadMidPartial * 4, 256);
}
inline size_t ws_ce(int B) { return align_up((size_t)B * 4, 256); }

}  // namespace nnue
