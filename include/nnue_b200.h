/*
 * nnue_b200.h -- C ABI of the B200-native NNUE hot path (libnnue_b200.so).
 *
 * The reference (marict/nnue-vision) has no C plugin ABI for this path: its
 * boundary is the Python module surface of nnue.py plus the .nnue file format
 * and the C++ NNUEEvaluator class.  Every entry point below names the
 * reference interface it stands in for (file:line relative to the reference
 * checkout); INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - pointers named *_d are DEVICE pointers owned by the caller and must stay
 *     valid until the work queued on `stream` has run; *_h are host pointers;
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*);
 *     the float-path calls never allocate and never synchronise -- scratch is
 *     passed in (`workspace_d`, sized by nnue_workspace_bytes);
 *   - return value: 0 = NNUE_OK, negative = error code (nnue_error_string).
 */
#ifndef NNUE_B200_H
#define NNUE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNUE_B200_ABI_VERSION 4

enum {
    NNUE_OK = 0,
    NNUE_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, inconsistent shape */
    NNUE_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels are built for */
    NNUE_ERR_CUDA = -3,          /* a CUDA runtime call or launch failed (see nnue_last_cuda_error) */
    NNUE_ERR_IO = -4,            /* cannot open / short read of a .nnue file */
    NNUE_ERR_FORMAT = -5,        /* .nnue magic / version / architecture check failed */
    NNUE_ERR_WORKSPACE = -6,     /* workspace too small */
    NNUE_ERR_RASTER = -7         /* conv raster does not fit the feature buffer (engine would overrun) */
};

/*
 * Shape of one float-path problem.  Mirrors the constructor arguments of
 * nnue.NNUE (nnue.py:454-464) plus the per-call image size; derived fields are
 * filled by nnue_shape_init and must not be edited afterwards.
 */
typedef struct nnue_shape {
    int32_t B, H, W;          /* batch, image height, image width (images are [B,3,H,W] fp32) */
    int32_t C, G;             /* num_features_per_square, grid_size (GridFeatureSet, nnue.py:81-90) */
    int32_t L1, L2, L3, NC;   /* l1_size, l2_size, l3_size, num_classes */
    int32_t stride;           /* conv stride (model attribute; nnue.py:519 computes it from input_size) */
    /* ---- derived ---- */
    int32_t F;                /* G*G*C rows of the feature-transformer table */
    int32_t Gh, Gw;           /* conv output raster: (H-1)/stride+1, (W-1)/stride+1 */
    int32_t P;                /* C*Gh*Gw flat CHW positions (may exceed F: clamp at nnue.py:701) */
    int32_t CW;               /* ceil(Gh*Gw/32): 32-bit words per channel in the sample-major bitmask */
    int32_t NW;               /* C*CW words per sample in bits_s */
    int32_t PP;               /* NW*32 padded positions (rows of bits_t, columns of dval) */
    int32_t BW;               /* ceil(B/32) words per padded position in bits_t */
} nnue_shape;

int nnue_b200_abi_version(void);
const char *nnue_error_string(int code);
/* text of the last CUDA error seen by this library on the calling thread ("" if none) */
const char *nnue_last_cuda_error(void);

/* Kernels this library has launched in this process so far (optionally reset to 0). */
unsigned long long nnue_launch_count(int reset);
/* Account for kernels launched on the library's behalf without passing through it: a host that replays a captured
 * CUDA graph of these calls adds the graph's kernel count per replay.  Returns the new total. */
unsigned long long nnue_launch_count_add(unsigned long long n);

/*
 * Tuning knobs (process-wide; defaults work).  Keys:
 *   "ft_fwd_staging"  0 = gather table rows from global/L2, 1 = auto (default), 2 = stage the
 *                     whole table in shared memory with bulk TMA copies whenever it fits.
 *   "ft_bwd_dw_owner" 1 (default) = row-owner weight gradient where the shape allows, 0 = always
 *                     the transposed-bitmask segment reduction.
 *   "input_bwd_fused" 1 (default) = dense value-gradient + conv-gradient kernels where the shape
 *                     allows, 0 = always the index-driven kernel pair.
 *   "input_bwd_variant" conv-gradient kernel: 0 = 16 warps x 2 channels, 1 (default) = 8 warps x 4.
 *   "input_bwd_swizzle" 1 (default) = the conv-gradient kernel stages 32 x 32 images with one tensor-map TMA copy in
 *                     the 128-byte swizzle mode (bank-conflict-free tap reads), 0 = dense rows by 1-D bulk copies.
 *   "extract_fixed"   1 (default) = extraction forward specialised on the channel count (4/8/16/32) with one
 *                     cell word per warp, 0 = the generic kernel.
 *   "extract_tma"     1 = TMA-staged extraction forward for CIFAR-sized images, 0 (default) = direct loads
 *                     (measured faster at config D: profiles/).
 *   "ft_mma"          1 (default) = tensor-core (bf16-split, fp32-exact products) feature-transformer
 *                     contractions for small tables, 0 = CUDA-core kernels only.
 *   "ft_umma"         1 (default) = tcgen05 / TMEM (UMMA) feature-transformer contractions for every L1 that is
 *                     a multiple of 64, 0 = the warp-level MMA / CUDA-core families.
 *   "input_bwd_onchip" 1 = value + conv + threshold gradients of CIFAR-shaped steps in one kernel with g_bin kept in
 *                     tensor memory (input_bwd_fused.cu), 0 (default) = the two dense kernels through HBM.  Measured at
 *                     config D: the one-kernel stage is 8 us shorter (112 vs 120 us) but its 212 KB / 512-column CTAs
 *                     leave no room for the table gradient's CTAs beside them, which puts ~18 us of that kernel back
 *                     on the critical path: 226 vs 216 us per step.
 *   "input_bwd_rows"  1 (default) = conv / threshold gradients of large images from row-staged tiles
 *                     (input_bwd_rows.cu), 0 = the direct-gather kernel pair.
 *   "ft_gather"       1 (default) = index-driven forward / value gradient through the TMA-staged row gather of
 *                     ft_gather.cu (L1 a multiple of 128), 0 = the direct-load kernels of ft.cu.
 *   "ft_gather_slab"  columns per CTA of the gather forward (128, 256, 512, 1024; 0 = widest that divides L1).
 *   "ft_form"         which formulation of the feature transformer serves a shape: 0 (default) = the cost model of
 *                     plan.cuh (dense bit-GEMM on the tensor cores vs index-driven row gather / segment reduction,
 *                     compared with measured rates), 1 = always dense, 2 = always gather.
 *   "ft_density_permille" fraction of active positions the cost model expects, in 1/1000 (default 400: the reference
 *                     model at initialisation has 330 - 430); the gather form wins below ~20.
 *   "ft_bwd_both"     1 (default) = one kernel for both feature-transformer gradients (small tables).
 *   "q_tc_min_batch"  integer inference: batches of at least this many samples (default 2048; 0 = never) run as
 *                     bitmask -> tcgen05 accumulate -> layer stack instead of the one fused kernel (L1 % 64 == 0).
 *   "head_umma"       1 (default) = layer 1 of wide stacks (L1 >= 256, 16 <= L2 <= 256) as a split-bf16 tcgen05 GEMM
 *                     (fp32-exact products) wherever scratch is passed, 0 = fp32 FMA GEMM kernels.
 *   "head_fused"      1 (default) = one-kernel head training step for small stacks, 0 = layer kernels.
 *   "head_mid"        1 (default) = layers 2 - 3 + cross-entropy + their backward of wide stacks (L2 <= 128, L3 <= 32,
 *                     NC <= 16) in one persistent kernel (head_mid.cu), 0 = the layer kernels; N > 1 caps its grid at N CTAs.
 *   "head_pair_epilogue" 1 (default) = the pairwise backward in the epilogue of the layer-1 input-gradient GEMM
 *                     (tensor-core form, L1 / 2 a multiple of 128), 0 = g_l0 through HBM and a separate kernel.
 *   "gemm_inline_a"   1 = the A operands of the layer-1 forward GEMM and of the wide value-gradient GEMM are split into
 *                     bf16 terms inside the kernels (bit-identical), 0 (default) = formatting kernels.  Measured no faster.
 *   "conv_bwd_packed" 1 (default) = the conv-gradient accumulators of 32 x 32 images updated two taps at a time by the
 *                     packed fp32 FMA (fma.rn.f32x2; bit-identical), 0 = scalar FMAs.
 *   "q_cta_max_batch" integer inference: batches of at most this many samples (default 592) run a CTA per sample, larger
 *                     ones (below q_tc_min_batch) a warp per sample.
 *   "q_conv_fixed"    1 (default) = 32 x 32 images at conv stride 4 take the TMA-staged conv + bitmask kernel in the
 *                     large-batch form, 0 = the general conv kernel.
 *   "q_stack_fused"   the smallest batch (default 8192; 0 = never) whose layer stack runs in the epilogue of the tcgen05
 *                     accumulate (L1 = 64, stack <= 32 / 32 / 64) instead of a third launch.
 * The Python binding applies NNUE_OPTIONS="key=value,..." from the environment when the library is loaded.
 */
int nnue_set_option(const char *key, int value);

/* Fill the derived fields; returns NNUE_ERR_INVALID_ARG on non-positive sizes. */
int nnue_shape_init(nnue_shape *s, int B, int H, int W, int C, int G, int L1, int L2, int L3, int NC,
                    int stride);
/* Bytes of scratch the float-path calls may use for this shape (max over all of them). */
size_t nnue_workspace_bytes(const nnue_shape *s);

/* ------------------------------------------------------------------------- *
 *  Float training path                                                       *
 * ------------------------------------------------------------------------- */

/*
 * Grid-feature extraction: 3x3 conv (pad 1, stride s, no bias) + per-channel hard
 * threshold -> active-feature bitmask.  Replaces `self.conv(images)` +
 * StraightThroughBinary.forward + NNUE._to_sparse_features
 * (nnue.py:640, 18-25, 590-635).
 *   images_d [B,3,H,W] f32; conv_w_d [C,3,3,3] f32; thr_d [C] f32
 *   bits_s_d [B][NW] u32   bit k of word (c*CW+j) = position (c, cell 32j+k) active
 *   bits_t_d [PP][BW] u32  transposed (position-major) copy, or NULL to skip
 *   xpad_d [B][PP] f32 pre-threshold activations in padded-position layout (kept for the
 *          threshold gradient in nnue_ft_bwd_dval), or NULL when no backward will follow
 *   conv_out_d [B,C,Gh,Gw] f32 the same activations in the reference's NCHW layout, or NULL
 *   nnz_d [B] i32 active positions per sample, or NULL
 */
int nnue_extract_fwd(const nnue_shape *s, const float *images_d, const float *conv_w_d,
                     const float *thr_d, uint32_t *bits_s_d, uint32_t *bits_t_d, float *xpad_d,
                     float *conv_out_d, int32_t *nnz_d, void *stream);

/*
 * bits -> the padded (indices, values) pair NNUE._to_sparse_features returns
 * (nnue.py:590-635): ascending CHW flat indices, -1 / 0.0 padding.
 *   idx_d [B,K] i64, val_d [B,K] f32, K >= max(1, max nnz)
 */
int nnue_sparse_from_bits(const nnue_shape *s, const uint32_t *bits_s_d, int K, int64_t *idx_d,
                          float *val_d, void *stream);

/*
 * Feature-transformer forward on the bitmask (values are all 1.0 on this path):
 * out[b] = bias + sum over active p of W[min(p, F-1)].  Replaces
 * FeatureTransformer.forward as called from NNUE.forward (nnue.py:653, 686-710).
 *   ft_w_d [F,L1] f32; ft_b_d [L1] f32; ft_out_d [B,L1] f32
 * Two forms: the row gather (table staged in shared memory by bulk TMA, rows read as LDS.128,
 * lane groups combined with warp shuffles) and, for small tables when a workspace is passed, a
 * tensor-core contraction bits x (W split into three exact bf16 terms) with fp32 accumulation.
 * workspace_d may be NULL (gather only).
 */
int nnue_ft_fwd(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d,
                const float *ft_b_d, float *ft_out_d, void *workspace_d, size_t workspace_bytes,
                void *stream);

/*
 * General FeatureTransformer.forward(feature_indices, feature_values) (nnue.py:686-710) for
 * callers that use `model.input(idx, val)` directly (tests/test_model.py:1026-1064):
 * arbitrary repeated / unsorted / out-of-range indices, -1 = skip, arbitrary values.
 *   idx_d [B,K] i64; val_d [B,K] f32
 */
int nnue_ft_fwd_indexed(int B, int K, int F, int L1, const int64_t *idx_d, const float *val_d,
                        const float *ft_w_d, const float *ft_b_d, float *ft_out_d, void *stream);
/*
 * The (row, sample, value) triples nnue_ft_bwd_indexed consumes, sorted by table row on the device: a chunked stable
 * counting sort of the B * K pairs of (idx, val) -- key min(idx, F - 1) (the clamp of nnue.py:701), pairs with idx < 0
 * sort behind every row as key F with value 0 -- that keeps (sample, slot) order inside a row, so the segment reduction
 * that follows is deterministic.  Always B * K triples come out: nothing is read back to the host.  Replaces the
 * argsort a reference maintainer would otherwise write in torch around nnue.py:702-708.
 *   idx_d [B,K] i64; val_d [B,K] f32; row_out_d / sample_out_d [B*K] i32; pval_out_d [B*K] f32
 */
size_t nnue_ft_sort_pairs_workspace_bytes(int B, int K, int F);
int nnue_ft_sort_pairs(int B, int K, int F, const int64_t *idx_d, const float *val_d, int32_t *row_out_d,
                       int32_t *sample_out_d, float *pval_out_d, void *workspace_d, size_t workspace_bytes, void *stream);

/*
 * Backward of the indexed form.  The weight gradient is a SORTED SEGMENT REDUCTION: the
 * caller passes the (row, sample, value) triples sorted by row (rows already clamped,
 * -1 entries removed); one warp sums each row's segment, no atomics.
 *   n_pairs triples: row_d i32 ascending, sample_d i32, pval_d f32
 *   g_out_d [B,L1]; g_w_d [F,L1] (fully written, zeros for untouched rows); g_b_d [L1];
 *   g_val_d [B,K] (d out / d val, 0 at idx<0), may be NULL
 */
int nnue_ft_bwd_indexed(int B, int K, int F, int L1, const int64_t *idx_d, const float *ft_w_d,
                        const float *g_out_d, int n_pairs, const int32_t *row_d,
                        const int32_t *sample_d, const float *pval_d, float *g_w_d, float *g_b_d,
                        float *g_val_d, void *stream);

/*
 * Pairwise product + 3-layer head, ReLU fused in the epilogues.  Replaces
 * nnue.py:660-669 + SimpleClassifier.forward (nnue.py:728-738).
 *   w1 [L2,L1] b1 [L2]; w2 [L3,L2] b2 [L3]; w3 [NC,L3] b3 [NC]
 *   act1_d [B,L2], act2_d [B,L3]: post-ReLU activations kept for the backward
 *   logits_d [B,NC]
 */
int nnue_head_fwd(const nnue_shape *s, const float *ft_out_d, const float *w1_d, const float *b1_d,
                  const float *w2_d, const float *b2_d, const float *w3_d, const float *b3_d,
                  float *act1_d, float *act2_d, float *logits_d, void *stream);

/*
 * Mean cross-entropy (train.py:250-254) fused with its gradient.
 *   labels_d [B] i64
 *   loss_d [1] f32 = sum_b CE_b * inv_count            (may be NULL: gradient only)
 *   g_logits_d [B,NC] = (softmax - onehot) * inv_count * g  (may be NULL: loss only), where
 *   g = *g_scale_d, a DEVICE scalar (the upstream gradient of the loss), or 1 when NULL.
 *   inv_count is 1/B on one GPU and 1/(global B) under data parallelism;
 *   per_sample_d [B] (may be NULL) receives the unscaled CE_b.
 */
int nnue_ce_fwd_bwd(int B, int NC, const float *logits_d, const int64_t *labels_d, float inv_count,
                    const float *g_scale_d, float *loss_d, float *per_sample_d, float *g_logits_d,
                    void *workspace_d, size_t workspace_bytes, void *stream);

/*
 * Backward of nnue_head_fwd: parameter gradients of the three Linear layers and the
 * gradient w.r.t. the feature-transformer output (pairwise backward fused).
 *   g_logits_d [B,NC] -> g_w{1,2,3}, g_b{1,2,3}, g_ft_d [B,L1]
 */
int nnue_head_bwd(const nnue_shape *s, const float *g_logits_d, const float *ft_out_d,
                  const float *act1_d, const float *act2_d, const float *w1_d, const float *w2_d,
                  const float *w3_d, float *g_w1_d, float *g_b1_d, float *g_w2_d, float *g_b2_d,
                  float *g_w3_d, float *g_b3_d, float *g_ft_d, void *workspace_d,
                  size_t workspace_bytes, void *stream);

/*
 * The training step of the head in one call: nnue_head_fwd + nnue_ce_fwd_bwd + nnue_head_bwd
 * (nnue.py:660-669, 713-738; train.py:250-254) with every activation kept on chip.  Small stacks
 * (L1 <= 64, L2 <= 32, L3 <= 8, NC <= 16) run ONE kernel, a sample per thread; larger ones run
 * the layer kernels on scratch carved from the workspace.
 *   loss_d [1] = sum_b CE_b * inv_count; g_ft_d [B,L1]; g_w*, g_b* as nnue_head_bwd
 */
int nnue_head_is_fused(const nnue_shape *s);   /* 1 when nnue_head_train runs the one-kernel form for this shape */
int nnue_head_uses_umma(const nnue_shape *s);  /* 1 when layer 1 runs as split-bf16 tcgen05 GEMMs (gemm_umma.cu) */
int nnue_head_train(const nnue_shape *s, const float *ft_out_d, const int64_t *labels_d, float inv_count,
                    const float *w1_d, const float *b1_d, const float *w2_d, const float *b2_d,
                    const float *w3_d, const float *b3_d, float *loss_d, float *g_ft_d, float *g_w1_d,
                    float *g_b1_d, float *g_w2_d, float *g_b2_d, float *g_w3_d, float *g_b3_d,
                    void *workspace_d, size_t workspace_bytes, void *stream);
/*
 * The same with a second stream: wide stacks whose first layer runs on the tensor cores (nnue_head_uses_umma) hand the
 * layer-1 weight-gradient chain (operand formatting, split-K GEMM, fold: independent of g_ft and of every later stage of
 * the step) to `side_stream`, behind an event recorded on `stream`; its scratch and the g_z1 rows it reads come from
 * `side_workspace_d` (nnue_head_side_workspace_bytes(s) bytes; 0 = this shape has no side chain), which must stay
 * untouched until the side stream has been joined.  g_w1_d is complete on `side_stream`, everything else on `stream`.
 * Null side arguments = nnue_head_train.  autograd has no counterpart: it runs the reference's nodes one after another.
 */
size_t nnue_head_side_workspace_bytes(const nnue_shape *s);
int nnue_head_train_overlapped(const nnue_shape *s, const float *ft_out_d, const int64_t *labels_d, float inv_count,
                               const float *w1_d, const float *b1_d, const float *w2_d, const float *b2_d,
                               const float *w3_d, const float *b3_d, float *loss_d, float *g_ft_d, float *g_w1_d,
                               float *g_b1_d, float *g_w2_d, float *g_b2_d, float *g_w3_d, float *g_b3_d,
                               void *workspace_d, size_t workspace_bytes, void *stream, void *side_workspace_d,
                               size_t side_workspace_bytes, void *side_stream);

/*
 * Feature-transformer weight/bias gradient: a segment reduction over (feature, sample)
 * pairs sorted by feature, no atomics, deterministic.  Rows p >= F fold onto row F-1.
 * Replaces the B IndexBackward / index_put nodes autograd builds for nnue.py:702-708.
 * Two forms, chosen by shape (nnue_wants_transposed_bits tells which):
 *   - row-owner (L1 <= 64): a lane owns one table row's gradient in registers and walks the
 *     samples in order off bits_s; bits_t_d may be NULL;
 *   - transposed bitmask (any other L1): the bit-matrix transpose is the sort; needs bits_t_d.
 *   bits_s_d [B][NW]; bits_t_d [PP][BW]; g_ft_d [B,L1]; g_w_d [F,L1] fully written; g_b_d [L1]
 */
int nnue_wants_transposed_bits(const nnue_shape *s);
/*
 * Both feature-transformer gradients in one pass over g_ft (small tables; returns
 * NNUE_ERR_UNSUPPORTED otherwise -- ask nnue_ft_bwd_is_fused first): g_w / g_b as nnue_ft_bwd_dw,
 * gbin_d [B][PP] as nnue_ft_bwd_gbin.  A lane keeps its table row and that row's gradient in registers.
 */
int nnue_ft_bwd_is_fused(const nnue_shape *s);
/*
 * Pre-formatted table tiles for the tcgen05 kernels.  nnue_ft_fwd / nnue_ft_bwd_gbin re-format the table into
 * split-bf16 tiles on every call; the formatting depends only on the weights, so a training step may do it once, on
 * another stream, while the images are being extracted, and then call the *_tables forms.
 *   nnue_ft_tables_bytes: size of the tile buffer, 0 when the tcgen05 kernels do not serve the shape.
 */
size_t nnue_ft_tables_bytes(const nnue_shape *s);
int nnue_ft_format_tables(const nnue_shape *s, const float *ft_w_d, void *tables_d, void *stream);
int nnue_ft_fwd_tables(const nnue_shape *s, const uint32_t *bits_s_d, const void *tables_d, const float *ft_b_d,
                       float *ft_out_d, void *stream);
int nnue_ft_bwd_gbin_tables(const nnue_shape *s, const uint32_t *bits_s_d, const void *tables_d, const float *g_ft_d,
                            float *gbin_d, void *workspace_d, size_t workspace_bytes, void *stream);
int nnue_ft_uses_mma(const nnue_shape *s);  /* 1 when the tensor-core contractions serve this shape */
int nnue_ft_uses_umma(const nnue_shape *s); /* 1 when they are the tcgen05 / TMEM kernels of ft_umma.cu (L1 a multiple of 64) */
int nnue_ft_bwd(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d, const float *g_ft_d,
                float *g_w_d, float *g_b_d, float *gbin_d, void *workspace_d, size_t workspace_bytes,
                void *stream);
int nnue_ft_bwd_dw(const nnue_shape *s, const uint32_t *bits_s_d, const uint32_t *bits_t_d,
                   const float *g_ft_d, float *g_w_d, float *g_b_d, void *workspace_d,
                   size_t workspace_bytes, void *stream);

/*
 * Gradient w.r.t. the feature values at ACTIVE positions: dval[b,p] = <W[min(p,F-1)], g_ft[b]>
 * (the autograd edge that reaches conv / threshold, nnue.py:602-603, 705), fused with the
 * straight-through threshold gradient (StraightThroughBinary.backward, nnue.py:36-52, k = 10):
 *   g_thr[c] = -sum over active (b, p in channel c) of dval[b,p] * k * sig * (1 - sig),
 *   sig = sigmoid(k * (x[b,p] - thr[c])).
 *   xpad_d [B][PP] from nnue_extract_fwd; dval_d [B][PP] f32, written only where the bit is set
 */
int nnue_ft_bwd_dval(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d,
                     const float *g_ft_d, const float *xpad_d, const float *thr_d, float *dval_d,
                     float *g_thr_d, void *workspace_d, size_t workspace_bytes, void *stream);

/*
 * Conv weight gradient: g_x = g_bin (straight-through, nnue.py:28-33) so
 * g_conv_w = conv2d_weight(images, g_bin) with g_bin = dval at active positions, 0 elsewhere.
 *   g_conv_w_d [C,3,3,3]
 */
int nnue_extract_bwd(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d,
                     const float *dval_d, float *g_conv_w_d, void *workspace_d,
                     size_t workspace_bytes, void *stream);

/*
 * Everything upstream of the feature transformer's input in one call: the value gradient at
 * active positions, the straight-through threshold gradient and the conv weight gradient
 * (= nnue_ft_bwd_dval + nnue_extract_bwd, nnue.py:28-52, 602-603, 640, 705).  For CIFAR-sized
 * images and L1 <= 64 this is ONE fused kernel that recomputes the conv from TMA-staged image
 * tiles and never materialises dval or the pre-threshold activations; other shapes run the
 * two-kernel path on scratch carved from the workspace.
 *   g_conv_w_d [C,3,3,3]; g_thr_d [C]
 */
int nnue_input_bwd_is_dense(const nnue_shape *s);  /* 1 when the two dense kernels below serve this shape */
/* dense half 1: g_bin[b,p] = bit(b,p) ? <W[min(p,F-1)], g_ft[b]> : 0, gbin_d [B][PP] f32 */
int nnue_ft_bwd_gbin(const nnue_shape *s, const uint32_t *bits_s_d, const float *ft_w_d,
                     const float *g_ft_d, float *gbin_d, void *workspace_d, size_t workspace_bytes,
                     void *stream);  /* with a workspace: tensor-core form; workspace_d may be NULL */
/* dense half 2: g_conv_w = conv2d_weight(images, g_bin); g_thr from the pre-threshold activations:
 * xpad_d [B][PP] as written by nnue_extract_fwd, or NULL to recompute them from the image taps */
int nnue_conv_bwd(const nnue_shape *s, const float *images_d, const float *gbin_d, const float *xpad_d,
                  const float *conv_w_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d,
                  void *workspace_d, size_t workspace_bytes, void *stream);
int nnue_input_bwd(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d,
                   const float *ft_w_d, const float *g_ft_d, const float *conv_w_d, const float *thr_d,
                   float *g_conv_w_d, float *g_thr_d, void *workspace_d, size_t workspace_bytes,
                   void *stream);
/*
 * The dense pair in ONE kernel for CIFAR-shaped steps (32 x 32 images, 97..128 raster cells, C <= 8, L1 32 / 64, stored
 * activations): the value gradient is computed on the tensor cores inside the conv-gradient CTA and read by its
 * consumers straight out of tensor memory -- g_bin [B][PP] is never written.  nnue_input_bwd_fused_ok tells whether
 * the shape qualifies AND the option "input_bwd_onchip" is on (off by default: see its note); scratch from
 * nnue_workspace_bytes.
 */
int nnue_input_bwd_fused_ok(const nnue_shape *s);
int nnue_input_bwd_fused(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *xpad_d,
                         const float *ft_w_d, const float *g_ft_d, const float *thr_d, float *g_conv_w_d, float *g_thr_d,
                         void *workspace_d, size_t workspace_bytes, void *stream);
/* Same, with the pre-threshold activations the forward stored (xpad_d [B][PP] from nnue_extract_fwd; NULL = recompute
 * them from the images, which is what nnue_input_bwd does).  nnue_input_bwd_wants_activations: 1 when passing them
 * saves a pass over the images for this shape (ImageNet-sized input). */
int nnue_input_bwd_wants_activations(const nnue_shape *s);
int nnue_input_bwd_stored(const nnue_shape *s, const float *images_d, const uint32_t *bits_s_d, const float *xpad_d,
                          const float *ft_w_d, const float *g_ft_d, const float *conv_w_d, const float *thr_d,
                          float *g_conv_w_d, float *g_thr_d, void *workspace_d, size_t workspace_bytes,
                          void *stream);

/* ------------------------------------------------------------------------- *
 *  Optimizer step over flat buffers (the step either side of the path)       *
 * ------------------------------------------------------------------------- */

/*
 * torch.nn.utils.clip_grad_norm_ + optimizer.step() of train.py:363-366, 457-471 on ONE flat fp32
 * buffer of n parameters (parameters, gradients and state are views of flat buffers).
 *   nnue_opt_grad_sqnorm: sqnorm_d[0] = sum g^2 (deterministic two-stage sum; workspace from
 *                         nnue_opt_workspace_bytes).  Needed only when max_norm > 0.
 *   nnue_opt_sgd_step:    torch.optim.SGD(lr, momentum, weight_decay) (dampening 0, no nesterov); gradients
 *                         are first scaled in place by min(1, max_norm / (sqrt(sqnorm) + 1e-6)) when
 *                         max_norm > 0; first_step != 0 initialises the momentum buffer with the gradient.
 *   nnue_opt_adam_step:   torch.optim.Adam(lr, (beta1, beta2), eps, weight_decay) (L2 decay, no amsgrad);
 *                         step counts from 1.
 */
/* dst[i] = src[i] * *scale_d (a DEVICE scalar): the upstream gradient of the scalar loss applied to the flat buffer
 * of parameter gradients in one pass (loss.backward() of train.py:352-361 with any upstream factor). */
int nnue_scale_flat(long long n, const float *src_d, const float *scale_d, float *dst_d, void *stream);
size_t nnue_opt_workspace_bytes(long long n);
int nnue_opt_grad_sqnorm(long long n, const float *g_d, float *sqnorm_d, void *workspace_d,
                         size_t workspace_bytes, void *stream);
int nnue_opt_sgd_step(long long n, float *p_d, float *g_d, float *buf_d, float lr, float momentum,
                      float weight_decay, float max_norm, const float *sqnorm_d, int first_step, void *stream);
int nnue_opt_adam_step(long long n, float *p_d, float *g_d, float *m_d, float *v_d, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float max_norm,
                       const float *sqnorm_d, void *stream);

/* ------------------------------------------------------------------------- *
 *  Quantized integer inference path (bit-exact vs serialize.py + engine)     *
 * ------------------------------------------------------------------------- */

typedef struct nnue_qmodel nnue_qmodel; /* opaque: parsed .nnue + device-resident blobs */

/*
 * Parse a .nnue v2 file (format: serialize.py:30-63, 103-136, 394-491) and upload it.
 * Replaces NNUEEvaluator::load_model (engine/src/nnue_engine.cpp:544-657) incl. its
 * validation.  Allocates device memory on the current device; synchronous.
 */
int nnue_q_load(const char *path, nnue_qmodel **out);
int nnue_q_load_memory(const void *bytes_h, size_t n_bytes, nnue_qmodel **out);
void nnue_q_free(nnue_qmodel *m);
/* dims[8] = F, L1, L2, L3, NC, OC (channels per square), G, n_buckets; thr = header threshold */
int nnue_q_dims(const nnue_qmodel *m, int32_t *dims, float *visual_threshold);

/*
 * Batched NNUEEvaluator::evaluate_logits (nnue_engine.cpp:704-734) + the density line of
 * nnue_inference.cpp:50-54: conv -> threshold -> int16-wrap accumulate -> clipped ReLU ->
 * pairwise -> layer stack, one fused kernel.
 *   images_d [B][H*W*3] f32: the raw buffer the engine is handed, read as HWC
 *   logits_d [B,NC] f32 (multiples of 1/64); density_d [B] f32 (may be NULL)
 *   bucket: layer-stack index (the engine ignores it, nnue_engine.cpp:480-481; values
 *           >= n_buckets fall back to 0 as nnue_engine.cpp:705-707 does)
 */
int nnue_q_infer(const nnue_qmodel *m, const float *images_d, int B, int H, int W, int bucket,
                 float *logits_d, float *density_d, void *stream);
/*
 * Same with caller-provided scratch (nnue_q_workspace_bytes(m, B) bytes; 0 when the model has no tensor-core form):
 * batches of at least "q_tc_min_batch" samples then run as bitmask -> tcgen05 accumulate -> layer stack, three
 * launches computing the same integers.  Neither call allocates, synchronises or writes to the model, so one model
 * may serve several streams at once as long as each call has its own scratch; nnue_q_infer (no scratch) always runs
 * the one-kernel form.
 */
/*
 * Host-only: the integer bound the fixed-shape conv kernel compares against instead of dividing.  The engine fires a
 * feature iff (float)clamp(a / conv_scale, -127, 127) > threshold (C++ truncating division; nnue_engine.cpp:93-103, 199).
 * Returns 0 with *a_min set (fires iff a >= *a_min), 1 = never fires, 2 = always fires, or a negative NNUE_ERR_*.
 */
int nnue_q_conv_bound(float threshold, int conv_scale, int32_t *a_min);
size_t nnue_q_workspace_bytes(const nnue_qmodel *m, int B);
int nnue_q_infer_ws(const nnue_qmodel *m, const float *images_d, int B, int H, int W, int bucket,
                    float *logits_d, float *density_d, void *workspace_d, size_t workspace_bytes, void *stream);
/*
 * Same through HOST buffers (the call evaluate.py:126-173 would make instead of one
 * subprocess per sample): H2D, kernel, D2H and a stream synchronise inside the call.
 */
int nnue_q_infer_host(const nnue_qmodel *m, const float *images_h, int B, int H, int W, int bucket,
                      float *logits_h, float *density_h);

/*
 * Incremental accumulators over S independent streams -- the batched form of NNUEEvaluator::refresh_accumulator /
 * update_features / evaluate_incremental / save_ / restore_accumulator (nnue_engine.cpp:739-821,
 * nnue_engine.h:583-594).  acc_d is caller-owned int16 [S][L1] (save / restore = copies of it).
 *   refresh != 0: acc = (int16)bias + rows(added)            (refresh_accumulator)
 *   refresh == 0: acc = acc - rows(removed) + rows(added)    (update_features), int16 wrap-around
 * Feature lists are CSR: *_off_d int32 [S + 1], *_idx_d int32 [nnz]; either pair may be NULL (no features);
 * indices outside [0, F) are ignored as the engine does.
 */
int nnue_q_acc_apply(const nnue_qmodel *m, int S, int refresh, const int32_t *add_off_d, const int32_t *add_idx_d,
                     const int32_t *rem_off_d, const int32_t *rem_idx_d, int16_t *acc_d, void *stream);
/* score_d[s] = LayerStack::forward(clipped acc[s]) -- the single score evaluate_incremental returns
 * (nnue_engine.cpp:382-478, 775-787).  NNUE_ERR_UNSUPPORTED when the file lacks bias L2 of the combined layer. */
int nnue_q_acc_score(const nnue_qmodel *m, int S, const int16_t *acc_d, int bucket, float *score_d, void *stream);

/* ------------------------------------------------------------------------- *
 *  Data-parallel exchange (the reference has no collective)        *
 * ------------------------------------------------------------------------- */

/*
 * One-shot all-reduce (sum, in rank order on every rank, in place) of `n` floats over NVLink peer memory.
 *   peer_recv_h[world]  host array of DEVICE pointers: rank r's symmetric receive area of
 *                       nnue_allreduce_recv_floats(world, n) floats, peer-mapped on every rank
 *   peer_flags_h[world] host array of device pointers: rank r's flag block, int32[nnue_allreduce_max_world()],
 *                       zeroed once before the first call
 *   state_d             local device uint32[2] {CTA counter, epoch}, zero before the first call; the kernel advances
 *                       the epoch itself (no per-step host argument: the launch can be captured in a CUDA graph)
 *   buf_d               local buffer of n floats (n % 4 == 0, 16-byte aligned): input and result; any aligned slice
 *                       of a larger buffer may be exchanged on its own, with its own recv / flags / state
 * Every rank must issue the same sequence of calls per (recv, flags, state) set.  The kernel pushes the local values
 * into every rank's receive area (posted NVLink stores), publishes the epoch, waits for the peers' and sums its own area.
 */
int nnue_allreduce_max_world(void);
/* slices of at most this many floats travel in the flagged 8-byte form (value + epoch in one store, no fence, no flag
 * round): lower latency at twice the bytes; larger ones (n % 4 == 0, 16-byte aligned) in the fenced bulk form */
size_t nnue_allreduce_ll_max_floats(void);
size_t nnue_allreduce_recv_floats(int world, size_t n);
int nnue_allreduce_oneshot(int world, int rank, void *const *peer_recv_h, void *const *peer_flags_h,
                           void *state_d, size_t n, float *buf_d, void *stream);

/*
 * The flagged form (n <= nnue_allreduce_ll_max_floats()) over one or two slices in ONE launch, each slice in a phase:
 *   0  push + collect (= nnue_allreduce_oneshot),  1  push only (waits for nobody),  2  collect only (poll the peers'
 *   values, sum in rank order, advance the slice's epoch).
 * A slice pushed by one call must be collected by a later call on a stream ordered behind it, before it is pushed again.
 * The data-parallel step pushes the slice that is final early (feature transformer, head) from the side stream and
 * collects it at the end of the step in the same launch that exchanges the last slice (conv weights, thresholds): one
 * point per step where ranks wait for each other.
 */
typedef struct nnue_allreduce_slice {
    void *const *peer_recv_h;   /* as nnue_allreduce_oneshot */
    void *const *peer_flags_h;
    void *state_d;
    size_t n;
    float *buf_d;
    int phase;
} nnue_allreduce_slice;
int nnue_allreduce_oneshot_slices(int world, int rank, const nnue_allreduce_slice *slices, int n_slices, void *stream);


#ifdef __cplusplus
}
#endif
#endif /* NNUE_B200_H */
