/*
 * oracle/nnue_int_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's quantized
 * integer inference path (SURVEY.md section 8a rows Q0'..Q6).  It exists only
 * so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check the CUDA path against it.  Nothing under nnue-vision_b200/ may import,
 * link or execute it.
 *
 * Parity status: PINNED.  tests/test_oracle_int.py checks this file
 *   (a) against golden vectors produced in the build container by the
 *       reference's own serialize.py + engine (tests/golden/make_int_golden.py),
 *   (b) against oracle/_ref (the reference engine compiled from its sources)
 *       on randomized cases whenever that build is present.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  The arithmetic is restated, not copied: one
 * general border-aware convolution instead of the reference's interior/border
 * split, a direct threshold scan instead of the bit-packed DynamicGrid, and a
 * scalar accumulate instead of the AVX2 intrinsics -- all three are
 * value-identical to the reference by construction and by test.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    /* header: serialize.py:30-63 / engine/src/nnue_engine.cpp:544-585 */
    uint32_t F, L1, L2, L3, n_buckets;
    float nnue2score, quantized_one, visual_threshold;
    /* conv block: serialize.py:103-136 / nnue_engine.cpp:11-46 */
    float conv_scale;
    uint32_t OC, IC, KH, KW;
    int8_t *conv_w;  /* OC*IC*KH*KW raw bytes (PyTorch OIHW order in the file) */
    int32_t *conv_b; /* OC */
    int G;           /* grid size derived as the engine does, nnue_engine.cpp:593-603 */
    /* feature transformer: serialize.py:394-420 / nnue_engine.cpp:161-186 */
    float ft_scale;
    int16_t *ft_w; /* F*L1 */
    int32_t *ft_b; /* L1 */
    /* layer stack (bucket 0): serialize.py:423-491 / nnue_engine.cpp:283-380 */
    float l1_scale, l2_scale, out_scale, l1_fact_scale;
    int8_t *w1;  /* (L2+1)*L1, only rows < L2 are used by the multiclass head */
    int32_t *b1; /* L2+1 */
    int8_t *w2;  /* L3*(2*L2), only columns < L2 are used */
    int32_t *b2; /* L3 */
    int8_t *wo;  /* NC*L3 */
    int32_t *bo; /* NC */
    uint32_t NC;
} oracle_model;

static int rd(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }
static int rd_u32(FILE *f, uint32_t *v) { return rd(f, v, 4); }
static int rd_f32(FILE *f, float *v) { return rd(f, v, 4); }

static void *rd_blob(FILE *f, size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) return NULL;
    if (rd(f, p, n)) { free(p); return NULL; }
    return p;
}

void oracle_free(oracle_model *m) {
    if (!m) return;
    free(m->conv_w); free(m->conv_b); free(m->ft_w); free(m->ft_b);
    free(m->w1); free(m->b1); free(m->w2); free(m->b2); free(m->wo); free(m->bo);
    free(m);
}

/* .nnue v2 parser.  Layout per serialize.py:500-528 (header, conv, FT, stack);
 * validation mirrors NNUEEvaluator::load_model, nnue_engine.cpp:544-657 and
 * LayerStack::load_from_stream, nnue_engine.cpp:283-380.  Returns NULL on any
 * malformed input (the engine returns false). */
oracle_model *oracle_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    oracle_model *m = (oracle_model *)calloc(1, sizeof(*m));
    char magic[4];
    uint32_t ver, n, rows, cols, layer_type;
    if (!m || rd(f, magic, 4) || memcmp(magic, "NNUE", 4) || rd_u32(f, &ver) || ver != 2) goto bad;
    if (rd_u32(f, &m->F) || rd_u32(f, &m->L1) || rd_u32(f, &m->L2) || rd_u32(f, &m->L3) ||
        rd_u32(f, &m->n_buckets) || rd_f32(f, &m->nnue2score) || rd_f32(f, &m->quantized_one) ||
        rd_f32(f, &m->visual_threshold))
        goto bad;
    /* conv */
    if (rd_u32(f, &layer_type) || rd_f32(f, &m->conv_scale) || rd_u32(f, &m->OC) ||
        rd_u32(f, &m->IC) || rd_u32(f, &m->KH) || rd_u32(f, &m->KW))
        goto bad;
    if (m->IC != 3 || m->KH != 3 || m->KW != 3 || m->OC == 0) goto bad;
    if (!(m->conv_w = (int8_t *)rd_blob(f, (size_t)m->OC * 27))) goto bad;
    if (rd_u32(f, &n) || n != m->OC) goto bad;
    if (!(m->conv_b = (int32_t *)rd_blob(f, (size_t)n * 4))) goto bad;
    /* grid: G = (int)sqrt(F / OC) with integer F/OC, nnue_engine.cpp:593-603 */
    if (m->F == 0 || m->F % m->OC) goto bad;
    m->G = (int)sqrt((double)(m->F / m->OC));
    if ((uint32_t)(m->G * m->G) * m->OC != m->F) goto bad;
    /* feature transformer */
    if (rd_f32(f, &m->ft_scale) || rd_u32(f, &rows) || rd_u32(f, &cols)) goto bad;
    if (rows != m->F || cols != m->L1 || cols == 0) goto bad;
    if (!(m->ft_w = (int16_t *)rd_blob(f, (size_t)rows * cols * 2))) goto bad;
    if (rd_u32(f, &n) || n != m->L1) goto bad;
    if (!(m->ft_b = (int32_t *)rd_blob(f, (size_t)n * 4))) goto bad;
    if (m->n_buckets < 1) goto bad;
    /* layer stack, bucket 0 (the engine ignores the bucket index, nnue_engine.cpp:480-481) */
    if (rd_f32(f, &m->l1_scale) || rd_f32(f, &m->l2_scale) || rd_f32(f, &m->out_scale) ||
        rd_f32(f, &m->l1_fact_scale))
        goto bad;
    if (rd_u32(f, &rows) || rd_u32(f, &cols)) goto bad; /* L1 block: (L2+1) x L1 */
    if (rows != m->L2 + 1 || cols != m->L1 || m->L2 < 1) goto bad;
    if (!(m->w1 = (int8_t *)rd_blob(f, (size_t)rows * cols))) goto bad;
    if (rd_u32(f, &n)) goto bad;
    if (!(m->b1 = (int32_t *)rd_blob(f, (size_t)n * 4))) goto bad;
    if (rd_u32(f, &rows) || rd_u32(f, &cols)) goto bad; /* L1-fact block: skipped by the multiclass head */
    if (cols != m->L1 || rows <= m->L2) goto bad;
    if (fseek(f, (long)((size_t)rows * cols), SEEK_CUR)) goto bad;
    if (rd_u32(f, &n) || fseek(f, (long)((size_t)n * 4), SEEK_CUR)) goto bad;
    if (rd_u32(f, &rows) || rd_u32(f, &cols)) goto bad; /* L2 block: L3 x 2*L2 */
    if (cols != 2 * m->L2 || rows != m->L3) goto bad;
    if (!(m->w2 = (int8_t *)rd_blob(f, (size_t)rows * cols))) goto bad;
    if (rd_u32(f, &n)) goto bad;
    if (!(m->b2 = (int32_t *)rd_blob(f, (size_t)n * 4))) goto bad;
    if (rd_u32(f, &rows) || rd_u32(f, &cols)) goto bad; /* output block: NC x L3 */
    if (cols != m->L3 || rows < 1) goto bad;
    m->NC = rows;
    if (!(m->wo = (int8_t *)rd_blob(f, (size_t)rows * cols))) goto bad;
    if (rd_u32(f, &n)) goto bad;
    if (!(m->bo = (int32_t *)rd_blob(f, (size_t)n * 4))) goto bad;
    fclose(f);
    return m;
bad:
    fclose(f);
    oracle_free(m);
    return NULL;
}

/* dims[] = F, L1, L2, L3, NC, OC, G, n_buckets */
void oracle_dims(const oracle_model *m, int *dims) {
    dims[0] = (int)m->F; dims[1] = (int)m->L1; dims[2] = (int)m->L2; dims[3] = (int)m->L3;
    dims[4] = (int)m->NC; dims[5] = (int)m->OC; dims[6] = m->G; dims[7] = (int)m->n_buckets;
}
float oracle_threshold(const oracle_model *m) { return m->visual_threshold; }

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Engine stride rule: ceil((H-1)/(G-1)), nnue_engine.cpp:710-718 (differs from
 * the Python model's floor rule, nnue.py:519). */
int oracle_stride(const oracle_model *m, int H) {
    int s = m->G > 1 ? (H - 1 + m->G - 2) / (m->G - 1) : (H > 1 ? H : 1);
    return s < 1 ? 1 : s;
}

/*
 * One sample.  img is the raw float buffer the engine is handed, read as HWC.
 * Optional debug outputs (may be NULL): conv_dbg int8[F] (the zero-filled
 * G*G*OC buffer after the conv), acc_dbg int16[L1] (accumulator before the
 * clipped ReLU).  Returns the number of active features, or -1 when the conv
 * raster would not fit the F-byte buffer (the reference would overrun it).
 */
int oracle_eval(const oracle_model *m, const float *img, int H, int W, float *logits,
                int8_t *conv_dbg, int16_t *acc_dbg) {
    const int OC = (int)m->OC, F = (int)m->F, L1 = (int)m->L1, L2 = (int)m->L2, L3 = (int)m->L3;
    const int s = oracle_stride(m, H);
    const int oh = (H - 1) / s + 1, ow = (W - 1) / s + 1; /* nnue_engine.cpp:49-50 with k=3,pad=1 */
    if ((long long)oh * ow * OC > F) return -1;
    int8_t *buf = (int8_t *)calloc((size_t)F, 1); /* zeroed: nnue_engine.cpp:720 */
    int16_t *acc = (int16_t *)malloc((size_t)L1 * 2);
    int16_t *pw = (int16_t *)malloc((size_t)L1 * 2);
    int8_t *h1 = (int8_t *)malloc((size_t)L2);
    int8_t *h2 = (int8_t *)malloc((size_t)L3);
    const int iscale = (int32_t)m->conv_scale;

    /* Q1 ConvLayer::forward, nnue_engine.cpp:48-157.  Pixel quantisation is a
     * float multiply then a truncating cast (:68); weights are indexed
     * [oc][kh][kw][ic] over the file's bytes (:69); the division truncates
     * toward zero (:92); output raster is compact, width ow (:93). */
    for (int oy = 0; oy < oh; ++oy)
        for (int ox = 0; ox < ow; ++ox)
            for (int oc = 0; oc < OC; ++oc) {
                int32_t a = m->conv_b[oc];
                for (int kh = 0; kh < 3; ++kh) {
                    const int iy = oy * s + kh - 1;
                    if (iy < 0 || iy >= H) continue;
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ix = ox * s + kw - 1;
                        if (ix < 0 || ix >= W) continue;
                        for (int ic = 0; ic < 3; ++ic) {
                            const float scaled = img[(iy * W + ix) * 3 + ic] * m->conv_scale;
                            a += (int32_t)scaled * (int32_t)m->conv_w[((oc * 3 + kh) * 3 + kw) * 3 + ic];
                        }
                    }
                }
                buf[(oy * ow + ox) * OC + oc] = (int8_t)clampi(a / iscale, -127, 127);
            }
    if (conv_dbg) memcpy(conv_dbg, buf, (size_t)F);

    /* Q2 DynamicGrid::from_conv_output + extract_features, nnue_engine.h:236-285:
     * feature i is active iff (float)buf[i] > threshold, channels >= 64 never.
     * Q3 ft_forward, simd_scalar.cpp:78-95: int16 accumulate with wraparound,
     * bias truncated to int16 first. */
    for (int i = 0; i < L1; ++i) acc[i] = (int16_t)m->ft_b[i];
    int n_active = 0;
    for (int i = 0; i < F; ++i) {
        if (i % OC >= 64) continue;
        if ((float)buf[i] > m->visual_threshold) {
            ++n_active;
            const int16_t *row = m->ft_w + (size_t)i * L1;
            for (int j = 0; j < L1; ++j) acc[j] = (int16_t)(uint16_t)((uint16_t)acc[j] + (uint16_t)row[j]);
        }
    }
    if (acc_dbg) memcpy(acc_dbg, acc, (size_t)L1 * 2);

    /* Q4 clipped ReLU, nnue_engine.cpp:726-729 */
    const int16_t qone = (int16_t)m->quantized_one;
    for (int i = 0; i < L1; ++i) acc[i] = acc[i] < 0 ? 0 : (acc[i] > qone ? qone : acc[i]);

    /* Q5 LayerStack::forward_multiclass, nnue_engine.cpp:480-536 */
    const int half = L1 / 2;
    for (int i = 0; i < L1; ++i) pw[i] = 0;
    for (int i = 0; i < half; ++i) {
        const int32_t a = acc[i], b = acc[i + half];
        pw[i] = (int16_t)clampi((a * b) / 128, 0, 127);
        pw[i + half] = (int16_t)clampi(a, 0, 127);
    }
    for (int o = 0; o < L2; ++o) { /* dense_forward_scalar, simd_scalar.cpp:117-136 */
        int32_t a = m->b1[o];
        for (int j = 0; j < L1; ++j) a += (int32_t)pw[j] * (int32_t)m->w1[(size_t)o * L1 + j];
        const float r = (float)a / m->l1_scale;
        h1[o] = (int8_t)clampi((int32_t)r, 0, 127);
    }
    const int i2 = (int32_t)m->l2_scale;
    for (int o = 0; o < L3; ++o) { /* nnue_engine.cpp:512-523: row stride 2*L2, first L2 columns */
        int32_t a = m->b2[o];
        for (int j = 0; j < L2; ++j) a += (int32_t)h1[j] * (int32_t)m->w2[(size_t)o * 2 * L2 + j];
        int r = clampi(a / i2, -127, 127);
        h2[o] = (int8_t)(r < 0 ? 0 : r);
    }
    for (uint32_t c = 0; c < m->NC; ++c) { /* nnue_engine.cpp:526-533 */
        int32_t a = m->bo[c];
        for (int j = 0; j < L3; ++j) a += (int32_t)h2[j] * (int32_t)m->wo[(size_t)c * L3 + j];
        logits[c] = (float)a / m->out_scale;
    }
    free(buf); free(acc); free(pw); free(h1); free(h2);
    return n_active;
}

/* Batch wrapper: images [B][H][W][3] raw float bytes, logits [B][NC],
 * density[b] = (float)n_active / F as nnue_inference.cpp:50-54 computes it.
 * Reentrant; callers may split [b0,b1) ranges over threads. */
int oracle_eval_batch(const oracle_model *m, const float *imgs, int b0, int b1, int H, int W,
                      float *logits, float *density) {
    const size_t img_elems = (size_t)H * W * 3;
    for (int b = b0; b < b1; ++b) {
        int n = oracle_eval(m, imgs + (size_t)b * img_elems, H, W, logits + (size_t)b * m->NC, NULL, NULL);
        if (n < 0) return -1;
        density[b] = (float)n / (float)(int)m->F;
    }
    return 0;
}
