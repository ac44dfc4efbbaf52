/*
 * mltcnn_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See mltcnn_oracle.h for scope and the reference file:line each function follows.
 *
 * Arithmetic: fp32 throughout, BatchNorm applied un-folded in eval mode
 * (eps = 1e-5, nn.BatchNorm2d default) so that this oracle is independent of the
 * BN folding done by the product's weight packer.
 */
#include "mltcnn_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BN_EPS 1e-5f

typedef struct {
    int cin, cout, k, stride; /* k = 1 or 3; pad = k/2 (arch.py:37-48) */
    float *w;                 /* rearranged [kh][kw][cin][cout] */
} conv_t;

typedef struct {
    int c;
    float *gamma, *beta, *mean, *var;
} bn_t;

typedef struct {
    conv_t conv1, conv2, sc;
    bn_t bn1, bn2, scbn;
    int has_sc;
} block_t;

struct mlto_model {
    int arch;          /* 128 = CTU model (mlt_ctu_or_pq_arch.py), 64 / 32 / 16 = smaller-CU model (mlt_cu_or_pq_arch.py) */
    int nstage, nhead; /* 4 stages + 3 heads (CTU) or 5 stages + 4 heads (CU); head i hangs off stage i + 1 */
    conv_t conv1;      /* arch.py:244 (bn1 at :246 is defined but unused, :277-278) */
    block_t blk[5][2]; /* layer0..3(4), two BasicBlocks each (arch.py:247-254,306; mlt_cu_or_pq_arch.py:69-80,130) */
    float *fc_w[4], *fc_b[4];
    int fc_in[4], fc_out[4];
};

/* ---------------------------------------------------------------- loading */

static int read_f32(FILE *f, float *dst, size_t n) { return fread(dst, sizeof(float), n, f) == n ? 0 : -1; }

static int load_conv(FILE *f, conv_t *c, int cin, int cout, int k, int stride)
{
    size_t n = (size_t)cout * cin * k * k;
    float *raw = (float *)malloc(n * sizeof(float));
    c->cin = cin; c->cout = cout; c->k = k; c->stride = stride;
    c->w = (float *)malloc(n * sizeof(float));
    if (!raw || !c->w || read_f32(f, raw, n)) { free(raw); return -1; }
    /* torch OIHW -> [kh][kw][cin][cout] */
    for (int o = 0; o < cout; o++)
        for (int i = 0; i < cin; i++)
            for (int y = 0; y < k; y++)
                for (int x = 0; x < k; x++)
                    c->w[(((size_t)y * k + x) * cin + i) * cout + o] = raw[(((size_t)o * cin + i) * k + y) * k + x];
    free(raw);
    return 0;
}

static int load_bn(FILE *f, bn_t *b, int c)
{
    b->c = c;
    b->gamma = (float *)malloc(4 * (size_t)c * sizeof(float));
    if (!b->gamma) return -1;
    b->beta = b->gamma + c; b->mean = b->beta + c; b->var = b->mean + c;
    return read_f32(f, b->gamma, 4 * (size_t)c);
}

mlto_model *mlto_load(const char *path)
{
    static const int planes_ctu[4] = {32, 64, 128, 256};     /* mlt_ctu_or_pq_arch.py:247-254 */
    static const int planes_cu[5] = {32, 64, 96, 128, 256};  /* mlt_cu_or_pq_arch.py:69-80 */
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    uint32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 0x52544c4du /* "MLTR" */ || hdr[1] != 1 ||
        (hdr[2] != 128 && hdr[2] != 64 && hdr[2] != 32 && hdr[2] != 16)) {
        fclose(f);
        return NULL;
    }
    mlto_model *m = (mlto_model *)calloc(1, sizeof(*m));
    m->arch = (int)hdr[2];
    m->nstage = m->arch == 128 ? 4 : 5;
    m->nhead = m->nstage - 1;
    const int *planes = m->arch == 128 ? planes_ctu : planes_cu;
    int bad = load_conv(f, &m->conv1, 2, 32, 3, 1);
    int in_planes = 32;
    for (int L = 0; L < m->nstage && !bad; L++) {
        for (int b = 0; b < 2 && !bad; b++) {
            block_t *B = &m->blk[L][b];
            int stride = b == 0 ? 2 : 1; /* _make_layer: strides = [2, 1] (arch.py:265-271) */
            bad |= load_conv(f, &B->conv1, in_planes, planes[L], 3, stride);
            bad |= load_bn(f, &B->bn1, planes[L]);
            bad |= load_conv(f, &B->conv2, planes[L], planes[L], 3, 1);
            bad |= load_bn(f, &B->bn2, planes[L]);
            B->has_sc = (stride != 1 || in_planes != planes[L]); /* arch.py:45 */
            if (B->has_sc) {
                bad |= load_conv(f, &B->sc, in_planes, planes[L], 1, stride);
                bad |= load_bn(f, &B->scbn, planes[L]);
            }
            in_planes = planes[L];
        }
    }
    static const int fout[4] = {2, 3, 4, 6}; /* arch.py:241; mlt_cu_or_pq_arch.py:62 */
    for (int i = 0; i < m->nhead && !bad; i++) {
        const int fin_i = planes[i + 1] + 2; /* pooled features + poc + qp */
        m->fc_in[i] = fin_i; m->fc_out[i] = fout[i];
        m->fc_w[i] = (float *)malloc((size_t)fin_i * fout[i] * sizeof(float));
        m->fc_b[i] = (float *)malloc((size_t)fout[i] * sizeof(float));
        bad |= read_f32(f, m->fc_w[i], (size_t)fin_i * fout[i]);
        bad |= read_f32(f, m->fc_b[i], (size_t)fout[i]);
    }
    fclose(f);
    if (bad) { mlto_free(m); return NULL; }
    return m;
}

void mlto_free(mlto_model *m)
{
    if (!m) return;
    free(m->conv1.w);
    for (int L = 0; L < 5; L++)
        for (int b = 0; b < 2; b++) {
            block_t *B = &m->blk[L][b];
            free(B->conv1.w); free(B->conv2.w); free(B->sc.w);
            free(B->bn1.gamma); free(B->bn2.gamma); free(B->scbn.gamma);
        }
    for (int i = 0; i < 4; i++) { free(m->fc_w[i]); free(m->fc_b[i]); }
    free(m);
}

/* ---------------------------------------------------------------- staging */

void mlto_stage(const int16_t *org, int org_stride, const int16_t *pred, int pred_stride, float *x)
{
    /* cv::Mat::convertTo(CV_32FC1, 1.0/1023, 0) on CV_16U computes
     * (float)v * (float)alpha in fp32 (OpenCV cvt_32f; pinned by the cv2 KAT in
     * tests/golden/stage_kat.npz).  (float)(1.0/1023) has bit pattern 0x3A802008. */
    const float alpha = (float)(1.0 / 1023);
    float *x_org = x, *x_resi = x + MLTO_CTU * MLTO_CTU;
    for (int i = 0; i < MLTO_CTU; i++) {
        for (int j = 0; j < MLTO_CTU; j++) {
            /* EncCu.cpp:816,827 : (uint16_t) cast of the Pel (int16) samples */
            uint16_t o = (uint16_t)org[(size_t)i * org_stride + j];
            uint16_t p = (uint16_t)pred[(size_t)i * pred_stride + j];
            /* EncCu.cpp:833 : cv::absdiff on CV_16UC1 = true |o - p| */
            uint16_t r = o > p ? (uint16_t)(o - p) : (uint16_t)(p - o);
            float fo = (float)o * alpha, fr = (float)r * alpha; /* EncCu.cpp:836,838 */
            /* EncCu.cpp:848-867 : clamp to [0,1] */
            if (fo < 0.0f) fo = 0.0f; else if (fo > 1.0f) fo = 1.0f;
            if (fr < 0.0f) fr = 0.0f; else if (fr > 1.0f) fr = 1.0f;
            x_org[i * MLTO_CTU + j] = fo;  /* channel 0 (EncCu.cpp:869-875: cat order org, resi) */
            x_resi[i * MLTO_CTU + j] = fr; /* channel 1 */
        }
    }
}

/* ---------------------------------------------------------------- layers (NHWC fp32) */

static void conv2d(const conv_t *c, const float *in, int h, int w, float *out)
{
    const int k = c->k, pad = k / 2, s = c->stride, ci = c->cin, co = c->cout;
    const int ho = (h + 2 * pad - k) / s + 1, wo = (w + 2 * pad - k) / s + 1;
    for (int oy = 0; oy < ho; oy++) {
        for (int ox = 0; ox < wo; ox++) {
            float *o = out + ((size_t)oy * wo + ox) * co;
            for (int n = 0; n < co; n++) o[n] = 0.0f;
            for (int ky = 0; ky < k; ky++) {
                int iy = oy * s + ky - pad;
                if (iy < 0 || iy >= h) continue;
                for (int kx = 0; kx < k; kx++) {
                    int ix = ox * s + kx - pad;
                    if (ix < 0 || ix >= w) continue;
                    const float *ip = in + ((size_t)iy * w + ix) * ci;
                    const float *wp = c->w + ((size_t)(ky * k + kx) * ci) * co;
                    for (int i = 0; i < ci; i++) {
                        const float v = ip[i];
                        const float *wr = wp + (size_t)i * co;
                        for (int n = 0; n < co; n++) o[n] += v * wr[n];
                    }
                }
            }
        }
    }
}

static void bn_apply(const bn_t *b, float *x, size_t npix, int relu)
{
    for (size_t p = 0; p < npix; p++) {
        float *v = x + p * b->c;
        for (int c = 0; c < b->c; c++) {
            float y = (v[c] - b->mean[c]) / sqrtf(b->var[c] + BN_EPS) * b->gamma[c] + b->beta[c];
            v[c] = (relu && y < 0.0f) ? 0.0f : y;
        }
    }
}

/* BasicBlock.forward, arch.py:52-57. in: [h][w][cin] -> out: [ho][wo][planes] */
static void basic_block(const block_t *B, const float *in, int h, float *out, float *tmp, float *tmp2)
{
    const int ho = (h - 1) / B->conv1.stride + 1, co = B->conv1.cout; /* k=3, pad=1 (and the 1x1, pad=0 shortcut): 128->64 ... 2->1, 1->1 */
    const size_t npix = (size_t)ho * ho;
    conv2d(&B->conv1, in, h, h, tmp);
    bn_apply(&B->bn1, tmp, npix, 1);  /* relu(bn1(conv1(x))) */
    conv2d(&B->conv2, tmp, ho, ho, out);
    bn_apply(&B->bn2, out, npix, 0);  /* bn2(conv2(out)) */
    if (B->has_sc) {
        conv2d(&B->sc, in, h, h, tmp2);
        bn_apply(&B->scbn, tmp2, npix, 0);
        for (size_t i = 0; i < npix * co; i++) out[i] += tmp2[i];
    } else {
        for (size_t i = 0; i < npix * co; i++) out[i] += in[i]; /* identity shortcut */
    }
    for (size_t i = 0; i < npix * co; i++) out[i] = out[i] < 0.0f ? 0.0f : out[i];
}

static void gap_fc(const mlto_model *m, int idx, const float *act, int hw, int c, int poc, int qp,
                   float *logits, float *gap_out)
{
    float feat[258];
    /* F.adaptive_avg_pool2d(out,(1,1)) : arch.py:282,288,294 */
    for (int k = 0; k < c; k++) {
        float s = 0.0f;
        for (int p = 0; p < hw * hw; p++) s += act[(size_t)p * c + k];
        feat[k] = s / (float)(hw * hw);
        if (gap_out) gap_out[k] = feat[k];
    }
    /* torch.cat([lvl, poc, qp], dim=1): raw integers promoted to float (arch.py:274-275,284) */
    feat[c] = (float)poc;
    feat[c + 1] = (float)qp;
    for (int o = 0; o < m->fc_out[idx]; o++) {
        float s = m->fc_b[idx][o];
        const float *wr = m->fc_w[idx] + (size_t)o * m->fc_in[idx];
        for (int k = 0; k < c + 2; k++) s += wr[k] * feat[k];
        logits[o] = s;
    }
}

void mlto_forward_ex(const mlto_model *m, const float *x, int poc, int qp, float logits[9], float *gap1,
                     float *gap2, float *gap3)
{
    const int S = MLTO_CTU;
    float *nhwc = (float *)malloc((size_t)S * S * 2 * sizeof(float));
    float *a = (float *)malloc((size_t)S * S * 32 * sizeof(float));
    float *b = (float *)malloc((size_t)64 * 64 * 32 * sizeof(float));
    float *t1 = (float *)malloc((size_t)64 * 64 * 32 * sizeof(float));
    float *t2 = (float *)malloc((size_t)64 * 64 * 32 * sizeof(float));
    for (int p = 0; p < S * S; p++) { nhwc[2 * p] = x[p]; nhwc[2 * p + 1] = x[S * S + p]; }

    conv2d(&m->conv1, nhwc, S, S, a);                  /* arch.py:278 : conv1, no BN / ReLU */
    basic_block(&m->blk[0][0], a, 128, b, t1, t2);     /* layer0 -> 64x64x32 */
    basic_block(&m->blk[0][1], b, 64, a, t1, t2);
    basic_block(&m->blk[1][0], a, 64, b, t1, t2);      /* layer1 -> 32x32x64 */
    basic_block(&m->blk[1][1], b, 32, a, t1, t2);
    gap_fc(m, 0, a, 32, 64, poc, qp, logits + 0, gap1);   /* branch1, arch.py:281-285 */
    basic_block(&m->blk[2][0], a, 32, b, t1, t2);      /* layer2 -> 16x16x128 */
    basic_block(&m->blk[2][1], b, 16, a, t1, t2);
    gap_fc(m, 1, a, 16, 128, poc, qp, logits + 2, gap2);  /* branch2, arch.py:287-291 */
    basic_block(&m->blk[3][0], a, 16, b, t1, t2);      /* layer3 -> 8x8x256 */
    basic_block(&m->blk[3][1], b, 8, a, t1, t2);
    gap_fc(m, 2, a, 8, 256, poc, qp, logits + 5, gap3);   /* branch3, arch.py:293-297 */
    free(nhwc); free(a); free(b); free(t1); free(t2);
}

/* ---------------------------------------------------------------- smaller-CU model (64 / 32 / 16 px) */

void mlto_cu_stage(int size, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride, float *x)
{
    /* the same lines as mlto_stage, for a cuw x cuh block (EncCu.cpp:810-867 run with cuw = cuh = size) */
    const float alpha = (float)(1.0 / 1023);
    for (int i = 0; i < size; i++)
        for (int j = 0; j < size; j++) {
            uint16_t o = (uint16_t)org[(size_t)i * org_stride + j];
            uint16_t p = (uint16_t)pred[(size_t)i * pred_stride + j];
            uint16_t r = o > p ? (uint16_t)(o - p) : (uint16_t)(p - o);
            float fo = (float)o * alpha, fr = (float)r * alpha;
            if (fo < 0.0f) fo = 0.0f; else if (fo > 1.0f) fo = 1.0f;
            if (fr < 0.0f) fr = 0.0f; else if (fr > 1.0f) fr = 1.0f;
            x[i * size + j] = fo;
            x[size * size + i * size + j] = fr;
        }
}

/* MltCnnL4ORPQv4.forward, mlt_cu_or_pq_arch.py:99-127: logits = lvl1[2], lvl2[3], lvl3[4], lvl4[6] */
void mlto_cu_forward(const mlto_model *m, int size, const float *x, int poc, int qp, float logits[15])
{
    const int S = size;
    const size_t big = (size_t)S * S * 32;
    float *nhwc = (float *)malloc((size_t)S * S * 2 * sizeof(float));
    float *a = (float *)malloc(big * sizeof(float)), *b = (float *)malloc(big * sizeof(float));
    float *t1 = (float *)malloc(big * sizeof(float)), *t2 = (float *)malloc(big * sizeof(float));
    for (int p = 0; p < S * S; p++) { nhwc[2 * p] = x[p]; nhwc[2 * p + 1] = x[S * S + p]; }
    conv2d(&m->conv1, nhwc, S, S, a); /* :105 conv1, no BN / ReLU */
    int h = S, off = 0;
    for (int L = 0; L < m->nstage; L++) {
        basic_block(&m->blk[L][0], a, h, b, t1, t2); /* stride-2 block */
        h = (h - 1) / 2 + 1;
        basic_block(&m->blk[L][1], b, h, a, t1, t2);
        if (L >= 1) { /* branch L hangs off layer L (:108-127) */
            gap_fc(m, L - 1, a, h, m->blk[L][1].conv2.cout, poc, qp, logits + off, NULL);
            off += m->fc_out[L - 1];
        }
    }
    free(nhwc); free(a); free(b); free(t1); free(t2);
}

typedef struct {
    const mlto_model *m;
    int n, tid, nthreads, size;
    const int16_t *orgpred;
    const int32_t *pocqp;
    float *logits;
} cu_job;

static void *cu_worker(void *arg)
{
    const cu_job *j = (const cu_job *)arg;
    const size_t plane = (size_t)j->size * j->size;
    float *x = (float *)malloc(2 * plane * sizeof(float));
    for (int i = j->tid; i < j->n; i += j->nthreads) {
        const int16_t *o = j->orgpred + (size_t)i * 2 * plane;
        mlto_cu_stage(j->size, o, j->size, o + plane, j->size, x);
        mlto_cu_forward(j->m, j->size, x, j->pocqp[2 * i], j->pocqp[2 * i + 1], j->logits + (size_t)15 * i);
    }
    free(x);
    return NULL;
}

void mlto_cu_predict_batch(const mlto_model *m, int size, int n, const int16_t *orgpred, const int32_t *pocqp,
                           float *logits, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t th[256];
    cu_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (cu_job){m, n, t, nthreads, size, orgpred, pocqp, logits};
        if (t > 0) pthread_create(&th[t], NULL, cu_worker, &jobs[t]);
    }
    cu_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

void mlto_forward(const mlto_model *m, const float *x, int poc, int qp, float logits[9])
{
    mlto_forward_ex(m, x, poc, qp, logits, NULL, NULL, NULL);
}

int mlto_predict(const mlto_model *m, const int16_t *org, int org_stride, const int16_t *pred,
                 int pred_stride, int poc, int qp, float logits[9])
{
    float *x = (float *)malloc((size_t)2 * MLTO_CTU * MLTO_CTU * sizeof(float));
    mlto_stage(org, org_stride, pred, pred_stride, x);
    mlto_forward(m, x, poc, qp, logits);
    free(x);
    /* EncCu.cpp:913-921 : argmax(1) of the 3rd tuple element; first maximum wins (torch.argmax) */
    int best = 0;
    for (int k = 1; k < 4; k++)
        if (logits[5 + k] > logits[5 + best]) best = k;
    return best;
}

typedef struct {
    const mlto_model *m;
    int n, tid, nthreads;
    const int16_t *orgpred;
    const int32_t *pocqp;
    float *logits;
    int32_t *split;
} batch_job;

static void *batch_worker(void *arg)
{
    const batch_job *j = (const batch_job *)arg;
    const size_t plane = (size_t)MLTO_CTU * MLTO_CTU;
    for (int i = j->tid; i < j->n; i += j->nthreads) {
        const int16_t *o = j->orgpred + (size_t)i * 2 * plane;
        j->split[i] = mlto_predict(j->m, o, MLTO_CTU, o + plane, MLTO_CTU, j->pocqp[2 * i], j->pocqp[2 * i + 1],
                                   j->logits + (size_t)9 * i);
    }
    return NULL;
}

void mlto_predict_batch(const mlto_model *m, int n, const int16_t *orgpred, const int32_t *pocqp,
                        float *logits, int32_t *split, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t th[256];
    batch_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (batch_job){m, n, t, nthreads, orgpred, pocqp, logits, split};
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---- frame-level pre-pass (see mltcnn_oracle.h) ---------------------------------------------------------------- */
int mlto_picture_ctus(int w, int h, int32_t *xy, int cap)
{
    int n = 0;
    /* raster order of EncSlice::encodeCtus (EncSlice.cpp:1529); gate of EncCu.cpp:755 */
    for (int y = 0; y + MLTO_CTU <= h; y += MLTO_CTU)
        for (int x = 0; x + MLTO_CTU <= w; x += MLTO_CTU) {
            if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; }
            n++;
        }
    return n;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void mlto_picture_block_pred(const int16_t *ref, int ref_stride, int w, int h, int size, int x, int y, int mvx, int mvy, int16_t *pred)
{
    /* reading outside the picture returns the nearest border sample: that is what the replicated margins of a
     * reference picture hold (Picture.cpp:1117) */
    for (int r = 0; r < size; r++) {
        const int16_t *row = ref + (size_t)clampi(y + r + mvy, 0, h - 1) * ref_stride;
        for (int c = 0; c < size; c++) pred[r * size + c] = row[clampi(x + c + mvx, 0, w - 1)];
    }
}

void mlto_picture_pred(const int16_t *ref, int ref_stride, int w, int h, int x, int y, int mvx, int mvy, int16_t *pred)
{
    mlto_picture_block_pred(ref, ref_stride, w, h, MLTO_CTU, x, y, mvx, mvy, pred);
}

/* Integer full-search block matching of the CTU at (x, y) (the library's mlt_estimate_picture_mv): candidates dx, dy in
 * [-range, range], cost = sum |org - ref(clamped)|; smallest (cost, preference) wins, preference 0 = zero MV, then raster
 * order of (dy, dx).  Returns the winning cost. */
uint32_t mlto_picture_me(const int16_t *org, int org_stride, const int16_t *ref, int ref_stride, int w, int h, int x, int y, int range,
                         int16_t mv[2])
{
    const int D = 2 * range + 1, centre = range * D + range;
    uint64_t best = ~(uint64_t)0;
    for (int k = 0; k < D * D; k++) {
        const int dy = k / D - range, dx = k % D - range;
        uint32_t cost = 0;
        for (int r = 0; r < MLTO_CTU; r++) {
            const int16_t *o = org + (size_t)(y + r) * org_stride + x;
            const int16_t *q = ref + (size_t)clampi(y + r + dy, 0, h - 1) * ref_stride;
            for (int c = 0; c < MLTO_CTU; c++) {
                const int d = (int)o[c] - (int)q[clampi(x + c + dx, 0, w - 1)];
                cost += (uint32_t)(d < 0 ? -d : d);
            }
        }
        const uint32_t pref = k == centre ? 0u : (k < centre ? (uint32_t)k + 1u : (uint32_t)k);
        const uint64_t key = ((uint64_t)cost << 16) | pref;
        if (key < best) { best = key; mv[0] = (int16_t)dx; mv[1] = (int16_t)dy; }
    }
    return (uint32_t)(best >> 16);
}
