/*
 * mltcnn_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the MLT-CNN inter CU-split hot path of
 * smu-ivpl/FastInterCU-VVC, written from the reference's behaviour:
 *   - staging / normalisation : vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:810-877
 *   - network forward         : mlt-cnn-python/codes/models/archs/mlt_ctu_or_pq_arch.py:32-57,239-299
 *   - select + argmax         : vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:912-921
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this.  The product path (libmltcnn.so) never
 * does; it fails loudly when the CUDA library is missing.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 4).  This oracle is pinned against outputs of the reference's own
 * Python architecture file, imported from /root/reference by
 * tools/gen_golden.py, committed under tests/golden/ (see DESIGN.md).
 */
#ifndef MLTCNN_ORACLE_H
#define MLTCNN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLTO_CTU 128

typedef struct mlto_model mlto_model; /* raw (un-folded) fp32 parameters */

/* Load the raw "MLTR" parameter blob written by oracle/weights_io.py
 * (same tensors and key order as the reference state_dict,
 * model2torchScript.py:23-32).  Returns NULL on error. */
mlto_model *mlto_load(const char *path);
void mlto_free(mlto_model *m);

/* EncCu.cpp:810-867 : int16 org/pred (strided) -> fp32 [2][128][128]
 * channel 0 = org/1023, channel 1 = |org-pred|/1023, both clamped to [0,1]. */
void mlto_stage(const int16_t *org, int org_stride, const int16_t *pred, int pred_stride,
                float *x /* [2*128*128] NCHW */);

/* mlt_ctu_or_pq_arch.py:273-299 : logits[0..1]=lvl1, [2..4]=lvl2, [5..8]=lvl3 */
void mlto_forward(const mlto_model *m, const float *x, int poc, int qp, float logits[9]);

/* Same, but also returns the three global-average-pooled feature vectors
 * (64, 128, 256 floats) for layer-level debugging of the CUDA path. */
void mlto_forward_ex(const mlto_model *m, const float *x, int poc, int qp, float logits[9],
                     float *gap1, float *gap2, float *gap3);

/* Whole hook: stage + forward + argmax of the 3rd head (EncCu.cpp:912-921).
 * Returns predictedSplitMode in {0 NS, 1 QT, 2 BT_H, 3 BT_V}. */
int mlto_predict(const mlto_model *m, const int16_t *org, int org_stride, const int16_t *pred,
                 int pred_stride, int poc, int qp, float logits[9]);

/* Batch helper: CTU i is handled by pthread (i mod nthreads).
 * orgpred: [n][2][128][128] int16 dense, pocqp: [n][2] int32, logits: [n][9], split: [n]. */
void mlto_predict_batch(const mlto_model *m, int n, const int16_t *orgpred, const int32_t *pocqp,
                        float *logits, int32_t *split, int nthreads);

/* ---- smaller-CU model `GapBigMltCuORPQ` (64 / 32 / 16 px; mlt_cu_or_pq_arch.py:59-130): the hook's cuw != 128 branch
 * (gate EncCu.cpp:754, model file per size :899, elements()[0] :916-919).  The MLTR blob header carries arch = size. */
void mlto_cu_stage(int size, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride,
                   float *x /* [2*size*size] NCHW */);
/* logits[0..1]=lvl1, [2..4]=lvl2, [5..8]=lvl3, [9..14]=lvl4 */
void mlto_cu_forward(const mlto_model *m, int size, const float *x, int poc, int qp, float logits[15]);
/* orgpred: [n][2][size][size] int16 dense, pocqp: [n][2], logits: [n][15] */
void mlto_cu_predict_batch(const mlto_model *m, int size, int n, const int16_t *orgpred, const int32_t *pocqp,
                           float *logits, int nthreads);

/* ---- frame-level pre-pass (the library's mlt_predict_picture).  The reference has no such
 * pass; what is restated here are the two reference behaviours it is built from: the eligibility gate of EncCu.cpp:755
 * (a 128x128 CTU lying fully inside the picture, raster order as EncSlice::encodeCtus walks them, EncSlice.cpp:1529) and
 * integer-sample motion compensation from a reference picture whose border is extended by sample replication
 * (Picture::extendPicBorder, Picture.cpp:1117). */
/* xy: [cap][2] luma positions of the eligible CTUs; returns their number */
int mlto_picture_ctus(int w, int h, int32_t *xy, int cap);
/* pred[r][c] = ref[clamp(y + r + mvy, 0, h-1)][clamp(x + c + mvx, 0, w-1)], r, c in [0, 128) */
void mlto_picture_pred(const int16_t *ref, int ref_stride, int w, int h, int x, int y, int mvx, int mvy,
                       int16_t *pred /* [128*128] dense */);
/* the same for a size x size block (the smaller-CU pre-pass, mlt_cu_predict_picture); pred = [size*size] dense */
void mlto_picture_block_pred(const int16_t *ref, int ref_stride, int w, int h, int size, int x, int y, int mvx, int mvy,
                             int16_t *pred);

/* integer full-search block matching of the CTU at (x, y), see mltcnn_oracle.c; mv = (x, y); returns the winning cost */
uint32_t mlto_picture_me(const int16_t *org, int org_stride, const int16_t *ref, int ref_stride, int w, int h, int x, int y,
                         int range, int16_t mv[2]);

#ifdef __cplusplus
}
#endif
#endif
