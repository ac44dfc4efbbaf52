/*
 * mltcnn_cu.h -- C ABI of libmltcnn.so for the SMALLER-CU models (64x64, 32x32, 16x16 luma CUs).
 *
 * smu-ivpl/FastInterCU-VVC trains one `GapBigMltCuORPQ` network per square CU size below the CTU
 * (mlt-cnn-python/codes/models/archs/mlt_cu_or_pq_arch.py:59-130; exported per size by
 * model2torchScript.py:14-22 as MLTORPQ_splitMode_<cuw>.pt).  The hook in VTM's EncCu::xCompressCU is already
 * written for them -- the size test (vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:754, the 64/32/16 terms are
 * commented out), the per-size model file name (:899) and the output selection `elements()[0]` for cuw != 128
 * (:916-919) -- so enabling a size is a one-line change there.  These entry points are the drop-in for that
 * branch, with the same contract as mltcnn.h: plain C types, return codes instead of exceptions, NO CPU
 * fallback (a failing call makes the host pass predictedSplitMode = -1, EncCu.cpp:694,923-926).
 *
 * Staging is the same arithmetic on a cuw x cuh block (EncCu.cpp:810-867): (uint16) cast, cv::absdiff,
 * convertTo(1/1023), clamp.  The network has five stride-2 stages (32, 64, 96, 128, 256 channels) and four heads
 * with 2 / 3 / 4 / 6 classes after layer1..layer4 (mlt_cu_or_pq_arch.py:99-127).
 */
#ifndef MLTCNN_CU_H
#define MLTCNN_CU_H

#include <stdint.h>

#include "mltcnn.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MLT_CU_NLOGITS 15 /* lvl1[2] + lvl2[3] + lvl3[4] + lvl4[6] */

/* One prediction.  split[0] is what EncCu.cpp:916-921 computes for cuw != 128: argmax(1) of the FIRST output. */
typedef struct mlt_cu_result {
    int32_t split[4];             /* argmax of level 1 (2 classes), 2 (3), 3 (4), 4 (6); first maximum wins (torch.argmax) */
    float logits[MLT_CU_NLOGITS]; /* raw outputs as returned by cnn.forward (EncCu.cpp:909) */
    float probs[MLT_CU_NLOGITS];  /* per-level softmax */
} mlt_cu_result;

typedef struct mlt_cu_ctx mlt_cu_ctx;

/* Replaces torch::jit::load(".../MLTORPQ_splitMode_<cuw>.pt") + eval() (EncCu.cpp:894-905), once per CU size.
 * weights_path: MLTW blob written by `python -m fastintercu_vvc_b200.pack_weights --cu <size> model.pth out.mltw`.
 * cu_size: 64, 32 or 16.  max_batch: largest n of the batch calls. */
MLT_API int mlt_cu_create(mlt_cu_ctx **ctx, const char *weights_path, int cuda_device, int cu_size, int max_batch);
MLT_API void mlt_cu_destroy(mlt_cu_ctx *ctx);

/* Drop-in for EncCu.cpp:806-921 at cuw = cuh = cu_size: org / pred point at the CU's top-left luma sample. */
MLT_API int mlt_cu_predict(mlt_cu_ctx *ctx, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride,
                           int poc, int qp, mlt_cu_result *out);
/* All same-size CUs of a CTU / frame / several encodes in one call (descs as in mltcnn.h). */
MLT_API int mlt_cu_predict_batch(mlt_cu_ctx *ctx, int n, const mlt_ctu_desc *descs, mlt_cu_result *out);
/* Dense host batch: orgpred[n][2][size][size] int16 (plane 0 = org, 1 = pred), pocqp[n][2]. */
MLT_API int mlt_cu_predict_batch_dense(mlt_cu_ctx *ctx, int n, const int16_t *orgpred, const int32_t *pocqp, mlt_cu_result *out);
/* Pipelined form of mlt_cu_predict_batch_dense (same contract as mlt_submit_batch_dense / mlt_collect in mltcnn.h): at
 * most two batches in flight, the H2D copy of batch k + 1 runs under batch k's kernels; `orgpred` must stay valid and
 * unchanged until that batch is collected (pinned memory recommended), `pocqp` is consumed before the call returns;
 * synchronous calls are refused (MLT_E_STATE) while batches are in flight. */
MLT_API int mlt_cu_submit_batch_dense(mlt_cu_ctx *ctx, int n, const int16_t *orgpred, const int32_t *pocqp);
MLT_API int mlt_cu_collect(mlt_cu_ctx *ctx, mlt_cu_result *out, int *n_out);
/* Frame-level pre-pass for one CU size (the CTU form and its rationale are in mltcnn.h,
 * mlt_predict_picture): every cu_size x cu_size block of the picture's raster that lies fully inside the picture (the
 * gate of EncCu.cpp:755 at cuw = cuh = cu_size; raster order, width / cu_size blocks per row) is inferred in ONE batch.
 * Block i: org = the picture's original luma block, pred = integer-sample motion compensation out of `ref_luma` with
 * replicated borders (Picture.cpp:1117) by mv[i] (x, y) -- mv = NULL: zero MV.  Both planes are width x height; they
 * are gathered into the dense batch on the device.  Returns the number of results written
 * (== mlt_cu_picture_cu_count) or a negative code.  Changes the encoder's decisions relative to the reference hook
 * (different pred); the per-CU calls above remain the exact drop-in. */
MLT_API int mlt_cu_picture_cu_count(int cu_size, int width, int height);
MLT_API int mlt_cu_predict_picture(mlt_cu_ctx *ctx, const int16_t *org_luma, int org_stride, const int16_t *ref_luma, int ref_stride,
                                   int width, int height, int poc, const int16_t *mv, int qp, mlt_cu_result *out, int capacity);
/* Device-resident batch on the caller's stream (cudaStream_t as void*); asynchronous w.r.t. the host.  Ordered by the
 * library after the previous device call and before every later call of any entry point (see mlt_predict_batch_device);
 * MLT_E_STATE while submitted batches are uncollected. */
MLT_API int mlt_cu_predict_batch_device(mlt_cu_ctx *ctx, int n, const int16_t *d_orgpred, const int32_t *d_pocqp,
                                        mlt_cu_result *d_out, void *cuda_stream);

/* ---- introspection / test hooks ---- */
MLT_API const char *mlt_cu_last_error(const mlt_cu_ctx *ctx);
MLT_API int mlt_cu_size(const mlt_cu_ctx *ctx);
/* Kernel table of conv `layer` (0..19) of the cu_size network, no GPU needed: info = {cin, cout, stride as executed,
 * hout, extra-operand channels, parity-planar output, images per tile, flat tile, weight-slab channels G, GX}.
 * The weight packer's layout must agree with it (tests/test_host_cpu.py). */
MLT_API int mlt_cu_layer_info(int cu_size, int layer, int32_t info[10]);
MLT_API uint64_t mlt_cu_launch_count(const mlt_cu_ctx *ctx);
/* Copy back an intermediate activation of the last (single) batch as fp32 NHWC [n][H][H][C];
 * layer = 0 (conv1 out) .. 20 (layer4.1 out).  Returns the element count, or a negative code. */
MLT_API int64_t mlt_cu_debug_activation(mlt_cu_ctx *ctx, int layer, float *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* MLTCNN_CU_H */
