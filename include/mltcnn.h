/*
 * mltcnn.h -- C ABI of libmltcnn.so: B200-native (sm_100a) MLT-CNN inter CU-split predictor.
 *
 * Drop-in boundary for the inference block that smu-ivpl/FastInterCU-VVC pastes inline into
 * VTM-11.0's EncCu::xCompressCU (reference: vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:803-926).
 * Each entry point below names the reference lines it replaces.  Plain C types only: no torch,
 * no OpenCV, no C++ exceptions cross this boundary.  There is NO CPU fallback: every call fails
 * with a negative code if the CUDA device / kernels are unavailable, and the host then passes
 * predictedSplitMode = -1 to the unchanged EncModeCtrl::setNewModeList (EncModeCtrl.cpp:110-149),
 * which is exactly the reference's error convention (EncCu.cpp:694,902-905,923-926).
 *
 * Threading: an mlt_ctx is single-threaded / non-reentrant, one per process per GPU, calls are
 * synchronous -- the same contract as the reference hook (one host thread, blocking libtorch calls).
 */
#ifndef MLTCNN_H
#define MLTCNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MLT_API __attribute__((visibility("default")))
#else
#define MLT_API
#endif

#define MLT_ABI_VERSION 1
#define MLT_CTU_SIZE 128 /* the deployed model covers 128x128 luma CTUs only (EncCu.cpp:754) */

/* return codes: 0 = success, negative = failure (never throws) */
enum {
    MLT_OK = 0,
    MLT_E_INVAL = -1,    /* bad argument */
    MLT_E_IO = -2,       /* weight file unreadable */
    MLT_E_FORMAT = -3,   /* weight file is not an MLTW blob for this architecture */
    MLT_E_CUDA = -4,     /* CUDA runtime / kernel error (see mlt_last_error) */
    MLT_E_NOMEM = -5,    /* host or device allocation failed */
    MLT_E_NODEVICE = -6, /* no sm_100 CUDA device */
    MLT_E_BATCH = -7,    /* n exceeds the max_batch the context was created with */
    MLT_E_STATE = -8     /* call sequence error (e.g. no picture begun) */
};

/* split classes = VTM PartSplit numbering (UnitPartitioner.h:56-64) */
enum { MLT_SPLIT_NONE = 0, MLT_SPLIT_QT = 1, MLT_SPLIT_BT_H = 2, MLT_SPLIT_BT_V = 3 };

/* mlt_result.flags bits: per-level split flags the partitioning loop can prune with */
enum {
    MLT_FLAG_L1_SPLIT = 1u << 0, /* level 1 (2 classes): 1 = split */
    MLT_FLAG_L2_QT = 1u << 1,    /* level 2 (3 classes): QT  */
    MLT_FLAG_L2_MTT = 1u << 2,   /* level 2: MTT (BT/TT family) */
    MLT_FLAG_L3_QT = 1u << 3,    /* level 3 (4 classes): QT  */
    MLT_FLAG_L3_BT_H = 1u << 4,  /* level 3: horizontal binary split */
    MLT_FLAG_L3_BT_V = 1u << 5,  /* level 3: vertical binary split */
    MLT_FLAG_CONSISTENT = 1u << 8 /* the three levels agree (NS/NS/NS, split/QT/QT or split/MTT/BT_x) */
};

/* One prediction.  split_l3 is what EncCu.cpp:913-921 computes:
 * argmax(1) of the 3rd output (4 classes) == predictedSplitMode. */
typedef struct mlt_result {
    int32_t split_l3;  /* 0 NS, 1 QT, 2 BT_H, 3 BT_V */
    int32_t split_l2;  /* 0 NS, 1 QT, 2 MTT */
    int32_t split_l1;  /* 0 NS, 1 split */
    uint32_t flags;    /* MLT_FLAG_* */
    float logits[9];   /* raw outputs (lvl1[2], lvl2[3], lvl3[4]) as returned by cnn.forward (EncCu.cpp:909) */
    float probs[9];    /* per-level softmax of the logits */
} mlt_result;

/* One CTU of a batch: pointers + strides in Pel (int16) units, as VTM's AreaBuf{buf,stride}
 * (Buffer.h:94-97) hands them out; poc / qp as EncCu.cpp:806-807. */
typedef struct mlt_ctu_desc {
    const int16_t *org;
    const int16_t *pred;
    int32_t org_stride;
    int32_t pred_stride;
    int32_t poc;
    int32_t qp;
} mlt_ctu_desc;

typedef struct mlt_ctx mlt_ctx;

/* Replaces torch::jit::load(path, device) + cnn.eval() (EncCu.cpp:894-905) -- done ONCE, not per CTU.
 * weights_path: MLTW blob written by `python -m fastintercu_vvc_b200.pack_weights` from the same
 * state_dict ('params' key, optional 'module.' prefixes) model2torchScript.py:23-32 reads.
 * cuda_device: ordinal after CUDA_VISIBLE_DEVICES (the reference hard-codes at::kCUDA = 0, EncCu.cpp:804).
 * max_batch: largest n accepted by the batch calls (device + pinned buffers are sized once, here). */
MLT_API int mlt_create(mlt_ctx **ctx, const char *weights_path, int cuda_device);
MLT_API int mlt_create_ex(mlt_ctx **ctx, const char *weights_path, int cuda_device, int max_batch);
MLT_API void mlt_destroy(mlt_ctx *ctx);

/* The exact drop-in for EncCu.cpp:806-921: stage (cast, absdiff, /1023, clamp), forward, argmax.
 * org/pred: top-left sample of the 128x128 luma block inside the caller's Pel buffers. */
MLT_API int mlt_predict_ctu(mlt_ctx *ctx, const int16_t *org, int org_stride, const int16_t *pred, int pred_stride,
                    int poc, int qp, mlt_result *out);

/* Batched form for all CTUs of a frame / of several independent encodes (north-star item 4). */
MLT_API int mlt_predict_batch(mlt_ctx *ctx, int n, const mlt_ctu_desc *descs, mlt_result *out);

/* Dense host batch: orgpred[n][2][128][128] int16 (plane 0 = org, plane 1 = pred), pocqp[n][2]. */
MLT_API int mlt_predict_batch_dense(mlt_ctx *ctx, int n, const int16_t *orgpred, const int32_t *pocqp, mlt_result *out);

/* Pipelined form of mlt_predict_batch_dense for throughput runs over frames of independent encodes (north-star item 4;
 * no counterpart in the reference, whose hook is one blocking call per CTU): mlt_submit_batch_dense enqueues the H2D
 * copies, kernels and the D2H of the results and returns; mlt_collect blocks for the OLDEST submitted batch and copies
 * its n results out.  At most two batches may be in flight (the second one's H2D runs while the first computes; a second
 * set of input buffers is allocated on first use).  `orgpred` must stay valid and unchanged until that batch is
 * collected (pinned memory recommended); `pocqp` is consumed before the call returns.  Synchronous calls are refused
 * (MLT_E_STATE) while batches are in flight. */
MLT_API int mlt_submit_batch_dense(mlt_ctx *ctx, int n, const int16_t *orgpred, const int32_t *pocqp);
MLT_API int mlt_collect(mlt_ctx *ctx, mlt_result *out, int *n_out);

/* 10-bit packed transport of the same dense batch (no counterpart in the reference, which moves 128 KiB of fp32 per CTU,
 * EncCu.cpp:875): with InternalBitDepth 10 every Pel is in [0, 1023], so a CTU's 2 x 16384 int16 samples travel as a
 * little-endian bit stream of 10-bit fields, MLT_CTU_PACKED10_BYTES = 40960 bytes instead of 65536, and are unpacked on the
 * device in front of the stem kernel.  On multi-GPU hosts the aggregate H2D rate bounds the batch API; this is -37.5 % bytes.
 * mlt_pack10 is the host-side packer (plain C, any thread; count % 4 == 0): it returns the number of samples OUTSIDE
 * [0, 1023] -- such a block cannot be packed (the (uint16_t)-cast semantics of EncCu.cpp:816,827 need all 16 bits) and must
 * go through the int16 entry points.  Results are byte-identical to mlt_predict_batch_dense on the unpacked samples.
 * mlt_submit_batch_packed10 pairs with mlt_collect exactly like mlt_submit_batch_dense. */
#define MLT_CTU_PACKED10_BYTES 40960
MLT_API uint64_t mlt_pack10(const int16_t *src, uint64_t count, uint8_t *dst);
MLT_API int mlt_predict_batch_packed10(mlt_ctx *ctx, int n, const uint8_t *packed, const int32_t *pocqp, mlt_result *out);
MLT_API int mlt_submit_batch_packed10(mlt_ctx *ctx, int n, const uint8_t *packed, const int32_t *pocqp);

/* Device-resident batch on the caller's stream (cudaStream_t passed as void*; NULL = default stream).
 * d_orgpred / d_pocqp / d_out are device pointers; asynchronous w.r.t. the host.  The call uses the context's own
 * activation buffers: the library orders it after the previous device call (whatever its stream) and orders every later
 * call of any entry point after it, so back-to-back calls need no synchronisation by the caller; the caller only has to
 * keep d_orgpred / d_pocqp / d_out alive and untouched until its stream reaches the end of the call.  Refused with
 * MLT_E_STATE while submitted batches are uncollected. */
MLT_API int mlt_predict_batch_device(mlt_ctx *ctx, int n, const int16_t *d_orgpred, const int32_t *d_pocqp,
                             mlt_result *d_out, void *cuda_stream);

/* Per-picture staging (north-star item 1): upload the picture's original luma plane once at the top of
 * EncSlice::encodeCtus (EncSlice.cpp:1479-1528); per-CTU calls then ship only the 32 KiB prediction
 * block.  (x, y) = luma position of the CTU; it must lie fully inside the picture (EncCu.cpp:755). */
MLT_API int mlt_begin_picture(mlt_ctx *ctx, const int16_t *org_luma, int stride, int width, int height, int poc);
MLT_API int mlt_predict_ctu_in_picture(mlt_ctx *ctx, int x, int y, const int16_t *pred, int pred_stride, int qp,
                               mlt_result *out);

/* Frame-level pre-pass (no counterpart in the reference, whose `pred` is the RDO loop's own
 * best AFFINE / MERGE_SKIP prediction, EncCu.cpp:820-830, and therefore serialises the calls CTU by CTU): after
 * mlt_begin_picture, infer EVERY eligible CTU of the picture (128x128, fully inside, EncCu.cpp:755; raster order) in one
 * batch at the top of EncSlice::encodeCtus (EncSlice.cpp:1479), from a neighbour-independent prediction built on the
 * device: pred(x, y) = ref(clamp(x + mvx, 0, w-1), clamp(y + mvy, 0, h-1)) -- integer-sample motion compensation from
 * a reference luma plane (e.g. slice->getRefPic(REF_PIC_LIST_0, 0)->getRecoBuf().Y(), EncSlice.cpp:1562) whose border
 * is extended by sample replication as Picture::extendPicBorder does (Picture.cpp:1117).
 * ref_luma: width x height of the picture begun, ref_stride in samples.  mv: [n][2] (x, y) integer luma-sample MVs per
 * eligible CTU, or NULL = zero MV.  ctu_qp: [n] per-CTU QPs, or NULL = slice_qp for all (currTestMode.qp, EncCu.cpp:807).
 * Returns the number of results written (== mlt_picture_ctu_count) or a negative code.  This changes the encoder's
 * decisions relative to the reference hook (different pred) and needs its own BD-rate study; the per-CTU calls above
 * remain the exact drop-in. */
MLT_API int mlt_picture_ctu_count(const mlt_ctx *ctx);
/* Optional motion for the pre-pass: integer full-search block matching on the device, one MV per eligible CTU --
 * cost(dx, dy) = sum over the 128x128 block of |org - ref(clamp(x + dx), clamp(y + dy))| for dx, dy in [-range, range]
 * (range <= 16), reference borders replicated; the smallest (cost, preference) wins, preference 0 = the zero MV, then
 * raster order of (dy, dx).  mv_out: [n][2] (x, y), cost_out: [n] winning costs or NULL.  The reference plane stays
 * on the device: a following mlt_predict_picture may pass ref_luma = NULL to reuse it (and ref_luma = NULL here reuses
 * the plane of a previous call for the same picture).  Returns n or a negative code. */
MLT_API int mlt_estimate_picture_mv(mlt_ctx *ctx, const int16_t *ref_luma, int ref_stride, int range, int16_t *mv_out,
                            uint32_t *cost_out);
MLT_API int mlt_predict_picture(mlt_ctx *ctx, const int16_t *ref_luma, int ref_stride, const int16_t *mv, const int32_t *ctu_qp,
                        int slice_qp, mlt_result *out, int capacity);

/* Optional: page-lock a long-lived host buffer (VTM allocates a Picture's PelStorage once and keeps it for the whole
 * encode, Picture::create, Picture.cpp:202-213) so that mlt_begin_picture / mlt_predict_picture / the batch calls DMA straight
 * out of it instead of going through the driver's pageable staging.  Plain wrappers, so the host side needs no CUDA
 * headers; unpin before the buffer is freed.  Pinning the same range twice is not an error. */
MLT_API int mlt_pin_host_buffer(mlt_ctx *ctx, const void *ptr, uint64_t bytes);
MLT_API int mlt_unpin_host_buffer(mlt_ctx *ctx, const void *ptr);

/* ---- introspection / test hooks (not needed by the encoder) ---- */
MLT_API const char *mlt_strerror(int rc);
MLT_API const char *mlt_last_error(const mlt_ctx *ctx); /* detail of the last failure ("" if none) */
MLT_API int mlt_abi_version(void);
/* Engine: 0 = tcgen05 implicit-GEMM kernels (product path), 1 = fp32 CUDA-core kernels (GPU-side
 * cross-check used by tests; never selected implicitly). */
MLT_API int mlt_set_engine(mlt_ctx *ctx, int engine);
/* Number of kernels launched by this context so far (bench.py's gpu_launches). */
MLT_API uint64_t mlt_launch_count(const mlt_ctx *ctx);
/* Per-kernel device times of the LAST batch (CUDA events recorded on the stream the kernels were launched
 * on): ms[0] = fused staging+conv1, ms[1..16] = the 16 tcgen05 convs in forward order, ms[17] = head.
 * mlt_get_profile returns the number of entries written (18) or a negative code. */
MLT_API int mlt_set_profiling(mlt_ctx *ctx, int on);
MLT_API int mlt_get_profile(mlt_ctx *ctx, float *ms, int capacity);
/* Bit-exactness probe of the staging arithmetic (EncCu.cpp:810-867) run on the GPU:
 * out = fp32 [n][2][128][128] (channel 0 = org/1023, channel 1 = |org-pred|/1023, clamped). */
MLT_API int mlt_debug_stage(mlt_ctx *ctx, int n, const mlt_ctu_desc *descs, float *out);
/* The prediction blocks the last mlt_predict_picture built on the device: out = int16 [n][128][128] (bit-exactness
 * probe of the pre-pass gather).  Returns n or a negative code. */
MLT_API int mlt_debug_picture_pred(mlt_ctx *ctx, int16_t *out, int capacity);
/* Copy back an intermediate activation of the last batch as fp32 NHWC [n][H][W][C];
 * layer = 0 (conv1 out) .. 16 (layer3.1 out).  Returns the element count, or a negative code. */
MLT_API int64_t mlt_debug_activation(mlt_ctx *ctx, int layer, float *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* MLTCNN_H */
