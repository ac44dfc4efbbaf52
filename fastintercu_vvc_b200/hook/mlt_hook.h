// mlt_hook.h -- host-side (C++14, no CUDA / torch / OpenCV headers) mirror of the reference's MLT-CNN hook, written
// against the C ABI of libmltcnn.so (include/mltcnn.h).  This is the code a VTM maintainer drops into
// EncCu::xCompressCU in place of the inline libtorch block (EncCu.cpp:803-926); see INTEGRATION.md.
//
//   reference                                   here
//   ------------------------------------------  -----------------------------------------------------------
//   useCNN gate            EncCu.cpp:746-756     mlt_hook::useCNN(...)
//   at::kCUDA / model path EncCu.cpp:804,899     env MLT_DEVICE / MLT_WEIGHTS (MLT_DISABLE=1 -> anchor run)
//   jit::load per CTU      EncCu.cpp:894-905     SplitPredictor::instance() loads the MLTW blob ONCE
//   staging + forward      EncCu.cpp:806-921     SplitPredictor::predict(org, stride, pred, stride, poc, qp)
//   (none: pred comes from RDO, :820-830)        SplitPredictor::prepassPicture(ref, ...) + pictureSplit(x, y): whole
//                                                frame in one batch before the CTU loop (env MLT_PREPASS=1)
//   error convention       EncCu.cpp:694,923     any failure -> "error\n" on stderr, returns -1 (full RDO)
//   consumer               EncModeCtrl.cpp:110   unchanged in VTM; restated below only so tests can pin it
#pragma once
#include <cstdint>
#include <vector>

struct mlt_ctx;
struct mlt_cu_ctx;

namespace mlt_hook {

// EncCu.cpp:752-756 : luma tree, non-I slice, 128x128 CU, fully inside the picture.
bool useCNN(int chType, bool isIntraSlice, int cuw, int cuh, int cux, int cuy, int picWidth, int picHeight);
// The same gate with the square sizes the reference left commented out at EncCu.cpp:754 (64x64, 32x32, 16x16) switched
// on through `sizeMask` (bit 0: 64, bit 1: 32, bit 2: 16; 0 == the reference as shipped).  Env MLT_CU_SIZES="64,32,16".
bool useCNN(int chType, bool isIntraSlice, int cuw, int cuh, int cux, int cuy, int picWidth, int picHeight, unsigned sizeMask);
unsigned cuSizeMaskFromEnv();

class SplitPredictor {
public:
    // Process-wide instance (the reference hook is single-threaded: ENABLE_SPLIT_PARALLELISM 0, TypeDef.h:113).
    static SplitPredictor &instance();

    bool enabled() const { return m_ctx != nullptr; } // false: MLT_DISABLE=1, or creation failed (message on stderr)

    // == predictedSplitMode of EncCu.cpp:913-921: 0 NS, 1 QT, 2 BT_H, 3 BT_V; -1 on failure (caller passes it on
    // to setNewModeList, which then leaves the mode stack untouched, EncModeCtrl.cpp:147-148).
    int predict(const int16_t *org, int orgStride, const int16_t *pred, int predStride, int poc, int qp);

    // The one call the VTM patch makes per eligible CU (integration/apply_vtm_patch.py): dispatches on what was prepared for the
    // current picture -- the pre-pass table (pictureSplit), the staged org plane (predictInPicture), or nothing (predict /
    // predictCu for cuw < 128) -- then appends "poc x y qp split" to $MLT_TRACE, the call's inputs to $MLT_DUMP_INPUTS
    // (records {int32 poc, int32 qp, int16 org[cuw*cuw], int16 pred[cuw*cuw]}, 128x128 calls only) and adds the call's wall
    // time to the totals printed at exit when MLT_STATS=1.  Same return convention as predict().
    int predictAt(int cuw, int cux, int cuy, const int16_t *org, int orgStride, const int16_t *pred, int predStride, int poc, int qp);
    static bool pictureStagingFromEnv(); // env MLT_PICTURE_STAGING=1: EncSlice uploads the org plane once per picture

    // Optional per-picture staging from EncSlice::encodeCtus (EncSlice.cpp:1479-1528): upload the original luma
    // once, then only the 32 KiB prediction block per CTU.
    bool beginPicture(const int16_t *orgLuma, int stride, int width, int height, int poc);
    int predictInPicture(int cux, int cuy, const int16_t *pred, int predStride, int qp);

    // Page-lock a long-lived picture buffer once (VTM keeps a Picture's PelStorage for the whole encode, Picture.cpp:202-213) so
    // that beginPicture / prepassPicture DMA straight out of it; unpin before the buffer is freed.  false if disabled / failed.
    bool pinHostBuffer(const void *ptr, uint64_t bytes);
    bool unpinHostBuffer(const void *ptr);

    // Frame-level pre-pass, called once per inter picture from EncSlice::encodeCtus after
    // beginPicture: every eligible CTU (gate of EncCu.cpp:755) is inferred in ONE batch from a neighbour-independent
    // prediction -- integer-MV motion compensation out of `refLuma` (border-replicated, Picture.cpp:1117); `mv` = [n][2]
    // (x, y) per eligible CTU in raster order or nullptr for zero MV.  xCompressCU then reads pictureSplit(cux, cuy)
    // instead of calling predict(): -1 when the CTU is not eligible / no pre-pass ran for this picture / it failed.
    // searchRange > 0 with mv == nullptr: the MVs come from the library's integer full search (mlt_estimate_picture_mv,
    // +-searchRange samples, <= 16; env MLT_PREPASS_RANGE), the reference plane is uploaded once for both steps.
    bool prepassPicture(const int16_t *refLuma, int refStride, const int16_t *mv, int sliceQp, int searchRange = 0);
    static int prepassRangeFromEnv(); // env MLT_PREPASS_RANGE (0 = zero MV)
    int pictureSplit(int cux, int cuy) const;
    static bool prepassFromEnv(); // env MLT_PREPASS=1

    // Smaller CUs (cuw = 64 / 32 / 16): the hook's `elements()[0]` branch (EncCu.cpp:916-919) == argmax of the FIRST head
    // of the per-size model MLTORPQ_splitMode_<cuw> (EncCu.cpp:899); weights from env MLT_WEIGHTS_<cuw>, loaded once on
    // first use.  -1 on failure or when that size has no weights.
    int predictCu(int cuw, const int16_t *org, int orgStride, const int16_t *pred, int predStride, int poc, int qp);

    // The pre-pass for one smaller CU size (cuw = 64 / 32 / 16): every cuw x cuw block of the size's raster inside the picture
    // in one batch (mlt_cu_predict_picture); pictureSplitCu reads the level-1 argmax (what predictCu returns) of the block
    // whose top-left corner is (cux, cuy), or -1.  beginPicture() forgets these tables as well.
    bool prepassPictureCu(int cuw, const int16_t *orgLuma, int orgStride, const int16_t *refLuma, int refStride, int width, int height,
                          int poc, const int16_t *mv, int sliceQp);
    int pictureSplitCu(int cuw, int cux, int cuy) const;

    ~SplitPredictor();
    SplitPredictor(const SplitPredictor &) = delete;
    SplitPredictor &operator=(const SplitPredictor &) = delete;

private:
    SplitPredictor();
    mlt_ctx *m_ctx = nullptr;
    int m_ctxCap = 512; // batch capacity of m_ctx (eligible CTUs of the largest picture seen so far)
    mlt_cu_ctx *m_cu[3] = {nullptr, nullptr, nullptr}; // 64, 32, 16
    bool m_cuTried[3] = {false, false, false};
    int m_cuCap[3] = {0, 0, 0};
    bool m_disabled = false;
    int m_picPoc = -1;          // poc of the picture staged by beginPicture, -1: none
    bool m_picStaged = false;
    const char *m_tracePath = nullptr, *m_dumpPath = nullptr;
    bool m_stats = false;
    uint64_t m_calls = 0;
    double m_seconds = 0;
    std::vector<int> m_picSplit; // pre-pass decisions of the current picture, eligible CTUs in raster order
    int m_picCols = 0, m_picRows = 0, m_picW = 0, m_picH = 0;
    std::vector<int> m_cuSplit[3]; // per size: pre-pass decisions of the current picture, raster order
    int m_cuCols[3] = {0, 0, 0}, m_cuRows[3] = {0, 0, 0};
    mlt_cu_ctx *cuContext(int idx, int cuw, int minBatch);
};

// ---- restatement of the consumer's semantics (EncModeCtrl.cpp:95-149), used by tests only ------------------
enum EncTestModeType { // numbering of EncModeCtrl.h:56-78 with REUSE_CU_RESULTS
    ETM_HASH_INTER, ETM_MERGE_SKIP, ETM_INTER_ME, ETM_AFFINE, ETM_MERGE_GEO, ETM_INTRA, ETM_PALETTE,
    ETM_SPLIT_QT, ETM_SPLIT_BT_H, ETM_SPLIT_BT_V, ETM_SPLIT_TT_H, ETM_SPLIT_TT_V, ETM_POST_DONT_SPLIT,
    ETM_RECO_CACHED, ETM_TRIGGER_IMV_LIST, ETM_IBC, ETM_IBC_MERGE, ETM_INVALID
};
struct TestMode {
    EncTestModeType type;
    int qp;
};
struct ModeListState {
    std::vector<TestMode> testModes; // back() is tested next (EncModeCtrl.cpp:95-107)
    bool didHorzSplit = false, didVertSplit = false;
    bool untouched = true;
};
// canSplit: result of partitioner.canSplit(PartSplit(predictedSplitMode), cs) (UnitPartitioner.cpp:458-475)
void setNewModeList(ModeListState &st, int predictedSplitMode, int qp, bool canSplit);

} // namespace mlt_hook
