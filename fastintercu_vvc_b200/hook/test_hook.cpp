// test_hook.cpp -- CPU-only checks of the host-side hook mirror (run by tests/test_hook.py).
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <vector>

#include "mlt_hook.h"

using namespace mlt_hook;

static ModeListState ra_stack_128()
{
    // RA cfg, 128x128 inter CTU after MERGE_SKIP was popped: bottom -> top
    ModeListState s;
    for (EncTestModeType t : {ETM_SPLIT_QT, ETM_SPLIT_BT_V, ETM_SPLIT_BT_H, ETM_POST_DONT_SPLIT, ETM_INTRA, ETM_INTER_ME})
        s.testModes.push_back({t, 37});
    return s;
}

int main()
{
    // ---- gate (EncCu.cpp:752-756): eligible CTUs per inter frame 416x240 -> 3, 1920x1080 -> 120, 3840x2160 -> 480
    struct { int w, h, want; } pics[] = {{416, 240, 3}, {1920, 1080, 120}, {3840, 2160, 480}};
    for (auto &p : pics) {
        int n = 0;
        for (int y = 0; y < p.h; y += 128)
            for (int x = 0; x < p.w; x += 128) n += useCNN(0, false, 128, 128, x, y, p.w, p.h);
        assert(n == p.want);
    }
    assert(!useCNN(1, false, 128, 128, 0, 0, 416, 240)); // chroma tree
    assert(!useCNN(0, true, 128, 128, 0, 0, 416, 240));  // I slice
    assert(!useCNN(0, false, 64, 64, 0, 0, 416, 240));   // smaller CUs are gated off (EncCu.cpp:754)
    assert(!useCNN(0, false, 128, 64, 0, 0, 416, 240));
    // the sizes the reference left commented out at EncCu.cpp:754, switched on through the mask
    assert(!useCNN(0, false, 64, 64, 0, 0, 416, 240, 0u) && useCNN(0, false, 128, 128, 0, 0, 416, 240, 0u));
    assert(useCNN(0, false, 64, 64, 64, 64, 416, 240, 1u) && !useCNN(0, false, 32, 32, 0, 0, 416, 240, 1u));
    assert(useCNN(0, false, 32, 32, 384, 208, 416, 240, 2u) && !useCNN(0, false, 32, 32, 400, 208, 416, 240, 2u));
    assert(useCNN(0, false, 16, 16, 400, 224, 416, 240, 7u) && !useCNN(0, false, 16, 8, 0, 0, 416, 240, 7u));
    assert(!useCNN(0, false, 8, 8, 0, 0, 416, 240, 7u) && !useCNN(0, true, 64, 64, 0, 0, 416, 240, 7u));
    {
        int n64 = 0; // 1920x1080: 64-px CUs fully inside the picture = 30 x 16
        for (int y = 0; y < 1080; y += 64)
            for (int x = 0; x < 1920; x += 64) n64 += useCNN(0, false, 64, 64, x, y, 1920, 1080, 1u);
        assert(n64 == 30 * 16);
        setenv("MLT_CU_SIZES", "64,16", 1);
        assert(cuSizeMaskFromEnv() == 5u);
    }

    // ---- consumer semantics
    for (int pred = 1; pred <= 3; pred++) {
        ModeListState s = ra_stack_128();
        setNewModeList(s, pred, 37, true);
        assert(s.testModes.size() == 2 && s.testModes[0].type == EncTestModeType(pred + 6) && s.testModes[0].qp == 37);
        assert(s.testModes[1].type == ETM_POST_DONT_SPLIT);
        assert(s.didHorzSplit == (pred == 2) && s.didVertSplit == (pred == 3));
    }
    {
        ModeListState s = ra_stack_128(); // predicted BT_H not allowed here -> QT
        s.didHorzSplit = true;
        setNewModeList(s, 2, 30, false);
        assert(s.testModes.size() == 2 && s.testModes[0].type == ETM_SPLIT_QT && !s.didHorzSplit);
    }
    {
        ModeListState s = ra_stack_128(); // no split: splits removed, the rest kept in order
        setNewModeList(s, 0, 37, true);
        assert(s.testModes.size() == 3 && s.testModes[0].type == ETM_POST_DONT_SPLIT && s.testModes[2].type == ETM_INTER_ME);
    }
    {
        ModeListState s = ra_stack_128(); // failure: untouched
        setNewModeList(s, -1, 37, true);
        assert(s.untouched && s.testModes.size() == 6);
    }

    // ---- predictor failure convention: no usable GPU / weights here -> -1, never throws, never falls back
    setenv("MLT_WEIGHTS", "/nonexistent/weights.mltw", 1);
    char trace_path[] = "/tmp/mlt_hook_trace_XXXXXX", dump_path[] = "/tmp/mlt_hook_dump_XXXXXX";
    close(mkstemp(trace_path));
    close(mkstemp(dump_path));
    setenv("MLT_TRACE", trace_path, 1);       // read once, by the constructor
    setenv("MLT_DUMP_INPUTS", dump_path, 1);
    SplitPredictor &p = SplitPredictor::instance();
    std::vector<int16_t> blk(128 * 128, 512);
    const int r = p.predict(blk.data(), 128, blk.data(), 128, 1, 32);
    if (!p.enabled()) assert(r == -1);
    setenv("MLT_WEIGHTS_64", "/nonexistent/cu64.mltw", 1);
    assert(p.predictCu(64, blk.data(), 128, blk.data(), 128, 1, 32) == -1); // no weights / no GPU -> -1, full RDO
    assert(p.predictCu(48, blk.data(), 128, blk.data(), 128, 1, 32) == -1);
    // frame-level pre-pass: without a usable predictor nothing is ever reported as a decision
    assert(p.pictureSplit(0, 0) == -1 && p.pictureSplitCu(64, 0, 0) == -1 && p.pictureSplitCu(48, 0, 0) == -1);
    if (!p.enabled()) {
        assert(!p.beginPicture(blk.data(), 128, 128, 128, 1));
        assert(!p.prepassPicture(blk.data(), 128, nullptr, 32) && !p.prepassPicture(blk.data(), 128, nullptr, 32, 8));
        assert(!p.prepassPictureCu(64, blk.data(), 128, blk.data(), 128, 128, 128, 1, nullptr, 32));
        assert(p.pictureSplit(0, 0) == -1 && p.pictureSplitCu(64, 0, 0) == -1);
    }
    // predictAt = the one call the VTM patch makes (integration/apply_vtm_patch.py): same -1 convention, one trace line and one
    // input record per 128x128 call (poc, qp, strided org block, strided pred block)
    if (!p.enabled()) {
        std::vector<int16_t> pic(160 * 200);
        for (size_t i = 0; i < pic.size(); i++) pic[i] = (int16_t)(i % 1000);
        assert(p.predictAt(128, 128, 256, pic.data() + 3, 200, blk.data(), 128, 7, 35) == -1);
        assert(p.predictAt(64, 64, 0, pic.data(), 200, blk.data(), 128, 7, 35) == -1); // no trace record for the inputs of smaller CUs
        FILE *tf = std::fopen(trace_path, "r");
        int a, b, c, d, e, lines = 0;
        while (tf && std::fscanf(tf, "%d %d %d %d %d", &a, &b, &c, &d, &e) == 5) {
            if (lines == 0) assert(a == 7 && b == 128 && c == 256 && d == 35 && e == -1);
            lines++;
        }
        if (tf) std::fclose(tf);
        assert(lines == 2);
        FILE *df = std::fopen(dump_path, "rb");
        assert(df);
        int32_t hdr[2];
        std::vector<int16_t> rec(2 * 128 * 128);
        assert(std::fread(hdr, 4, 2, df) == 2 && hdr[0] == 7 && hdr[1] == 35);
        assert(std::fread(rec.data(), 2, rec.size(), df) == rec.size());
        assert(rec[0] == pic[3] && rec[128] == pic[200 + 3] && rec[127 * 128 + 127] == pic[127 * 200 + 127 + 3]); // stride honoured
        assert(rec[128 * 128] == 512);
        assert(std::fgetc(df) == EOF); // exactly one record
        std::fclose(df);
    }
    std::remove(trace_path);
    std::remove(dump_path);
    setenv("MLT_PREPASS", "1", 1);
    setenv("MLT_PREPASS_RANGE", "40", 1);
    assert(SplitPredictor::prepassFromEnv() && SplitPredictor::prepassRangeFromEnv() == 16);
    setenv("MLT_PREPASS", "0", 1);
    unsetenv("MLT_PREPASS_RANGE");
    assert(!SplitPredictor::prepassFromEnv() && SplitPredictor::prepassRangeFromEnv() == 0);
    std::printf("test_hook: OK (predictor %s, predict -> %d)\n", p.enabled() ? "enabled" : "disabled", r);
    return 0;
}
