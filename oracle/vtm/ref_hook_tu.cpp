// ref_hook_tu.cpp -- CPU ORACLE / BASELINE (test infrastructure, NOT product code).
// Runs the REFERENCE'S OWN STATEMENTS for the path -- vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:803-926, extracted at
// build time into oracle/_ref/obj/gen/hook_block.inc by oracle/vtm/Makefile (never committed) -- outside the encoder:
// this file only supplies the few names that block reads from EncCu::xCompressCU's scope (bestCS, currTestMode, cuw, cuh,
// predictedSplitMode, xMalloc) and a timing loop.  Token edits applied to the block (integration/apply_vtm_patch.py
// --hook-block): at::kCUDA -> at::kCPU for the CPU binary (:804), model directory from $MLT_REF_MODEL_DIR (:899), and the
// module kept across calls (:894) unless MLT_REF_LOAD_PER_CALL=1 restores the per-CTU torch::jit::load of the hook as written.
// OpenCV comes from the shim (oracle/vtm/opencv_shim).
//   usage: ref_hook_tu ctus.bin seconds threads        (model: $MLT_REF_MODEL_DIR/MLTORPQ_splitMode_128.pt)
//   ctus.bin: int32 n, then n x { int32 poc, int32 qp, int16 org[128*128], int16 pred[128*128] }
//   stdout  : "split <i> <predictedSplitMode>" for the first 64 CTUs, then "<ctus> <seconds> <threads>"
#include <ATen/Parallel.h>
#include <torch/script.h>

#include <opencv2/opencv.hpp>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>
using namespace std;

#define xMalloc(type, len) ((type *)malloc(sizeof(type) * (len)))

namespace {
struct PlaneStub { short *buf; int stride; };
struct BufStub { PlaneStub y; PlaneStub &Y() { return y; } };
struct SliceStub { int poc; int getPOC() const { return poc; } };
struct CodingStructureStub {
    SliceStub *slice;
    BufStub org, pred;
    BufStub &getOrgBuf() { return org; }
    BufStub &getPredBuf() { return pred; }
};
struct TestModeStub { int qp; };

int referenceHook(CodingStructureStub *bestCS, TestModeStub currTestMode, int cuw, int cuh)
{
    int predictedSplitMode = -1; // EncCu.cpp:694
    {
#include "hook_block.inc"
    }
    return predictedSplitMode;
}
} // namespace

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: ref_hook_tu ctus.bin seconds threads\n"); return 2; }
    const double budget = atof(argv[2]);
    const int threads = atoi(argv[3]);
    if (threads > 0) at::set_num_threads(threads);
    FILE *f = fopen(argv[1], "rb");
    int32_t n = 0;
    if (!f || fread(&n, 4, 1, f) != 1 || n <= 0) return 2;
    constexpr int S = 128;
    vector<int32_t> poc(n), qp(n);
    vector<int16_t> org((size_t)n * S * S), pred((size_t)n * S * S);
    for (int i = 0; i < n; i++) {
        if (fread(&poc[i], 4, 1, f) != 1 || fread(&qp[i], 4, 1, f) != 1) return 2;
        if (fread(&org[(size_t)i * S * S], 2, S * S, f) != (size_t)S * S) return 2;
        if (fread(&pred[(size_t)i * S * S], 2, S * S, f) != (size_t)S * S) return 2;
    }
    fclose(f);
    long done = 0;
    double dt = 0;
    for (int warm = 1; warm >= 0; warm--) {
        done = 0;
        const auto ts = chrono::steady_clock::now();
        for (;;) {
            const int i = (int)(done % n);
            SliceStub sl{poc[i]};
            CodingStructureStub cs{&sl, {{&org[(size_t)i * S * S], S}}, {{&pred[(size_t)i * S * S], S}}};
            const int split = referenceHook(&cs, TestModeStub{qp[i]}, S, S);
            if (!warm && done < 64) printf("split %ld %d\n", done, split);
            done++;
            dt = chrono::duration<double>(chrono::steady_clock::now() - ts).count();
            if (warm ? done >= 3 : (dt >= budget && done >= min<long>(n, 64))) break;
        }
    }
    printf("%ld %.6f %d\n", done, dt, at::get_num_threads());
    return 0;
}
