// hook_cu_sim.cpp -- drives the smaller-CU branch of the host-side hook (mlt_hook.h: useCNN with the size mask of
// EncCu.cpp:754 switched on, predictCu == elements()[0] of EncCu.cpp:916-919) over the square CUs of one size of a
// picture, the way the partitioner would hand them to xCompressCU: pointers + strides into the picture buffers.
// Used by tests/test_gpu_cu_parity.py to check the C++ hook against the ctypes binding on the GPU.
//   file: int32 {width, height, stride, poc, qp, cuw}, int16 org[height][stride], int16 pred[height][stride]
//   out : one line per eligible CU (raster order): "x y split split_from_the_picture_pre_pass"
// The fourth column is the frame-level pre-pass (prepassPictureCu + pictureSplitCu) run with the file's pred plane as
// the reference picture and zero MVs -- by construction the same prediction the per-CU calls were given.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mlt_hook.h"

int main(int argc, char **argv)
{
    if (argc != 2) { std::fprintf(stderr, "usage: hook_cu_sim picture.bin (MLT_WEIGHTS_<cuw> and MLT_CU_SIZES must be set)\n"); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[6];
    if (std::fread(hdr, sizeof(int), 6, f) != 6) return 2;
    const int w = hdr[0], h = hdr[1], stride = hdr[2], poc = hdr[3], qp = hdr[4], cuw = hdr[5];
    std::vector<int16_t> org((size_t)h * stride), pred((size_t)h * stride);
    if (std::fread(org.data(), sizeof(int16_t), org.size(), f) != org.size()) return 2;
    if (std::fread(pred.data(), sizeof(int16_t), pred.size(), f) != pred.size()) return 2;
    std::fclose(f);
    const unsigned mask = mlt_hook::cuSizeMaskFromEnv();
    mlt_hook::SplitPredictor &p = mlt_hook::SplitPredictor::instance();
    int seen = 0;
    std::vector<int> perCu;
    for (int y = 0; y < h; y += cuw)
        for (int x = 0; x < w; x += cuw) {
            if (!mlt_hook::useCNN(0, false, cuw, cuw, x, y, w, h, mask)) continue; // partial CUs and masked-off sizes are skipped
            perCu.push_back(p.predictCu(cuw, org.data() + (size_t)y * stride + x, stride, pred.data() + (size_t)y * stride + x, stride, poc, qp));
            seen++;
        }
    if (seen > 0 && !p.prepassPictureCu(cuw, org.data(), stride, pred.data(), stride, w, h, poc, nullptr, qp)) return 3;
    int k = 0;
    for (int y = 0; y < h; y += cuw)
        for (int x = 0; x < w; x += cuw) {
            const bool use = mlt_hook::useCNN(0, false, cuw, cuw, x, y, w, h, mask);
            const int pre = p.pictureSplitCu(cuw, x, y);
            if (seen > 0 && !mlt_hook::useCNN(0, false, cuw, cuw, x, y, w, h, 7u) && pre != -1) return 5; // partial CUs carry no decision
            if (use) std::printf("%d %d %d %d\n", x, y, perCu[(size_t)k++], pre);
        }
    return seen > 0 ? 0 : 4;
}
