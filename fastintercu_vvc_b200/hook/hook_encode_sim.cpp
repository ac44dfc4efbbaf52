// hook_encode_sim.cpp -- drives the host-side hook (mlt_hook.h) the way VTM's CTU loop would (EncSlice.cpp:1529 ->
// EncCu.cpp:746-756, 800-928) on a picture read from a file: raster scan of 128x128 CTUs, useCNN gate, one predict() per
// eligible CTU with pointers + strides into the picture buffer, then the same through the per-picture staging path.
// Used by tests/test_gpu_parity.py to check the C++ hook against the ctypes binding on the GPU.
//   file: int32 {width, height, stride, poc, n_ctu}, int16 org[height][stride], then per eligible CTU (raster order):
//         int32 qp, int16 pred[128][128]
//   out : one line per eligible CTU: "x y split split_in_picture"
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mlt_hook.h"

int main(int argc, char **argv)
{
    if (argc != 2) { std::fprintf(stderr, "usage: hook_encode_sim picture.bin (MLT_WEIGHTS must be set)\n"); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[5];
    if (std::fread(hdr, sizeof(int), 5, f) != 5) return 2;
    const int w = hdr[0], h = hdr[1], stride = hdr[2], poc = hdr[3], nctu = hdr[4];
    std::vector<int16_t> org((size_t)h * stride);
    if (std::fread(org.data(), sizeof(int16_t), org.size(), f) != org.size()) return 2;
    mlt_hook::SplitPredictor &p = mlt_hook::SplitPredictor::instance();
    if (!p.enabled()) { std::fprintf(stderr, "predictor disabled\n"); return 3; }
    if (!p.beginPicture(org.data(), stride, w, h, poc)) return 3;
    std::vector<int16_t> pred(128 * 128);
    int seen = 0;
    for (int y = 0; y < h; y += 128)
        for (int x = 0; x < w; x += 128) {
            if (!mlt_hook::useCNN(0, false, 128, 128, x, y, w, h)) continue; // partial CTUs are skipped like EncCu.cpp:755
            int qp;
            if (std::fread(&qp, sizeof(int), 1, f) != 1 || std::fread(pred.data(), sizeof(int16_t), pred.size(), f) != pred.size()) return 2;
            const int a = p.predict(org.data() + (size_t)y * stride + x, stride, pred.data(), 128, poc, qp);
            const int b = p.predictInPicture(x, y, pred.data(), 128, qp);
            std::printf("%d %d %d %d\n", x, y, a, b);
            seen++;
        }
    std::fclose(f);
    return seen == nctu ? 0 : 4;
}
