// hook_prepass_sim.cpp -- drives the frame-level pre-pass of the host-side hook (mlt_hook.h) the way a patched
// EncSlice::encodeCtus / EncCu::xCompressCU pair would: per picture beginPicture(org) +
// prepassPicture(ref, mv, sliceQp) once, then the CTU raster reads pictureSplit(x, y) under the unchanged useCNN gate.
// Used by tests/test_gpu_parity.py to check the C++ path against the per-CTU drop-in call on the GPU.
//   file: int32 {width, height, stride, poc, slice_qp, has_mv}, int16 org[height][stride], int16 ref[height][stride],
//         then if has_mv == 1: int16 mv[n][2] for the n eligible CTUs (raster order); has_mv < 0: MVs from the library's
//         block matching with search range -has_mv
//   out : one line per CTU of the raster (partial ones included): "x y eligible split"
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mlt_hook.h"

int main(int argc, char **argv)
{
    if (argc != 2) { std::fprintf(stderr, "usage: hook_prepass_sim picture.bin (MLT_WEIGHTS must be set)\n"); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[6];
    if (std::fread(hdr, sizeof(int), 6, f) != 6) return 2;
    const int w = hdr[0], h = hdr[1], stride = hdr[2], poc = hdr[3], qp = hdr[4], hasMv = hdr[5];
    std::vector<int16_t> org((size_t)h * stride), ref((size_t)h * stride);
    if (std::fread(org.data(), sizeof(int16_t), org.size(), f) != org.size()) return 2;
    if (std::fread(ref.data(), sizeof(int16_t), ref.size(), f) != ref.size()) return 2;
    const int n = (w / 128) * (h / 128);
    std::vector<int16_t> mv((size_t)2 * n);
    if (hasMv == 1 && std::fread(mv.data(), sizeof(int16_t), mv.size(), f) != mv.size()) return 2;
    std::fclose(f);
    mlt_hook::SplitPredictor &p = mlt_hook::SplitPredictor::instance();
    if (!p.enabled()) { std::fprintf(stderr, "predictor disabled\n"); return 3; }
    if (p.pictureSplit(0, 0) != -1) return 5; // nothing before the first pre-pass
    if (!p.pinHostBuffer(org.data(), org.size() * sizeof(int16_t)) || !p.pinHostBuffer(ref.data(), ref.size() * sizeof(int16_t))) return 3;
    if (!p.beginPicture(org.data(), stride, w, h, poc)) return 3;
    if (!p.prepassPicture(ref.data(), stride, hasMv == 1 ? mv.data() : nullptr, qp, hasMv < 0 ? -hasMv : 0)) return 3;
    for (int y = 0; y < h; y += 128)
        for (int x = 0; x < w; x += 128) {
            const bool use = mlt_hook::useCNN(0, false, 128, 128, x, y, w, h);
            const int split = p.pictureSplit(x, y);
            if (!use && split != -1) return 5; // partial CTUs never carry a decision (EncCu.cpp:755)
            std::printf("%d %d %d %d\n", x, y, use ? 1 : 0, split);
        }
    // a new picture forgets the previous picture's decisions until its own pre-pass has run
    if (!p.beginPicture(org.data(), stride, w, h, poc + 1)) return 3;
    if (p.pictureSplit(0, 0) != -1) return 5;
    if (!p.unpinHostBuffer(org.data()) || !p.unpinHostBuffer(ref.data())) return 3;
    return 0;
}
