// ref_libtorch.cpp -- CPU ORACLE / BASELINE (test infrastructure, NOT product code): the reference hook's libtorch call
// sequence as a stand-alone C++ program, for timing the reference's CPU path the way VTM would run it
// (vtm-mlt-cpp/source/Lib/EncoderLib/EncCu.cpp:803-926 with at::kCPU): per CTU
//   xMalloc copies + (uint16_t) casts (:810-830), |org - pred| (:832-833, cv::absdiff on CV_16UC1), * (float)(1/1023) and
//   clamp (:835-867), from_blob NHWC x2 + cat(dim 3) + permute{0,3,1,2} (:869-877), poc / qp int tensors (:881-882),
//   torch::jit::load (:894-905; once, or per call = the hook as written), forward + toTuple (:909),
//   elements()[2] + argmax(1).item (:912-921).
// OpenCV is not linked: its two calls are restated inline (absdiff exact; convertTo = (float)v * (float)alpha, the
// formula pinned by tests/golden/stage_kat.npz).  Written from the reference's behaviour, not copied from it.
//   usage: ref_libtorch model.pt ctus.bin seconds threads mode     mode: 0 = load once, 1 = load per call (as written)
//   ctus.bin: int32 n, then n x { int32 poc, int32 qp, int16 org[128*128], int16 pred[128*128] }
//   stdout  : "<ctus> <seconds> <threads>" and one line "split <i> <argmax> <logit0..3>" for the first 8 CTUs
#include <ATen/Parallel.h>
#include <torch/script.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv)
{
    if (argc != 6) { std::fprintf(stderr, "usage: ref_libtorch model.pt ctus.bin seconds threads mode\n"); return 2; }
    const double budget = std::atof(argv[3]);
    const int threads = std::atoi(argv[4]), mode = std::atoi(argv[5]);
    if (threads > 0) at::set_num_threads(threads);
    FILE *f = std::fopen(argv[2], "rb");
    int32_t n = 0;
    if (!f || std::fread(&n, 4, 1, f) != 1 || n <= 0) return 2;
    constexpr int S = 128;
    std::vector<int32_t> poc(n), qp(n);
    std::vector<int16_t> org((size_t)n * S * S), pred((size_t)n * S * S);
    for (int i = 0; i < n; i++) {
        if (std::fread(&poc[i], 4, 1, f) != 1 || std::fread(&qp[i], 4, 1, f) != 1) return 2;
        if (std::fread(&org[(size_t)i * S * S], 2, S * S, f) != (size_t)S * S) return 2;
        if (std::fread(&pred[(size_t)i * S * S], 2, S * S, f) != (size_t)S * S) return 2;
    }
    std::fclose(f);
    const at::Device device = at::kCPU;
    torch::NoGradGuard no_grad;
    torch::jit::script::Module cnn;
    try {
        cnn = torch::jit::load(argv[1], device);
        cnn.eval();
    } catch (const c10::Error &) { std::fprintf(stderr, "error loading the model\n"); return 3; }
    const float alpha = (float)(1.0 / 1023);
    std::vector<float> orgF((size_t)S * S), resF((size_t)S * S);
    std::vector<uint16_t> orgU((size_t)S * S), predU((size_t)S * S);
    long done = 0;
    const auto t0 = std::chrono::steady_clock::now();
    double dt = 0;
    for (int warm = 1; warm >= 0; warm--) {
        done = 0;
        const auto ts = std::chrono::steady_clock::now();
        for (;;) {
            const int i = (int)(done % n);
            const int16_t *o = &org[(size_t)i * S * S], *p = &pred[(size_t)i * S * S];
            for (int k = 0; k < S * S; k++) { // dense copies with the (uint16_t) cast, absolute difference, scale, clamp
                orgU[k] = (uint16_t)o[k];
                predU[k] = (uint16_t)p[k];
                const uint16_t r = orgU[k] > predU[k] ? orgU[k] - predU[k] : predU[k] - orgU[k];
                float a = (float)orgU[k] * alpha, b = (float)r * alpha;
                orgF[k] = a < 0.f ? 0.f : (a > 1.f ? 1.f : a);
                resF[k] = b < 0.f ? 0.f : (b > 1.f ? 1.f : b);
            }
            at::Tensor tOrg = torch::from_blob(orgF.data(), {1, S, S, 1}, at::kFloat);
            at::Tensor tRes = torch::from_blob(resF.data(), {1, S, S, 1}, at::kFloat);
            at::Tensor in = torch::cat({tOrg, tRes}, 3).to(device).permute({0, 3, 1, 2});
            at::Tensor tPoc = torch::tensor({poc[i]}).to(device), tQp = torch::tensor({qp[i]}).to(device);
            if (mode == 1) { cnn = torch::jit::load(argv[1], device); cnn.eval(); }
            auto outputs = cnn.forward({in, tPoc, tQp}).toTuple();
            at::Tensor out = outputs->elements()[2].toTensor().cpu().detach();
            const int split = out.argmax(1).item().toInt();
            if (!warm && done < 8)
                std::printf("split %ld %d %.9g %.9g %.9g %.9g\n", done, split, out[0][0].item<float>(), out[0][1].item<float>(),
                            out[0][2].item<float>(), out[0][3].item<float>());
            done++;
            dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - ts).count();
            if (warm ? done >= 3 : (dt >= budget && done >= 8)) break;
        }
    }
    (void)t0;
    std::printf("%ld %.6f %d\n", done, dt, at::get_num_threads());
    return 0;
}
