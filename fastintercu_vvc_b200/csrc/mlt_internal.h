// mlt_internal.h -- declarations shared by the kernels and the C-ABI layer of libmltcnn.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mltcnn.h"

namespace mlt {

// Device-side view of one CTU (all pointers are DEVICE pointers; strides in int16 elements).
// The library owns the device copies, so org/pred rows are always 16-byte aligned here.
struct CtuDev {
    const int16_t *org;
    const int16_t *pred;
    int32_t org_stride;
    int32_t pred_stride;
    int32_t poc;
    int32_t qp;
};

struct LayerDesc { // one 3x3 conv of the residual stack (forward order, after conv1)
    int cin, cout, stride, hout, sc; // sc = shortcut index or -1
};

// (float)(1.0/1023): the alpha OpenCV's convertTo uses at EncCu.cpp:835-838, bit pattern 0x3A802008
#define MLT_ALPHA_BITS 0x3A802008u

// ---- stage_conv1.cu
// fp32 normalised input tensor [n][2][128][128] (bit-exactness probe of EncCu.cpp:810-867)
cudaError_t launch_stage(const CtuDev *ctus, int n, float *out, cudaStream_t s);
// fused staging + conv1 (2->32, 3x3, no BN / ReLU; arch.py:278) -> NHWC [n][128][128][32]
cudaError_t launch_stage_conv1_h(const CtuDev *ctus, int n, const float *w /*[9][2][32]*/, __half *out, cudaStream_t s);
cudaError_t launch_stage_conv1_f(const CtuDev *ctus, int n, const float *w, float *out, cudaStream_t s);

// ---- conv_simt.cu : fp32 CUDA-core cross-check engine (tests only; never a fallback)
cudaError_t launch_conv_simt(const float *in, const float *w /*[k*k][cin][cout]*/, const float *bias, const float *res,
                             float *out, int nimg, int hin, int cin, int cout, int ksize, int stride, int relu,
                             cudaStream_t s);

// ---- conv_umma.cu : tcgen05 implicit-GEMM engine
struct ConvParams;
cudaError_t launch_conv_umma(int layer, const __half *in, const __half *w, const __half *bias, const __half *sc_in,
                             const __half *sc_w, const __half *res, __half *out, int nimg, int relu, int num_sms,
                             cudaStream_t s, long long *trace = nullptr);
cudaError_t conv_umma_init(); // opt in to large dynamic shared memory for every instantiation

// ---- head.cu : global average pools + FC heads + softmax + argmax + flags (arch.py:281-297, EncCu.cpp:913-921)
struct HeadParams {
    const void *act[3];   // layer1 / layer2 / layer3 outputs: haloed NHWC fp16 (product) or dense NHWC fp32 (cross-check)
    const float *fc_w[3]; // [out][in]
    const float *fc_b[3];
    const CtuDev *ctus;   // poc / qp
    mlt_result *out;
    int n;
};
cudaError_t launch_head_h(const HeadParams &p, cudaStream_t s);
cudaError_t launch_head_f(const HeadParams &p, cudaStream_t s);

// ---- misc
cudaError_t launch_unhalo_to_float(const __half *in, float *out, int nimg, int h, int c, cudaStream_t s);

} // namespace mlt
