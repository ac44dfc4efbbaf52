// conv_umma.cu -- instantiations + launcher of the tcgen05 implicit-GEMM conv kernel for the 16
// 3x3 convolutions of the MLT-CNN residual stack (shapes: SURVEY.md section 8a / arch.py:247-254).
#include <cstdlib>

#include "conv_umma.cuh"
#include "mlt_internal.h"

namespace mlt {

template <class C>
static cudaError_t init_one()
{
    return cudaFuncSetAttribute(conv_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
}

template <class C>
static cudaError_t launch_one(const ConvParams &p, int num_sms, cudaStream_t s)
{
    const int ntiles = C::num_tiles(p.nimg);
    if (ntiles <= 0) return cudaSuccess;
    const int grid = ntiles < num_sms ? ntiles : num_sms; // persistent: one CTA per SM
    conv_umma_kernel<C><<<grid, C::NTHREADS, C::SMEM_BYTES, s>>>(p);
    return cudaGetLastError();
}

//            CIN  COUT S  HOUT CSC
using L0a = ConvCfg<32, 32, 2, 64, 0>;    // layer0.0.conv1
using L0b = ConvCfg<32, 32, 1, 64, 32>;   // layer0.0.conv2 + shortcut(conv1 out)
using L0c = ConvCfg<32, 32, 1, 64, 0>;    // layer0.1.conv1 / conv2
using L1a = ConvCfg<32, 64, 2, 32, 0>;    // layer1.0.conv1
using L1b = ConvCfg<64, 64, 1, 32, 32>;   // layer1.0.conv2 + shortcut
using L1c = ConvCfg<64, 64, 1, 32, 0>;    // layer1.1.*
using L2a = ConvCfg<64, 128, 2, 16, 0>;   // layer2.0.conv1
using L2b = ConvCfg<128, 128, 1, 16, 64>; // layer2.0.conv2 + shortcut
using L2c = ConvCfg<128, 128, 1, 16, 0>;  // layer2.1.*
using L3a = ConvCfg<128, 256, 2, 8, 0>;   // layer3.0.conv1
using L3b = ConvCfg<256, 256, 1, 8, 128>; // layer3.0.conv2 + shortcut
using L3c = ConvCfg<256, 256, 1, 8, 0>;   // layer3.1.*

cudaError_t conv_umma_init()
{
    cudaError_t e;
#define MLT_INIT(C) if ((e = init_one<C>()) != cudaSuccess) return e;
    MLT_INIT(L0a) MLT_INIT(L0b) MLT_INIT(L0c) MLT_INIT(L1a) MLT_INIT(L1b) MLT_INIT(L1c)
    MLT_INIT(L2a) MLT_INIT(L2b) MLT_INIT(L2c) MLT_INIT(L3a) MLT_INIT(L3b) MLT_INIT(L3c)
#undef MLT_INIT
    return cudaSuccess;
}

cudaError_t launch_conv_umma(int layer, const __half *in, const __half *w, const __half *bias, const __half *sc_in,
                             const __half *sc_w, const __half *res, __half *out, int nimg, int relu, int num_sms,
                             cudaStream_t s, long long *trace)
{
    ConvParams p;
    p.in = in; p.w = w; p.bias = bias; p.sc_in = sc_in; p.sc_w = sc_w; p.res = res; p.out = out;
    p.nimg = nimg; p.relu = relu; p.trace = trace;
    static const int dbg = getenv("MLT_DEBUG_FLAGS") ? atoi(getenv("MLT_DEBUG_FLAGS")) : 0;
    p.dbg = dbg;
    switch (layer) {
    case 0: return launch_one<L0a>(p, num_sms, s);
    case 1: return launch_one<L0b>(p, num_sms, s);
    case 2: case 3: return launch_one<L0c>(p, num_sms, s);
    case 4: return launch_one<L1a>(p, num_sms, s);
    case 5: return launch_one<L1b>(p, num_sms, s);
    case 6: case 7: return launch_one<L1c>(p, num_sms, s);
    case 8: return launch_one<L2a>(p, num_sms, s);
    case 9: return launch_one<L2b>(p, num_sms, s);
    case 10: case 11: return launch_one<L2c>(p, num_sms, s);
    case 12: return launch_one<L3a>(p, num_sms, s);
    case 13: return launch_one<L3b>(p, num_sms, s);
    case 14: case 15: return launch_one<L3c>(p, num_sms, s);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace mlt
