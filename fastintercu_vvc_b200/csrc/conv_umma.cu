// conv_umma.cu -- instantiations, TMA tensor maps and launcher of the tcgen05 implicit-GEMM conv kernel for the 16
// 3x3 convolutions of the MLT-CNN residual stack (shapes: SURVEY.md section 8a / arch.py:247-254).
#include <cstdlib>

#include "conv_umma.cuh"
#include "mlt_internal.h"

namespace mlt {

//            CIN  COUT S  HOUT XC  OUT_PAR
using L0a = ConvCfg<32, 32, 2, 64, 0, 0>;     // layer0.0.conv1
using L0b = ConvCfg<32, 32, 1, 64, 32, 0, 1>;   // layer0.0.conv2 + shortcut(conv1 out)
using L0c = ConvCfg<32, 32, 1, 64, 0, 0>;     // layer0.1.conv1
using L0d = ConvCfg<32, 32, 1, 64, 32, 1>;    // layer0.1.conv2 + identity; output feeds a stride-2 block
using L1a = ConvCfg<32, 64, 2, 32, 0, 0>;     // layer1.0.conv1
using L1b = ConvCfg<64, 64, 1, 32, 32, 0, 1>;   // layer1.0.conv2 + shortcut
using L1c = ConvCfg<64, 64, 1, 32, 0, 0>;     // layer1.1.conv1
using L1d = ConvCfg<64, 64, 1, 32, 64, 1>;    // layer1.1.conv2 + identity
using L2a = ConvCfg<64, 128, 2, 16, 0, 0>;    // layer2.0.conv1
using L2b = ConvCfg<128, 128, 1, 16, 64, 0, 1>; // layer2.0.conv2 + shortcut
using L2c = ConvCfg<128, 128, 1, 16, 0, 0>;   // layer2.1.conv1
using L2d = ConvCfg<128, 128, 1, 16, 128, 1>; // layer2.1.conv2 + identity
using L3a = ConvCfg<128, 256, 2, 8, 0, 0>;    // layer3.0.conv1
using L3b = ConvCfg<256, 256, 1, 8, 128, 0, 1>; // layer3.0.conv2 + shortcut
using L3c = ConvCfg<256, 256, 1, 8, 0, 0>;    // layer3.1.conv1
using L3d = ConvCfg<256, 256, 1, 8, 256, 0>;  // layer3.1.conv2 + identity
// small batches (fewer tiles than SMs): the same four layers with the 256 output channels split over 4 CTAs ("layers" 16..19)
using L3aS = ConvCfg<128, 256, 2, 8, 0, 0, 0, 0, 4>;
using L3bS = ConvCfg<256, 256, 1, 8, 128, 0, 1, 0, 4>;
using L3cS = ConvCfg<256, 256, 1, 8, 0, 0, 0, 0, 4>;
using L3dS = ConvCfg<256, 256, 1, 8, 256, 0, 0, 0, 4>;

#define MLT_FOR_LAYER(li, F)                                                                                            \
    switch (li) {                                                                                                       \
    case 0: F(L0a); break;  case 1: F(L0b); break;  case 2: F(L0c); break;  case 3: F(L0d); break;                      \
    case 4: F(L1a); break;  case 5: F(L1b); break;  case 6: F(L1c); break;  case 7: F(L1d); break;                      \
    case 8: F(L2a); break;  case 9: F(L2b); break;  case 10: F(L2c); break; case 11: F(L2d); break;                     \
    case 12: F(L3a); break; case 13: F(L3b); break; case 14: F(L3c); break; case 15: F(L3d); break;                     \
    case 16: F(L3aS); break; case 17: F(L3bS); break; case 18: F(L3cS); break; case 19: F(L3dS); break;                 \
    default: return cudaErrorInvalidValue;                                                                              \
    }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;

cudaError_t conv_umma_init()
{
    cudaError_t e;
#define MLT_INIT(C) if ((e = cudaFuncSetAttribute(conv_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)) != cudaSuccess) return e
    for (int li = 0; li < 20; li++) { MLT_FOR_LAYER(li, MLT_INIT) }
#undef MLT_INIT
    if (!g_encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    return cudaSuccess;
}

// 5-D tiled map over a chunk-planar activation tensor: dims (x*8, img, row, chunk, unit*plane), fp16, no swizzle.
cudaError_t make_act_map(CUtensorMap *tm, const __half *base, const ActLayout &L, size_t units, int box_px, int box_img,
                         int box_rows, int box_chunks)
{
    if (!g_encode) return cudaErrorNotReady;
    const cuuint64_t hp = (cuuint64_t)L.hp(), ni = (cuuint64_t)L.nimg();
    cuuint64_t dims[5] = {hp * 8, ni, hp, (cuuint64_t)(L.C / 8), (cuuint64_t)units * L.npl()};
    cuuint64_t strides[4] = {hp * 16, ni * hp * 16, hp * ni * hp * 16, (cuuint64_t)(L.C / 8) * hp * ni * hp * 16};
    cuuint32_t box[5] = {(cuuint32_t)box_px * 8, (cuuint32_t)box_img, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half *>(base), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <class C>
static cudaError_t prepare_one(ConvParams &p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l, size_t images)
{
    // the layouts the kernel template assumes must be the layouts the buffers were allocated with
    if (in_l.C != C::CIN || in_l.H != C::HOUT * C::STRIDE || in_l.par != (C::STRIDE == 2) || in_l.pair != C::IN_PAIR) return cudaErrorInvalidValue;
    cudaError_t e = make_act_map(&p.in_map, in, in_l, in_l.units_for((int)images), C::BLKW, C::NB, C::PROWS, C::CH);
    if (e != cudaSuccess) return e;
    if constexpr (C::XC > 0) {
        if (!x || !x_l || x_l->C != C::XC || x_l->hp() != C::HOUT || x_l->pair != C::IN_PAIR) return cudaErrorInvalidValue;
        e = make_act_map(&p.x_map, x, *x_l, x_l->units_for((int)images), 8, C::NB, C::TR, C::GX / 8);
        if (e != cudaSuccess) return e;
        p.x_unit_mul = x_l->npl();
    } else {
        p.x_map = p.in_map;
        p.x_unit_mul = 1;
    }
    return cudaSuccess;
}

cudaError_t conv_umma_prepare(int layer, ConvParams *p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l,
                              size_t images)
{
#define MLT_PREP(C) return prepare_one<C>(*p, in, in_l, x, x_l, images)
    MLT_FOR_LAYER(layer, MLT_PREP)
#undef MLT_PREP
    return cudaSuccess;
}

ActLayout conv_umma_out_layout(int layer)
{
    ActLayout l{0, 0, 0, 0};
#define MLT_OUT(C) l = ActLayout{C::HOUT, C::COUT, C::OUT_PAR, C::OUT_PAIR}
    switch (layer) {
    case 0: MLT_OUT(L0a); break;  case 1: MLT_OUT(L0b); break;  case 2: MLT_OUT(L0c); break;  case 3: MLT_OUT(L0d); break;
    case 4: MLT_OUT(L1a); break;  case 5: MLT_OUT(L1b); break;  case 6: MLT_OUT(L1c); break;  case 7: MLT_OUT(L1d); break;
    case 8: MLT_OUT(L2a); break;  case 9: MLT_OUT(L2b); break;  case 10: MLT_OUT(L2c); break; case 11: MLT_OUT(L2d); break;
    case 12: case 16: MLT_OUT(L3a); break; case 13: case 17: MLT_OUT(L3b); break; case 14: case 18: MLT_OUT(L3c); break;
    case 15: case 19: MLT_OUT(L3d); break;
    default: break;
    }
#undef MLT_OUT
    return l;
}

template <class C>
static cudaError_t launch_one(const ConvParams &p, int num_sms, cudaStream_t s)
{
    const int ntiles = C::num_tiles(p.nimg);
    if (ntiles <= 0) return cudaSuccess;
    // persistent: one CTA per SM (channel-split layers: a multiple of the split, so that the tile pair of a pass shares weights)
    const int grid = ntiles < num_sms ? ntiles : num_sms / C::NSPLIT * C::NSPLIT;
    return launch_pdl(conv_umma_kernel<C>, dim3(grid), dim3(C::NTHREADS), C::SMEM_BYTES, s, p);
}

cudaError_t launch_conv_umma(int layer, const ConvParams &p, int num_sms, cudaStream_t s)
{
#define MLT_LAUNCH(C) return launch_one<C>(p, num_sms, s)
    MLT_FOR_LAYER(layer, MLT_LAUNCH)
#undef MLT_LAUNCH
    return cudaSuccess;
}

} // namespace mlt
