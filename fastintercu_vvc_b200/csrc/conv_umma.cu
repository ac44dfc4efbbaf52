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
// tiny batches (the in-encoder call): layer3 split 8 ways ("layers" 20..23) and layer2 split 4 ways ("layers" 24..27)
using L3aS8 = ConvCfg<128, 256, 2, 8, 0, 0, 0, 0, 8>;
using L3bS8 = ConvCfg<256, 256, 1, 8, 128, 0, 1, 0, 8>;
using L3cS8 = ConvCfg<256, 256, 1, 8, 0, 0, 0, 0, 8>;
using L3dS8 = ConvCfg<256, 256, 1, 8, 256, 0, 0, 0, 8>;
using L2aS = ConvCfg<64, 128, 2, 16, 0, 0, 0, 0, 4>;
using L2bS = ConvCfg<128, 128, 1, 16, 64, 0, 1, 0, 4>;
using L2cS = ConvCfg<128, 128, 1, 16, 0, 0, 0, 0, 4>;
using L2dS = ConvCfg<128, 128, 1, 16, 128, 1, 0, 0, 4>;

#define MLT_FOR_LAYER(li, F)                                                                                            \
    switch (li) {                                                                                                       \
    case 0: F(L0a); break;  case 1: F(L0b); break;  case 2: F(L0c); break;  case 3: F(L0d); break;                      \
    case 4: F(L1a); break;  case 5: F(L1b); break;  case 6: F(L1c); break;  case 7: F(L1d); break;                      \
    case 8: F(L2a); break;  case 9: F(L2b); break;  case 10: F(L2c); break; case 11: F(L2d); break;                     \
    case 12: F(L3a); break; case 13: F(L3b); break; case 14: F(L3c); break; case 15: F(L3d); break;                     \
    case 16: F(L3aS); break; case 17: F(L3bS); break; case 18: F(L3cS); break; case 19: F(L3dS); break;                 \
    case 20: F(L3aS8); break; case 21: F(L3bS8); break; case 22: F(L3cS8); break; case 23: F(L3dS8); break;             \
    case 24: F(L2aS); break; case 25: F(L2bS); break; case 26: F(L2cS); break; case 27: F(L2dS); break;                 \
    default: return cudaErrorInvalidValue;                                                                              \
    }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
cudaError_t conv_chain_init();

cudaError_t conv_umma_init()
{
    cudaError_t e;
#define MLT_INIT(C) if ((e = cudaFuncSetAttribute(conv_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)) != cudaSuccess) return e
    for (int li = 0; li < 28; li++) { MLT_FOR_LAYER(li, MLT_INIT) }
#undef MLT_INIT
    if ((e = conv_chain_init()) != cudaSuccess) return e;
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
    case 8: case 24: MLT_OUT(L2a); break;  case 9: case 25: MLT_OUT(L2b); break;  case 10: case 26: MLT_OUT(L2c); break; case 11: case 27: MLT_OUT(L2d); break;
    case 12: case 16: case 20: MLT_OUT(L3a); break; case 13: case 17: case 21: MLT_OUT(L3b); break; case 14: case 18: case 22: MLT_OUT(L3c); break;
    case 15: case 19: case 23: MLT_OUT(L3d); break;
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

// ---- conv_chain_kernel: the 15 convs after the stem on one cluster (small batches), see mlt_internal.h
constexpr int chain_max(int a, int b) { return a > b ? a : b; }
constexpr int CHAIN_SMEM = chain_max(chain_max(chain_max(L0b::SMEM_BYTES, L0d::SMEM_BYTES), chain_max(L1a::SMEM_BYTES, L1b::SMEM_BYTES)),
                                     chain_max(chain_max(chain_max(L1d::SMEM_BYTES, L2aS::SMEM_BYTES), chain_max(L2bS::SMEM_BYTES, L2dS::SMEM_BYTES)),
                                               chain_max(chain_max(L3aS8::SMEM_BYTES, L3bS8::SMEM_BYTES), chain_max(L3cS8::SMEM_BYTES, L3dS8::SMEM_BYTES))));
static_assert(CHAIN_SMEM + 1024 <= 227 * 1024, "room for the static tmem slot");
static_assert(CHAIN_SMEM >= L0c::SMEM_BYTES && CHAIN_SMEM >= L1c::SMEM_BYTES && CHAIN_SMEM >= L2cS::SMEM_BYTES, "chain smem covers every layer");

__global__ void __launch_bounds__(384, 1) conv_chain_kernel(const __grid_constant__ ChainParams cp)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 8) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); } // one allocation for all 15 layers (the widest needs 512 columns)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    griddep_launch_dependents();
    griddep_wait(); // the stem's outputs
    const int bid = (int)blockIdx.x, nblk = (int)gridDim.x; // the whole grid is one cluster
    // between two layers: this CTA's epilogue stores (generic proxy) must be visible to every CTA's TMA loads (async proxy)
    auto layer_barrier = [&]() {
        fence_proxy_async_all();
        __threadfence();
        cluster_arrive_release();
        cluster_wait_acquire();
        fence_proxy_async_all();
    };
    auto stamp = [&](int k) {
        if (cp.trace != nullptr && bid == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            cp.trace[k] = t;
        }
    };
    stamp(0);
#define MLT_CHAIN_STEP(i, C) conv_layer_body<C, true>(cp.p[i], smem, tmem, bid, nblk); stamp(1 + 2 * i); layer_barrier(); stamp(2 + 2 * i);
    MLT_CHAIN_STEP(0, L0b) MLT_CHAIN_STEP(1, L0c) MLT_CHAIN_STEP(2, L0d)
    MLT_CHAIN_STEP(3, L1a) MLT_CHAIN_STEP(4, L1b) MLT_CHAIN_STEP(5, L1c) MLT_CHAIN_STEP(6, L1d)
    MLT_CHAIN_STEP(7, L2aS) MLT_CHAIN_STEP(8, L2bS) MLT_CHAIN_STEP(9, L2cS) MLT_CHAIN_STEP(10, L2dS)
    MLT_CHAIN_STEP(11, L3aS8) MLT_CHAIN_STEP(12, L3bS8) MLT_CHAIN_STEP(13, L3cS8) MLT_CHAIN_STEP(14, L3dS8)
#undef MLT_CHAIN_STEP
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

cudaError_t conv_chain_init()
{
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

cudaError_t launch_conv_chain(const ChainParams &cp, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CHAIN_CLUSTER);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = CHAIN_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CHAIN_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, conv_chain_kernel, cp);
}

cudaError_t launch_conv_umma(int layer, const ConvParams &p, int num_sms, cudaStream_t s)
{
#define MLT_LAUNCH(C) return launch_one<C>(p, num_sms, s)
    MLT_FOR_LAYER(layer, MLT_LAUNCH)
#undef MLT_LAUNCH
    return cudaSuccess;
}

} // namespace mlt
