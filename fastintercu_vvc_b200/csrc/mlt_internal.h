// mlt_internal.h -- declarations shared by the kernels and the C-ABI layer of libmltcnn.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/mltcnn.h"
#include "../../include/mltcnn_cu.h"

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

// Launch with programmatic stream serialization (PDL): the kernel's prologue may overlap the tail of the previous
// kernel in the stream; the kernel itself orders its dependent accesses with griddepcontrol.wait (ptx.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("MLT_NO_PDL") != nullptr; // A/B switch for measurements
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- stage_conv1.cu
// fp32 normalised input tensor [n][2][128][128] (bit-exactness probe of EncCu.cpp:810-867)
cudaError_t launch_stage(const CtuDev *ctus, int n, float *out, cudaStream_t s);
// product path: staging + conv1 (2->32, 3x3, no BN / ReLU; arch.py:278) on tcgen05 -> activation 0, parity-planar fp16
cudaError_t launch_conv1_umma(const CtuDev *ctus, int n, const __half *wop /*[hi,lo][4][32][8]*/, __half *out, cudaStream_t s);
// fp32 cross-check engine: fused staging + conv1 on CUDA cores -> NHWC fp32 [n][128][128][32]
cudaError_t launch_stage_conv1_f(const CtuDev *ctus, int n, const float *w /*[9][2][32]*/, float *out, cudaStream_t s);

// ---- stem_umma.cu : staging + conv1 + layer0.0.conv1 fused (product path); conv1's full output never reaches HBM
cudaError_t stem_umma_init();
cudaError_t launch_stem_umma(const CtuDev *ctus, int n, const __half *w1 /*SEC_STEM_CONV1*/, const __half *w0, const float *bias /*fp32*/,
                             __half *act0q /*conv1 at even rows/cols [n][4][64][64][8]*/, __half *act1, int num_sms, cudaStream_t s);

// ---- stem5_umma.cu : the stem with conv1 and layer0.0.conv1 COMPOSED into one 5x5 stride-2 conv (no activation lies between them,
// arch.py:277-279) + conv1 at the even positions for layer0.0's shortcut; product path since round 2 (MLT_STEM_OLD=1: the kernel above)
cudaError_t stem5_umma_init();
cudaError_t launch_stem5_umma(const CtuDev *ctus, int n, const __half *w /*SEC_STEM5_W*/, const float *corrw /*SEC_STEM5_CORR*/, const float *bias,
                              __half *act0q, __half *act1, int num_sms, cudaStream_t s);
cudaError_t launch_cu_stem5_umma(int size, const CtuDev *cus, int n, const __half *w, const float *corrw, const float *bias, __half *act0q,
                                 __half *act1, int cap, int num_sms, cudaStream_t s);

// ---- stem5_cu16.cu : the composed stem for the 16-px CU network (a tile = two CUs side by side, a work unit = four CUs)
cudaError_t stem5_cu16_init();
cudaError_t launch_cu16_stem5(const CtuDev *cus, int n, const __half *w, const float *corrw, __half *act0q, __half *act1, int cap, int num_sms,
                              cudaStream_t s);

// the same fused stem for the 64- / 32-px CU networks (strip-layout outputs; a 16-px CU is smaller than one work unit)
cudaError_t launch_cu_stem_umma(int size, const CtuDev *cus, int n, const __half *w1, const __half *w0, const float *bias, __half *act0q,
                                __half *act1, int cap, int num_sms, cudaStream_t s);

// ---- conv_simt.cu : fp32 CUDA-core cross-check engine (tests only; never a fallback)
cudaError_t launch_conv_simt(const float *in, const float *w /*[k*k][cin][cout]*/, const float *bias, const float *res,
                             float *out, int nimg, int hin, int cin, int cout, int ksize, int stride, int relu,
                             cudaStream_t s);

// ---- conv_umma.cu : tcgen05 implicit-GEMM engine
// Chunk-planar activation layout.  Element offset of (unit u, plane p, chunk k, row y, sub-image s, column x, e):
//   (((((u * NPL + p) * (C/8) + k) * HP + y) * NIMG + s) * HP + x) * 8 + e
// PAR : four parity planes, plane = (row & 1) * 2 + (col & 1), (y, x) = (row >> 1, col >> 1), HP = H / 2
// PAIR: two images per unit (row-interleaved), image i = unit i >> 1, sub-image i & 1
// STRIP (smaller-CU networks): ONE unit holds the whole batch, strip = images per unit = buffer capacity; unit index 0
struct ActLayout {
    int H, C, par, pair;
    int strip = 0;
    __host__ __device__ int hp() const { return par ? H / 2 : H; }
    __host__ __device__ int npl() const { return par ? 4 : 1; }
    __host__ __device__ int nimg() const { return strip ? strip : (pair ? 2 : 1); }
    __host__ __device__ size_t chunk_stride() const { return (size_t)hp() * nimg() * hp() * 8; }
    __host__ __device__ size_t unit_elems() const { return chunk_stride() * (C / 8) * npl(); }
    __host__ __device__ size_t units_for(int images) const { return strip ? (size_t)1 : (pair ? (size_t)(images + 1) / 2 : (size_t)images); }
};

struct ConvParams {
    CUtensorMap in_map; // main operand: 5-D (x*8, img, row, chunk, unit*plane) over the layer input
    CUtensorMap x_map;  // extra operand (block input: shortcut conv / identity residual); unused when XC == 0
    const __half *w;    // packed [CIN/G][9][G/8][COUT][8]
    const __half *bias; // tcgen05 bias operand [2][COUT][8] fp16: k=0 -> hi(b), k=1 -> lo(b), rest 0 (pack_weights.py)
    const __half *x_w;  // packed [XC/GX][GX/8][COUT][8] (folded shortcut weights, or the identity)
    const float *bias_f32; // the same (shortcut-fused) bias as fp32 [COUT], added in the epilogue by the 32/64-channel layers
    __half *out;
    float *gap_part; // GAP layers only: pool partial sums [tile * NB + sub-image][4 lane quadrants][COUT] of this layer's output
    int nimg;
    int relu;
    long long *trace; // MLT_TRACE_LAYER debug: CTA 0's MMA warp stores clock64() at every tile start (<= 1024 entries), or nullptr
    int dbg;        // MLT_DEBUG_FLAGS (timing experiments only, results invalid): 1 no activation TMA, 2 no stores, 4 no TMEM reads
    int reverse;    // walk the tiles from the last image to the first (L2 reuse across consecutive layers)
    int x_unit_mul; // 4 when the extra operand tensor is parity-planar (plane 0 = even rows, even columns), else 1
    int strip_cap;  // strip layouts (smaller-CU networks): images per strip of the OUTPUT tensor (= buffer capacity)
};

cudaError_t conv_umma_init(); // opt in to large dynamic shared memory for every instantiation; resolve cuTensorMapEncodeTiled
ActLayout conv_umma_out_layout(int layer); // layout of the activation conv `layer` (0..15) writes
// fill p->in_map / x_map / x_unit_mul for conv `layer` reading `in` (and the block input `x` for the second conv of a block);
// `images` = capacity of the buffers in images
cudaError_t conv_umma_prepare(int layer, ConvParams *p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l,
                              size_t images);
cudaError_t launch_conv_umma(int layer, const ConvParams &p, int num_sms, cudaStream_t s);
// Small batches (the in-encoder call: one CTU): convs 1..15 -- everything between the stem and the head -- as ONE kernel on one
// thread-block cluster.  The 15 layers run back to back; a cluster barrier (hardware, ~1 us) replaces the kernel boundary
// (~8 us of launch + prologue + drain each).  p[i] = the ConvParams of conv i + 1 (layer3 = the channel-split variants).
constexpr int CHAIN_NCONV = 15, CHAIN_CLUSTER = 16, CHAIN_MAX_IMAGES = 2;
struct ChainParams {
    ConvParams p[CHAIN_NCONV];
    unsigned long long *trace; // MLT_CHAIN_TRACE debug: CTA 0 stores %globaltimer at kernel start, after every layer and after every barrier
};
cudaError_t launch_conv_chain(const ChainParams &cp, cudaStream_t s);

// layers 12..15 (layer3) also exist with their 256 output channels split over 4 CTAs, as "layers" 16..19: used when a
// batch has fewer tiles than SMs (p.w / p.x_w then point at the split-packed weights)
constexpr int CONV_SPLIT_FIRST = 12, CONV_SPLIT_OFFSET = 4, CONV_SPLIT_WAYS = 4;
// tiny batches: layer3 (12..15) split 8 ways = "layers" 20..23, layer2 (8..11) split 4 ways = "layers" 24..27
constexpr int CONV_SPLIT8_OFFSET = 8, CONV_SPLIT8_WAYS = 8, CONV_L2SPLIT_FIRST = 8, CONV_L2SPLIT_OFFSET = 16, CONV_L2SPLIT_WAYS = 4;
// 5-D tiled tensor map over a chunk-planar activation tensor: dims (x*8, img, row, chunk, unit*plane), box (px, img, rows, chunks, 1)
cudaError_t make_act_map(CUtensorMap *tm, const __half *base, const ActLayout &L, size_t units, int box_px, int box_img,
                         int box_rows, int box_chunks);

// ---- cu_net_*.cu / cu_stem.cu / cu_head.cu : the smaller-CU networks (64 / 32 / 16-px GapBigMltCuORPQ, mlt_cu_or_pq_arch.py:59-130)
constexpr int CU_NCONV = 20, CU_NACT = 21, CU_NHEAD = 4, CU_NLOGIT = 15;
// first conv that READS fp16 hi + lo activation pairs (the conv before it is the first to write one): layer2 on for the
// 64- / 32-px networks, layer3 on for the 16-px one -- what keeps max |dprob| <= 8.6e-4 on 2048 CUs per size with the
// fused stem's single-fp16 conv1 weights (profiles/r01/precision_cu_2048.log)
constexpr int cu_hilo_from(int size) { return size == 16 ? 12 : 8; }
struct CuLayerInfo { // one 3x3 conv of the CU network at a given CU size (forward order, after conv1)
    int cin, cout, stride, hout, xc, out_par, nb, flat; // stride as executed (a stride-2 conv on a 1x1 map runs as stride 1)
    int g, gx;                                          // channels per weight slab of the main / extra operand (packer layout)
    int gap_count;                                      // pool partial vectors per image its epilogue writes (0: none)
    int hilo_out;                                       // writes an fp16 hi + lo pair (lo tensor right behind the hi tensor)
};
cudaError_t cu_conv_init(int size);                    // opt in to large dynamic shared memory for this size's kernels
cudaError_t cu_conv_info(int size, int layer, CuLayerInfo *info);
// fill p->in_map / x_map for conv `layer` of the `size`-px network; every tensor is a strip of `cap` images
cudaError_t cu_conv_prepare(int size, int layer, ConvParams *p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l);
cudaError_t launch_cu_conv(int size, int layer, const ConvParams &p, int num_sms, cudaStream_t s);
// staging + conv1 on tcgen05 -> activation 0 (H = size, 32 channels, parity-planar strip of `cap` images)
cudaError_t launch_cu_conv1(int size, const CtuDev *cus, int n, const __half *wop, __half *out, int cap, cudaStream_t s);
struct CuHeadParams {
    const __half *act[CU_NHEAD]; // outputs of layer1..layer4 (strip layouts)
    ActLayout lay[CU_NHEAD];
    int hilo[CU_NHEAD];              // the stored activation is an fp16 hi + lo pair
    const float *gap_part[CU_NHEAD]; // maps >= 8x8: fp32 pool partial sums written by that conv's epilogue [image][gap_count][C], else nullptr
    int gap_count[CU_NHEAD];
    const float *fc_w[CU_NHEAD]; // [out][in]
    const float *fc_b[CU_NHEAD];
    const CtuDev *cus;           // poc / qp
    mlt_cu_result *out;
    int n;
};
cudaError_t launch_cu_head(const CuHeadParams &p, cudaStream_t s);

// ---- head.cu : global average pools + FC heads + softmax + argmax + flags (arch.py:281-297, EncCu.cpp:913-921)
struct HeadParams {
    const void *act[3];   // fp32 cross-check engine: layer1 / layer2 / layer3 outputs, dense NHWC fp32
    const float *gap_part[3]; // product path: per-tile pool partial sums written by the epilogues of convs 7 / 11 / 15
    const float *fc_w[3]; // [out][in]
    const float *fc_b[3];
    const CtuDev *ctus;   // poc / qp
    mlt_result *out;
    int n;
};
cudaError_t launch_head_h(const HeadParams &p, cudaStream_t s);
cudaError_t launch_head_f(const HeadParams &p, cudaStream_t s);

// ---- picture_pred.cu : frame-level pre-pass -- integer-MV prediction blocks for every eligible CTU of a picture from one
// reference luma plane with replicated borders (Picture.cpp:1117); writes plane 1 of the dense [n][2][128][128] batch
struct PicCtu { int32_t x, y, mvx, mvy; }; // CTU position and MV in luma samples
cudaError_t launch_picture_pred(const int16_t *ref, int pitch, int w, int h, const PicCtu *ctus, int n, int16_t *out, cudaStream_t s);
// the smaller-CU form: every size x size block of the picture's CU raster as a dense [n][2][size][size] batch (plane 0 = org
// block, plane 1 = integer-MV prediction, mv = device [n][2] or nullptr) + its (poc, qp) pairs
cudaError_t launch_picture_cu_gather(const int16_t *org, const int16_t *ref, int pitch, int w, int h, int size, int n, const int16_t *mv, int poc,
                                     int qp, int16_t *out, int32_t *pocqp, cudaStream_t s);

// ---- picture_me.cu : integer full-search block matching per eligible CTU (pre-pass MVs); R <= MLT_ME_MAX_RANGE
constexpr int MLT_ME_MAX_RANGE = 16;
constexpr size_t picture_me_smem_bytes(int R) { return (size_t)(16 * MLT_CTU_SIZE + (16 + 2 * R) * (MLT_CTU_SIZE + 2 * R + 2)) * sizeof(int16_t); }
// cost: scratch [n][(2R+1)^2] u32; mv: [n][2] int16 (x, y); best_cost: [n] u32 or nullptr.  2 kernels + 1 memset.
cudaError_t launch_picture_me(const int16_t *org, const int16_t *ref, int pitch, int w, int h, const PicCtu *ctus, int n, int R, unsigned *cost,
                              int16_t *mv, unsigned *best_cost, cudaStream_t s);

// ---- pack10.cu : 10-bit packed Pel transport (4 samples per 5 bytes); device side of mlt_*_packed10
cudaError_t launch_unpack10(const uint8_t *packed, int16_t *out, size_t samples, cudaStream_t s);

// ---- misc
cudaError_t launch_unpack_act(const __half *in, float *out, int nimg, const ActLayout &L, cudaStream_t s);

} // namespace mlt
