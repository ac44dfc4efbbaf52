// conv_umma.cuh -- implicit-GEMM 3x3 convolution (+ fused 1x1 stride-2 shortcut conv OR identity residual, bias,
// ReLU) on tcgen05 tensor cores with TMEM accumulators, fed by TMA, for the MLT-CNN residual stack
// (reference graph: mlt_ctu_or_pq_arch.py:52-57 BasicBlock, :273-299 forward; BN folded by pack_weights.py).
//
// GEMM view per tile:  D[128 pixels x COUT] = bias + sum over (cin_group, tap) A_tap[128 x G] * W_tap[COUT x G]^T
//                                                  + X[128 x XC] * Wx[COUT x XC]^T        (shortcut conv, or Wx = I)
//
// Design (B200-first, not an im2col port):
//  * Activations live in HBM "chunk-planar": [unit][plane][C/8][row][img][x][8 ch] fp16 (ActLayout below), i.e. the
//    8-channel / 16-byte K-chunk of the UMMA "K-major, no-swizzle" core matrix is the innermost unit and pixels of a
//    row are 16 B apart.  ONE 5-D TMA box load per (tile, cin_group) brings the input patch WITH its halo into shared
//    memory already in operand layout [chunk][patch row][img][patch px][8]; the convolution's zero padding is the
//    TMA out-of-bounds fill (negative / past-the-edge box coordinates), so there are no halo copies, no bounds
//    checks and no im2col.  Each of the nine taps is just a different START ADDRESS of the same smem patch in the
//    tcgen05.mma operand descriptor (SBO = patch-row pitch, LBO = chunk pitch).
//  * Tensors that feed a stride-2 block are stored as four (row, column) parity planes, so every tap of the stride-2
//    conv -- and the 1x1 stride-2 shortcut, which is exactly plane (even, even) -- is again a dense window.
//  * The 8x8 maps of the last stage are stored as row-interleaved image PAIRS so that a 128-row tile spans two images.
//  * The BasicBlock's second conv takes one extra K-slab per 32/64 input channels: the block input X times the folded
//    shortcut weights, or times the identity matrix for an identity residual (exact in fp16 x fp32-accumulate).  The
//    epilogue therefore never loads anything: TMEM -> regs -> ReLU -> fp16 -> global.
//  * Folded weights are pre-packed offline in exactly the smem operand layout; each (cin_group, tap) slab is one
//    contiguous cp.async.bulk into a ring, or the whole layer is resident in smem for the 32/64-channel layers.
//  * Warp-specialised persistent CTA (1 per SM, 12 warps): 2 x 4 epilogue warps taking alternate tiles, 1 MMA-issuer
//    warp (one elected thread issues tcgen05.mma; 2 issuer warps on alternate tiles for the resident-weight layers,
//    whose MMA stream is otherwise instruction-issue-bound), 1 weight-copy warp, 1 activation-TMA warp.  2-4 TMEM accumulator
//    stages overlap the epilogue of tile i with the MMAs of tiles i+1...; the bias enters the accumulator through one
//    extra K=16 MMA (ones x [hi(b), lo(b)]).
#pragma once
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

// FLAGS_ bit 0 (STRIP): the smaller-CU networks (64 / 32 / 16-px GapBigMltCuORPQ, mlt_cu_or_pq_arch.py:59-130).  Their maps
// shrink to 4x4, 2x2 and 1x1, so every tensor is stored as ONE strip [plane][C/8][row][img][x][8] over the whole batch
// (ActLayout::strip) and a tile takes NB consecutive images of it.  Maps of 8x8 and larger use the same tile shapes as
// the CTU network; maps <= 4x4 use FLAT tiles: the TMA box [row][NB images][x incl. halo] is addressed as 128
// CONSECUTIVE 16-byte positions (SBO = 128 B), tap (kh, kw) is again just a start-address shift, and the accumulator
// rows that land on halo positions (x >= HOUT) are simply not stored -- M efficiency HOUT / (HOUT + 2), on < 15 % of
// the network's FLOPs.
// NSPLIT_ > 1 (small batches of the CTU network's last stage): a tile's COUT output channels are split over NSPLIT CTAs
// ("items" = tile x split).  With fewer tiles than SMs every CTA otherwise streams the whole layer's weights (1.2 MB for
// 256 -> 256) through one SM and issues all of the tile's MMAs alone; split four ways each CTA streams a quarter
// ([split][cin_group][tap][G/8][NC][8], packed offline) and the four run on different SMs.  Per output element the MMA
// sequence is unchanged, so results are bit-identical to the unsplit kernel.  Measured: 22 -> 19 us per layer at n = 1
// (a deeper slab ring changes nothing: what remains is the per-slab issue loop, 72 iterations of wait / 2 MMAs / commit,
// and ~8 us of launch + prologue + drain per kernel), single-CTU hook latency 177 -> 162 us.
template <int CIN_, int COUT_, int STRIDE_, int HOUT_, int XC_, int OUT_PAR_, int XLO_ = 0, int FLAGS_ = 0, int NSPLIT_ = 1>
struct ConvCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, STRIDE = STRIDE_, HOUT = HOUT_, XC = XC_, OUT_PAR = OUT_PAR_;
    static constexpr int NSPLIT = NSPLIT_, NC = COUT / NSPLIT; // output channels per CTA = GEMM N
    static constexpr bool STRIP = (FLAGS_ & 1) != 0;
    static constexpr bool FLAT = STRIP && HOUT <= 4;
    // FLAGS_ bits 1 / 2 (HILO_IN / HILO_OUT): the last two stages of the CU networks keep their activations as an fp16
    // hi + lo PAIR (lo = fp16(x - hi), stored right behind the hi tensor): their small maps pool only 16 / 4 / 1 pixels,
    // so the fp16 rounding of activations does not average out there (tools/emulate_cu_precision.py).  A hi+lo input
    // costs nothing in weight traffic: the "tile pair" of the streamed-weight path becomes (hi box, lo box) of ONE tile,
    // both accumulating into the same TMEM accumulator.
    static constexpr bool HILO_IN = (FLAGS_ & 2) != 0, HILO_OUT = (FLAGS_ & 4) != 0;
    // XLO: the extra operand's weights come as an fp16 hi + lo pair (two MMA passes over the same activation stage): the
    // folded 1x1 shortcut weights are the largest single source of fp16 weight-rounding error and cost < 1 % to do exactly
    static constexpr int XP = 1 + XLO_;
    // input channels per A stage / weight slab: 32 for the stride-2 convs (four parity planes per stage) and for the
    // 256-channel layers (keeps the activation ring small so the streamed-weight ring can be deep)
    // (FLAT tiles carry NB images' whole patches per stage: 16 / 32 channels keep the stage small)
    static constexpr int G = (STRIDE == 2 && (COUT >= 256 || FLAT)) ? 16 : ((STRIDE == 2 || COUT >= 256 || FLAT) ? 32 : (CIN % 64 != 0 ? 32 : 64));
    static constexpr int NCG = CIN / G;
    static constexpr int CH = G / 8;               // 16-byte chunks per pixel per A stage
    // 1x1 maps: eight of the nine taps of a stride-1 3x3 conv only ever see zero padding, so the conv IS its centre tap --
    // no halo, 128 images per tile, one weight slab per channel group instead of nine
    static constexpr bool CENTER_ONLY = FLAT && HOUT == 1 && STRIDE == 1;
    static constexpr int HALO = CENTER_ONLY ? 0 : 1;
    static constexpr int TAP0 = CENTER_ONLY ? 4 : 0, NTAPS = CENTER_ONLY ? 1 : 9;
    static constexpr int BLKW = FLAT ? (STRIDE == 1 ? HOUT + 2 * HALO : HOUT + 1) : ((STRIDE == 1) ? 10 : 9); // patch pixels per block (halo included)
    static constexpr int NB = FLAT ? 128 / (HOUT * BLKW) : ((HOUT == 8) ? 2 : 1); // images per tile
    static constexpr int TR = FLAT ? HOUT : 128 / (8 * NB); // tile rows; tile = TR x (NB * 8) pixels [FLAT: TR x NB x BLKW positions]
    static constexpr int PITCH = BLKW * NB;
    static constexpr int PROWS = (STRIDE == 1) ? TR + 2 * HALO : TR + 1;
    static constexpr int NPLANES = (STRIDE == 1) ? 1 : 4;
    static constexpr int PLANE_PX = PROWS * PITCH;
    static constexpr int PLANE_BYTES = CH * PLANE_PX * 16;              // one TMA box
    static constexpr int PLANE_STRIDE = (PLANE_BYTES + 127) / 128 * 128; // TMA destinations are 128-byte aligned
    static constexpr int A_LBO = PLANE_PX * 16; // bytes between 8-channel chunks
    static constexpr int A_SBO = FLAT ? 128 : BLKW * 16; // bytes between 8-pixel groups (M direction); FLAT: consecutive positions
    static constexpr int A_MAIN_BYTES = NPLANES * PLANE_STRIDE;
    static constexpr int A_TX_BYTES = NPLANES * PLANE_BYTES;
    // extra operand: dense [chunk][128 pixels][8]
    static constexpr int GX = XC == 0 ? 16 : (XC < G ? XC : (XC % G == 0 ? G : 32)); // extra-operand channels per stage / weight slab
    static constexpr int NXS = XC == 0 ? 0 : XC / GX;
    static constexpr int XBOXW = FLAT ? BLKW : 8; // extra-operand box: same position pitch as the accumulator rows
    static constexpr int X_LBO = FLAT ? TR * PITCH * 16 : 128 * 16, X_SBO = 128;
    static constexpr int X_STAGE_BYTES = (GX / 8) * X_LBO;
    static constexpr int A_STAGE_BYTES = ((A_MAIN_BYTES > X_STAGE_BYTES ? A_MAIN_BYTES : X_STAGE_BYTES) + 127) / 128 * 128;
    static constexpr int SLAB_BYTES = G * NC * 2; // one (cin_group, tap) weight slab
    static constexpr int X_SLAB_BYTES = GX * NC * 2;
    static constexpr int W_MAIN_BYTES = NCG * 9 * SLAB_BYTES;
    static constexpr int W_X_BYTES = XC * NC * 2 * XP;
    static constexpr bool RESIDENT = (W_MAIN_BYTES + W_X_BYTES) <= 84 * 1024 && !FLAT && NSPLIT == 1; // (FLAT / channel split: streamed path)
    // weight-slab ring: deep enough that ring depth x MMA time per slab covers the ~2500-cycle L2 -> smem latency of a
    // bulk copy (one slab feeds G/16 MMAs of max(N/2, 32 + N/4) cycles), leaving room for >= MIN_NAS activation stages
    // streamed weights: the activation ring holds exactly two passes' worth of one step (MIN_NAS = 4: one stage per tile
    // of the pair, double buffered); all remaining shared memory goes to the weight-slab ring, whose depth x MMA time per
    // slab must cover the L2 -> smem latency of a bulk copy
    static constexpr int TP = RESIDENT ? 1 : 2;     // tiles per pass: streamed weight slabs are shared by a pair of tiles
    static constexpr int MIN_NAS = 4;
    // taps per ring slot: channel-split layers (tiny batches, one pass per CTA) take all nine taps of a cin group as ONE slot -- one
    // bulk copy and one wait / commit per 9 * G / 16 MMAs.  With a slot per tap the issue loop itself (wait, fence, elect, 2 MMAs, commit:
    // ~100 ns per iteration, 80 iterations per layer3 conv) was the whole 13-20 us of a layer3 conv on one CTU, whatever the split.
    static constexpr int TPS = (NSPLIT > 1 && !CENTER_ONLY) ? NTAPS : 1;
    static constexpr int RSLOT_BYTES = TPS * SLAB_BYTES;
    // ... and they reserve activation stages for (up to 8 of) the TMA boxes of a whole pass before the weight ring takes the rest: with
    // one pass per CTA both streams are bound by round trips to L2 (~1.5 us each), i.e. by how much of the pass is in flight at once
    static constexpr int A_WANT = (NCG + NXS) * (HILO_IN ? 2 : 1) < 8 ? (NCG + NXS) * (HILO_IN ? 2 : 1) : 8;
    static constexpr int A_FIT = 96 * 1024 / A_STAGE_BYTES; // ... within 96 KB
    static constexpr int A_RESERVE = NSPLIT > 1 ? (A_WANT < A_FIT ? A_WANT : (A_FIT > MIN_NAS ? A_FIT : MIN_NAS)) : MIN_NAS;
    static constexpr int NBS_MAX = (231000 - (A_RESERVE > MIN_NAS ? A_RESERVE : MIN_NAS) * A_STAGE_BYTES - COUT * 32 - 4096) / RSLOT_BYTES;
    // channel-split layers serve tiny batches (one pass per CTA): a 12-slab ring makes their weight stream latency-bound (measured on
    // one CTU: 13-18 us per layer3 conv, ~18 GB/s per CTA) -- there the ring takes every slab of a pass, so all copies are in flight at once
    static constexpr int PASS_SLABS = NCG * (NTAPS / TPS) + NXS * XP; // ring slots one pass consumes
    static constexpr int NBS = RESIDENT ? 0 : (NSPLIT > 1 ? (NBS_MAX > PASS_SLABS ? PASS_SLABS : NBS_MAX) : (NBS_MAX > 12 ? 12 : NBS_MAX));
    static_assert(RESIDENT || NBS >= 3 || NBS >= PASS_SLABS, "weight ring too shallow");
    static constexpr int B_BYTES = RESIDENT ? (W_MAIN_BYTES + W_X_BYTES) : NBS * RSLOT_BYTES;
    static_assert(X_SLAB_BYTES <= RSLOT_BYTES, "extra-operand slabs share the ring slots");
    static constexpr int BIAS_BYTES = COUT * 32;    // bias as a K=16 B operand
    static constexpr int ONES_BYTES = 2 * 128 * 16; // matching A operand: k=0,1 -> 1.0, rest 0
    static constexpr int A_BUDGET = 231000 - B_BYTES - BIAS_BYTES - ONES_BYTES;
    static constexpr int NAS_RAW = (A_BUDGET / A_STAGE_BYTES) > 8 ? 8 : (A_BUDGET / A_STAGE_BYTES);
    // A ring depth.  Resident-weight layers run TWO MMA-issuer warps on alternate tiles.  Each issuer owns HALF of the ring
    // (stages [iss * NAS/2, (iss + 1) * NAS/2)), so every full-barrier has exactly one consumer: with one shared ring two
    // issuers wait on the same barriers whenever the depth is not a multiple of a tile pair's stages, and a parity wait
    // cannot tell fill k from fill k + 2 -- the (32 -> 32, stride 2) layer with its 5-stage ring faulted after ~120 tiles
    // per CTA once TMA fills landed out of order (profiles/r01/README.md).
    static constexpr int SPT_ = 1 + NXS; // A stages per tile
    static constexpr int NAS = !RESIDENT ? NAS_RAW : NAS_RAW / 2 * 2;
    static constexpr int NAS_HALF = NAS / 2;
    static constexpr int NACC = (NC <= 128) ? 4 : 2; // TMEM accumulator stages (NACC * ACC_COLS <= 512 columns)
    static constexpr int ACC_COLS = (NC == 96) ? 128 : NC; // accumulator pitch in TMEM columns (power of two)
    // bias: for the smem-operand-bound 32/64-channel layers it is added in the epilogue from registers (an extra MMA
    // would cost 5 % / 3 % of the tile); for 128/256 channels it enters the accumulator through one K=16 MMA
    static constexpr bool BIAS_REG = COUT <= 64;
    // last conv of a stage whose output feeds a prediction head (layer1.1 / layer2.1 / layer3.1 conv2, arch.py:282,288,294):
    // the epilogue also emits per-tile global-average-pool partial sums, so the head never re-reads the activation
    static constexpr bool GAP = (XC == COUT) && COUT >= 64 && !FLAT; // (flat small-map layers: the head pools the stored activation)
    static constexpr int NBAR = 2 * NAS + 2 * (RESIDENT ? 1 : NBS) + 2 * NACC;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + NAS * A_STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_B + B_BYTES;
    static constexpr int OFF_ONES = OFF_BIAS + BIAS_BYTES;
    static constexpr int OFF_BAR = OFF_ONES + ONES_BYTES;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int SMEM_BYTES = OFF_TMEM + 16;
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 32) ? 32 : (NACC * ACC_COLS <= 64 ? 64 : (NACC * ACC_COLS <= 128 ? 128 : (NACC * ACC_COLS <= 256 ? 256 : 512)));
    // warp roles: 0-3 epilogue group 0, 4-7 epilogue group 1 (alternate tiles), 8 MMA issuer, 9 weight loader, 10 activation TMA
    // (resident-weight layers: a second issuer warp takes the odd tiles -- one warp cannot issue a tile's ~220 instructions
    //  in the 750 cycles its 18 N=32 MMAs take, see profiles/r01/README.md)
    static constexpr int W_MMA = 8, W_BLOAD = 9, W_ALOAD = 10, W_MMA2 = 11;
    static constexpr int NTHREADS = 384;
    static constexpr int TILES_PER_IMG = (NB != 1 || FLAT) ? 1 : (HOUT / 16) * (HOUT / 8);
    // layouts of the tensors this conv touches
    static constexpr int IN_PAIR = (NB == 2);
    static constexpr int OUT_PAIR = (HOUT == 8) || (HOUT == 16 && OUT_PAR);
    static constexpr int OHP = OUT_PAR ? HOUT / 2 : HOUT;           // rows (= columns) of one output plane
    static constexpr int ONPL = OUT_PAR ? 4 : 1, ONIMG = OUT_PAIR ? 2 : 1;
    static constexpr int OCHUNK = OHP * ONIMG * OHP * 8;            // elements between output channel chunks
    static constexpr int OUNIT = OCHUNK * (COUT / 8) * ONPL;       // elements per output unit
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(NAS >= 2, "need at least a double-buffered A ring");
    static_assert(A_LBO / 16 < 16384 && COUT * 16 / 16 < 16384, "descriptor field range");
    static_assert(COUT % 32 == 0 && G % 16 == 0 && GX % 16 == 0 && CIN % G == 0 && (XC == 0 || XC % GX == 0), "shape");
    static_assert(NACC * ACC_COLS <= 512, "TMEM columns");
    static_assert(!FLAT || (TR * PITCH <= 128 && NB >= 1 && NB <= 256), "flat tile must fit the 128 accumulator rows");
    static_assert(!CENTER_ONLY || !RESIDENT, "centre-tap layers use the streamed-weight path");
    static_assert(!HILO_IN || (!RESIDENT && STRIP), "hi+lo inputs ride on the tile-pair structure of the streamed-weight path");
    static_assert(NSPLIT == 1 || (!RESIDENT && !BIAS_REG && NC % 32 == 0 && !HILO_IN), "channel split: streamed-weight layers with the bias MMA");
    static constexpr int TSTEP = HILO_IN ? 1 : TP; // tiles per pass
    static_assert(STRIP || (HOUT >= 8 && COUT != 96), "the CTU network has no small maps");
    static_assert(BLKW * 8 <= 256 && PROWS <= 256, "TMA box extents");

    __host__ __device__ static int num_tiles(int nimg) { return ((NB != 1 || FLAT) ? (nimg + NB - 1) / NB : nimg * TILES_PER_IMG) * NSPLIT; } // work items
};

// MLT_CHAIN_TRACE debug: %globaltimer stamp k of this layer, written by CTA 0 only (chain mode, p.trace != nullptr)
__device__ __forceinline__ void layer_stamp(const ConvParams &p, int bid, int k)
{
    if (p.trace != nullptr && bid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[k] = (long long)t;
    }
}

// offset (in 16-byte units) of tap (kh, kw) inside the A stage
template <class C>
__device__ __forceinline__ int tap_offset16(int kh, int kw)
{
    if (C::STRIDE == 1) return (kh - 1 + C::HALO) * C::PITCH + (kw - 1 + C::HALO);
    // stride 2: input row 2*oy + kh - 1 -> odd-row plane for kh != 1; the box of an odd plane starts one row/column
    // earlier (row oy0 - 1), so kh == 2 is one patch row further down
    const int py = (kh != 1), ro = (kh == 2), px = (kw != 1), co = (kw == 2);
    return (py * 2 + px) * (C::PLANE_STRIDE / 16) + ro * C::PITCH + co;
}

// One layer on the tiles bid, bid + nblk, ... of the grid.  CHAIN = false: the body of conv_umma_kernel (one launch per layer: allocates
// and frees its TMEM, PDL-orders itself against the previous kernel).  CHAIN = true: one step of conv_chain_kernel, which runs the
// layers of a small batch back to back inside ONE thread-block cluster: TMEM is allocated once by the caller (chain_tmem), the
// mbarriers are re-initialised per layer and invalidated at its end, and the caller separates the layers by a cluster barrier.
template <class C, bool CHAIN>
__device__ __forceinline__ void conv_layer_body(const ConvParams &p, uint8_t *smem, const uint32_t chain_tmem, const int bid, const int nblk)
{
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint64_t *fullA = bars, *emptyA = bars + C::NAS;
    uint64_t *fullB = bars + 2 * C::NAS;
    uint64_t *emptyB = fullB + (C::RESIDENT ? 1 : C::NBS);
    uint64_t *accFull = emptyB + (C::RESIDENT ? 1 : C::NBS), *accEmpty = accFull + C::NACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = C::num_tiles(p.nimg);
    if (CHAIN && tid == 0) layer_stamp(p, bid, 0);

    if (tid == 0) {
        for (int i = 0; i < C::NAS; i++) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < (C::RESIDENT ? 1 : C::NBS); i++) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < C::NACC; i++) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 128); }
        mbar_fence_init();
    }
    // bias operand pair for the accumulator-initialising MMA:  D = ones[128 x 16] * biasB[COUT x 16]^T
    for (int i = tid; i < C::BIAS_BYTES / 16; i += C::NTHREADS)
        reinterpret_cast<uint4 *>(smem + C::OFF_BIAS)[i] = reinterpret_cast<const uint4 *>(p.bias)[i];
    for (int i = tid; i < C::ONES_BYTES / 16; i += C::NTHREADS)
        reinterpret_cast<uint4 *>(smem + C::OFF_ONES)[i] = i < 128 ? make_uint4(0x3C003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if constexpr (!CHAIN) {
        if (warp == C::W_MMA) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = CHAIN ? chain_tmem : *tmem_slot;
    if (CHAIN && tid == 0) layer_stamp(p, bid, 1);
    const uint32_t sA = smem_u32(smem + C::OFF_A), sB = smem_u32(smem + C::OFF_B);
    // PDL: let the next kernel of the stream start its prologue; everything that touches activations waits for the
    // previous kernel here -- the weight loader does not (weights are constants), so its copies overlap that tail
    if constexpr (!CHAIN) {
        griddep_launch_dependents();
        if (warp != C::W_BLOAD) griddep_wait();
    }

    if (warp < 8) {
        // ======================= epilogue: TMEM (bias, conv, shortcut / residual all accumulated) -> regs -> ReLU -> fp16
        // two groups of 4 warps take alternate tiles, so one group's global stores overlap the other's TMEM reads
        const int grp = warp >> 2, wq = warp & 3;
        const int m = wq * 32 + lane; // accumulator row == TMEM lane == pixel of the tile
        // FLAT: row m is position m of the flattened [row][image][x incl. halo] box; halo positions are not stored
        const int r = C::FLAT ? m / C::PITCH : m / (8 * C::NB), h = C::FLAT ? (m % C::PITCH) / C::BLKW : (m / 8) % C::NB,
                  c = C::FLAT ? m % C::BLKW : m % 8;
        const bool row_ok = !C::FLAT || (r < C::HOUT && c < C::HOUT);
        uint32_t acc_it = grp;
        float bias_r[C::BIAS_REG ? C::COUT : 1];
        if constexpr (C::BIAS_REG) {
#pragma unroll
            for (int j = 0; j < C::COUT; j++) bias_r[j] = __ldg(p.bias_f32 + j);
        }
        for (int tile = bid + grp * nblk; tile < ((C::RESIDENT && (p.dbg & 8)) ? 0 : ntiles); tile += 2 * nblk, acc_it += 2) {
            // p.reverse: this layer walks the images in the opposite direction to the previous one, so that it starts on
            // the activations the previous kernel wrote last (still in the 126 MB L2) -- consecutive layers zig-zag
            const int pitem = p.reverse ? ntiles - 1 - tile : tile;
            const int ptile = pitem / C::NSPLIT, n0 = (pitem % C::NSPLIT) * C::NC; // tile and first output channel of this item
            int img, oy, ox;
            if constexpr (C::NB != 1 || C::FLAT) { img = ptile * C::NB + h; oy = r; ox = c; }
            else {
                img = ptile / C::TILES_PER_IMG;
                const int rem = ptile % C::TILES_PER_IMG;
                oy = (rem / (C::HOUT / 8)) * 16 + r;
                ox = (rem % (C::HOUT / 8)) * 8 + c;
            }
            const bool valid = img < p.nimg && row_ok && !(p.dbg & 2);
            const int unit = C::OUT_PAIR ? img >> 1 : img, sub = C::OUT_PAIR ? img & 1 : 0;
            const int plane = C::OUT_PAR ? (oy & 1) * 2 + (ox & 1) : 0;
            const int yy = C::OUT_PAR ? oy >> 1 : oy, xx = C::OUT_PAR ? ox >> 1 : ox;
            // strip layout: [plane][COUT/8][OHP rows][strip_cap images][OHP][8]
            const size_t ochunk = C::STRIP ? (size_t)C::OHP * p.strip_cap * C::OHP * 8 : (size_t)C::OCHUNK;
            const size_t off = C::STRIP ? (size_t)plane * (ochunk * (C::COUT / 8)) + ((size_t)(yy * p.strip_cap + img) * C::OHP + xx) * 8
                                        : (size_t)unit * C::OUNIT + (size_t)plane * (C::OCHUNK * (C::COUT / 8)) +
                                              (size_t)((yy * C::ONIMG + sub) * C::OHP + xx) * 8;
            const uint32_t acc = acc_it % C::NACC;
            mbar_wait(&accFull[acc], (acc_it / C::NACC) & 1);
            tc_fence_after();
            if (CHAIN && tid == 0) layer_stamp(p, bid, 5);
#pragma unroll(C::BIAS_REG ? 2 : 1)
            for (int c0 = 0; c0 < ((p.dbg & 4) ? 0 : C::NC); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * C::ACC_COLS + c0, v);
                tmem_ld_wait();
                if constexpr (C::BIAS_REG) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + bias_r[c0 + j]);
                }
                const __half2 zero2 = __float2half2_rn(0.0f);
                __half2 hq[16]; // ReLU(fp16(x)) of this pixel's 32 channels: what is stored, and what the pool sums
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const __half2 t = __floats2half2_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                    hq[e] = p.relu ? __hmax2(t, zero2) : t; // max(round(x), 0) == round(max(x, 0))
                }
                if (valid) {
                    __half *op = p.out + off + (size_t)((n0 + c0) / 8) * ochunk;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint4 ov;
                        __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                        for (int e = 0; e < 4; e++) h2[e] = hq[q * 4 + e];
                        *reinterpret_cast<uint4 *>(op + (size_t)q * ochunk) = ov;
                    }
                    if constexpr (C::HILO_OUT) {
                        // lo = fp16(relu(x) - hi), same layout, one whole tensor further on
                        __half *lp = op + ochunk * (C::COUT / 8) * C::ONPL;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            uint4 ov;
                            __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 hf = __half22float2(hq[q * 4 + e]);
                                float x0 = __uint_as_float(v[2 * (q * 4 + e)]), x1 = __uint_as_float(v[2 * (q * 4 + e) + 1]);
                                if (p.relu) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
                                h2[e] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                            }
                            *reinterpret_cast<uint4 *>(lp + (size_t)q * ochunk) = ov;
                        }
                    }
                }
                if constexpr (C::GAP) {
                    // Global-average-pool partial sums of this tile (fixed shuffle tree => bit-reproducible): a transpose-
                    // reduce over the warp's 32 pixels leaves lane l with the sum of channel c0 + l (pair tiles: 16 pixels per
                    // image, two channels per lane); one partial per (tile, image, lane quadrant, channel) goes to HBM and the
                    // head adds the few partials of an image in a fixed order.
                    float gv[32]; // pooled from the fp32 accumulator values (bias and ReLU applied), not from their fp16 roundings
#pragma unroll
                    for (int e = 0; e < 32; e++) gv[e] = p.relu ? fmaxf(__uint_as_float(v[e]), 0.0f) : __uint_as_float(v[e]);
                    auto fold = [&](int off, int nkeep) { // lanes with bit `off` keep the upper half of the list
                        const bool upper = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            if (i < nkeep) {
                                const float send = upper ? gv[i] : gv[i + nkeep], keep = upper ? gv[i + nkeep] : gv[i];
                                gv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                    };
                    float *gp = p.gap_part + ((size_t)(ptile * C::NB + h) * 4 + wq) * C::COUT + n0 + c0;
                    if constexpr (C::NB == 1) {
                        fold(16, 16); fold(8, 8); fold(4, 4); fold(2, 2); fold(1, 1);
                        gp[lane] = gv[0]; // channel bits == lane bits
                    } else {
                        fold(16, 16); fold(4, 8); fold(2, 4); fold(1, 2); // lane bit 3 selects the image: not folded
                        const int cb = ((lane >> 4) & 1) * 16 + (lane & 7) * 2;
                        *reinterpret_cast<float2 *>(gp + cb) = make_float2(gv[0], gv[1]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&accEmpty[acc]);
            if (CHAIN && tid == 0) layer_stamp(p, bid, 6);
        }
    } else if (warp == C::W_MMA || (C::RESIDENT && warp == C::W_MMA2)) {
        // ======================= MMA issuer: the whole warp runs the (uniform) control flow and the waits,
        // one elected lane issues tcgen05.mma / tcgen05.commit
        constexpr uint32_t idesc = umma_idesc_f16(128, C::NC);
        constexpr uint32_t a_hi = umma_desc_hi(C::A_SBO), b_hi = umma_desc_hi(128), x_hi = umma_desc_hi(C::X_SBO);
        const uint32_t ones_lo = umma_desc_lo(smem_u32(smem + C::OFF_ONES), 128 * 16);
        const uint32_t bias_lo = umma_desc_lo(smem_u32(smem + C::OFF_BIAS), C::COUT * 16);
        uint32_t a_it = 0, b_it = 0, acc_it = 0;
        if constexpr (C::RESIDENT) {
            // ---- weights resident (32/64-channel layers): one activation stage (+ one extra-operand stage) per tile.
            // These layers are bound by the MMA stream itself (18 N=32 MMAs = 750 cycles per tile) and ONE warp needs
            // ~930 cycles to issue a tile's ~220 instructions (measured: clock64 per tile, profiles/r01/README.md), so
            // two issuer warps take alternate tiles.
            static_assert(!C::RESIDENT || (C::NCG == 1 && C::NXS <= 1 && C::TP == 1 && C::BIAS_REG), "resident-weight path assumptions");
            constexpr int SPT = 1 + C::NXS; // stages per tile
            static_assert(!C::RESIDENT || (C::NAS_HALF >= SPT && C::NAS % 2 == 0), "each issuer's half ring must hold one tile");
            mbar_wait(&fullB[0], 0);
            auto wait_tile = [&](uint32_t j) {
                if (p.dbg & 8) return; // debug: free-running MMA stream, no handshakes
                mbar_wait(&accEmpty[j % C::NACC], ((j / C::NACC) & 1) ^ 1);
#pragma unroll
                for (int sidx = 0; sidx < SPT; sidx++) {
                    const uint32_t u = (j >> 1) * SPT + sidx; // this issuer's stage counter
                    mbar_wait(&fullA[(j & 1) * C::NAS_HALF + u % C::NAS_HALF], (u / C::NAS_HALF) & 1);
                }
            };
            // two issuer warps: W_MMA takes the even local tiles, W_MMA2 the odd ones (different accumulators and stages;
            // tcgen05.commit tracks the MMAs of the issuing thread, so each tile's barriers see exactly its own MMAs)
            const uint32_t iss = warp == C::W_MMA ? 0u : 1u;
            uint32_t j = iss;
            for (int tile = bid + iss * nblk; tile < ntiles; tile += 2 * nblk, j += 2) {
                wait_tile(j); // each issuer has two tile-times per tile: the wait latency is off the critical path
                if (p.trace != nullptr && bid == 0 && j < 1024 && lane == 0) p.trace[j] = clock64();
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (j % C::NACC) * C::ACC_COLS;
                const uint32_t st0 = iss * C::NAS_HALF + ((j >> 1) * SPT) % C::NAS_HALF;
                const uint32_t a_lo0 = umma_desc_lo(sA + st0 * C::A_STAGE_BYTES, C::A_LBO);
                auto issue_taps = [&](int t0, int t1) {
#pragma unroll
                    for (int tap = t0; tap < t1; tap++) {
                        const uint32_t b_lo0 = umma_desc_lo(sB + tap * C::SLAB_BYTES, C::COUT * 16);
                        const uint32_t a_tap = a_lo0 + tap_offset16<C>(tap / 3, tap % 3);
#pragma unroll
                        for (int ks = 0; ks < C::G / 16; ks++)
                            umma_f16(d_tmem, umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                     umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, (tap == 0 && ks == 0) ? 0u : 1u);
                    }
                };
                if (elect_one_sync()) {
                    issue_taps(0, 9);
                    if (!(p.dbg & 16)) umma_commit(&emptyA[st0]);
                    if constexpr (C::XC > 0) {
                        const uint32_t st1 = iss * C::NAS_HALF + ((j >> 1) * SPT + 1) % C::NAS_HALF;
                        const uint32_t x_lo0 = umma_desc_lo(sA + st1 * C::A_STAGE_BYTES, C::X_LBO);
#pragma unroll
                        for (int part = 0; part < C::XP; part++) {
                            const uint32_t b_lo0 = umma_desc_lo(sB + C::W_MAIN_BYTES + part * C::X_SLAB_BYTES, C::COUT * 16);
#pragma unroll
                            for (int ks = 0; ks < C::GX / 16; ks++)
                                umma_f16(d_tmem, umma_desc_pack(x_lo0 + ks * (2 * C::X_LBO / 16), x_hi),
                                         umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, 1);
                        }
                        if (!(p.dbg & 16)) umma_commit(&emptyA[st1]);
                    }
                    if (!(p.dbg & 16)) umma_commit(&accFull[j % C::NACC]);
                }
                __syncwarp();
            }
            if (p.dbg & 8) { // drain before the TMEM is released
                if (elect_one_sync()) umma_commit(&emptyB[0]);
                __syncwarp();
                mbar_wait(&emptyB[0], 0);
            }
        } else
        // One pass = TP tiles (a PAIR when the weights are streamed): every weight slab fetched from L2 feeds the MMAs of
        // both tiles, which halves the L2 -> smem weight traffic that otherwise bounds the 128/256-channel layers.
        for (int tile = bid; tile < ntiles; tile += C::TSTEP * nblk) {
            // HILO_IN: the pair is (hi box, lo box) of the SAME tile and both halves feed one accumulator
            const int np = C::HILO_IN ? 2 : ((C::TP == 2 && tile + nblk < ntiles) ? 2 : 1);
            const int nacc = C::HILO_IN ? 1 : np; // accumulators of this pass
            uint32_t d_tmem[C::TP];
#pragma unroll
            for (int h = 0; h < C::TP; h++) {
                if (h < nacc) {
                    const uint32_t acc = (acc_it + h) % C::NACC;
                    mbar_wait(&accEmpty[acc], (((acc_it + h) / C::NACC) & 1) ^ 1);
                    d_tmem[h] = tmem_base + acc * C::ACC_COLS;
                }
            }
            if constexpr (C::HILO_IN) d_tmem[C::TP - 1] = d_tmem[0];
            tc_fence_after();
            // accumulator := bias  (ones[128 x 16] x biasB[COUT x 16]^T, hi + lo fp16 split => ~fp32-exact bias)
            if constexpr (!C::BIAS_REG) {
                if (elect_one_sync()) {
                    // (channel split: this pass's rows [split * NC, + NC) of the bias operand, 16 B per row)
                    const uint32_t bias_n0 = (uint32_t)(((p.reverse ? ntiles - 1 - tile : tile) % C::NSPLIT) * C::NC);
#pragma unroll
                    for (int h = 0; h < C::TP; h++)
                        if (h < nacc) umma_f16(d_tmem[h], umma_desc_pack(ones_lo, b_hi), umma_desc_pack(bias_lo + bias_n0, b_hi), idesc, 0);
                }
            }
#pragma unroll 1
            for (int cg = 0; cg < C::NCG; cg++, a_it += np) {
                uint32_t a_lo0[C::TP];
#pragma unroll
                for (int h = 0; h < C::TP; h++) {
                    if (h < np) {
                        const uint32_t st = (a_it + h) % C::NAS;
                        mbar_wait(&fullA[st], ((a_it + h) / C::NAS) & 1); // TMA complete_tx: data visible to the async proxy
                        if (CHAIN && lane == 0 && a_it == 0 && h == 0) layer_stamp(p, bid, 2);
                        a_lo0[h] = umma_desc_lo(sA + st * C::A_STAGE_BYTES, C::A_LBO);
                    }
                }
                tc_fence_after();
                if constexpr (C::RESIDENT) {
                    // weights resident: one elected thread streams all taps of this stage back to back (TP == 1)
                    if (elect_one_sync()) {
#pragma unroll
                        for (int tap = 0; tap < 9; tap++) {
                            const uint32_t b_lo0 = umma_desc_lo(sB + (cg * 9 + tap) * C::SLAB_BYTES, C::COUT * 16);
                            const uint32_t a_tap = a_lo0[0] + tap_offset16<C>(tap / 3, tap % 3);
#pragma unroll
                            for (int ks = 0; ks < C::G / 16; ks++)
                                umma_f16(d_tmem[0], umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                         umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc,
                                         (C::BIAS_REG && tap == 0 && ks == 0) ? (uint32_t)(cg != 0) : 1u);
                        }
                        umma_commit(&emptyA[a_it % C::NAS]);
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int tap = C::TAP0; tap < C::TAP0 + C::NTAPS; tap++) {
                        const int tslot = (tap - C::TAP0) % C::TPS; // position inside the ring slot
                        const uint32_t bs = b_it % C::NBS;
                        if (tslot == 0) {
                            mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                            tc_fence_after();
                            if (CHAIN && lane == 0 && b_it == 0) layer_stamp(p, bid, 3);
                        }
                        if (elect_one_sync()) {
                            const uint32_t b_lo0 = umma_desc_lo(sB + bs * C::RSLOT_BYTES + tslot * C::SLAB_BYTES, C::NC * 16);
#pragma unroll
                            for (int h = 0; h < C::TP; h++) {
                                if (h < np) {
                                    const uint32_t a_tap = a_lo0[h] + tap_offset16<C>(tap / 3, tap % 3);
#pragma unroll
                                    for (int ks = 0; ks < C::G / 16; ks++)
                                        umma_f16(d_tmem[h], umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                                 umma_desc_pack(b_lo0 + ks * (2 * C::NC), b_hi), idesc,
                                                 (C::BIAS_REG && tap == C::TAP0 && ks == 0 && !(C::HILO_IN && h == 1)) ? (uint32_t)(cg != 0) : 1u);
                                }
                            }
                            if (tslot == C::TPS - 1) umma_commit(&emptyB[bs]);
                            if (tap == C::TAP0 + C::NTAPS - 1) {
#pragma unroll
                                for (int h = 0; h < C::TP; h++)
                                    if (h < np) umma_commit(&emptyA[(a_it + h) % C::NAS]);
                            }
                        }
                        __syncwarp();
                        if (tslot == C::TPS - 1) b_it++;
                    }
                }
            }
            if constexpr (C::XC > 0) {
                // extra operand: block input (1x1 stride-2 shortcut conv, or identity residual), GX channels per stage
#pragma unroll 1
                for (int xs = 0; xs < C::NXS; xs++, a_it += np) {
                    uint32_t a_lo0[C::TP];
#pragma unroll
                    for (int h = 0; h < C::TP; h++) {
                        if (h < np) {
                            const uint32_t st = (a_it + h) % C::NAS;
                            mbar_wait(&fullA[st], ((a_it + h) / C::NAS) & 1);
                            a_lo0[h] = umma_desc_lo(sA + st * C::A_STAGE_BYTES, C::X_LBO);
                        }
                    }
                    tc_fence_after();
#pragma unroll
                    for (int part = 0; part < C::XP; part++) { // hi (and lo) weights over the same activation stage
                        uint32_t bs = 0;
                        if constexpr (!C::RESIDENT) {
                            bs = b_it % C::NBS;
                            mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                            tc_fence_after();
                            b_it++;
                        }
                        if (elect_one_sync()) {
                            const uint32_t b_lo0 = C::RESIDENT ? umma_desc_lo(sB + C::W_MAIN_BYTES + (xs * C::XP + part) * C::X_SLAB_BYTES, C::COUT * 16)
                                                               : umma_desc_lo(sB + bs * C::RSLOT_BYTES, C::NC * 16);
#pragma unroll
                            for (int h = 0; h < C::TP; h++) {
                                if (h < np) {
#pragma unroll
                                    for (int ks = 0; ks < C::GX / 16; ks++)
                                        umma_f16(d_tmem[h], umma_desc_pack(a_lo0[h] + ks * (2 * C::X_LBO / 16), x_hi),
                                                 umma_desc_pack(b_lo0 + ks * (2 * C::NC), b_hi), idesc, 1);
                                }
                            }
                            if constexpr (!C::RESIDENT) umma_commit(&emptyB[bs]);
                            if (part == C::XP - 1) {
#pragma unroll
                                for (int h = 0; h < C::TP; h++)
                                    if (h < np) umma_commit(&emptyA[(a_it + h) % C::NAS]);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            if (elect_one_sync()) {
#pragma unroll
                for (int h = 0; h < C::TP; h++)
                    if (h < nacc) umma_commit(&accFull[(acc_it + h) % C::NACC]);
            }
            __syncwarp();
            if (CHAIN && lane == 0) layer_stamp(p, bid, 4);
            acc_it += nacc;
        }
    } else if (warp == C::W_BLOAD) {
        // ======================= weight loader (bulk copies on the TMA engine), one elected lane issues
        const uint8_t *gw = reinterpret_cast<const uint8_t *>(p.w);
        const uint8_t *gx = reinterpret_cast<const uint8_t *>(p.x_w);
        if constexpr (C::RESIDENT) {
            // whole layer stays in shared memory for the lifetime of this persistent CTA
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&fullB[0], C::W_MAIN_BYTES + C::W_X_BYTES);
                for (int s = 0; s < C::NCG * 9; s++)
                    bulk_g2s(sB + s * C::SLAB_BYTES, gw + (size_t)s * C::SLAB_BYTES, C::SLAB_BYTES, &fullB[0]);
                for (int s = 0; s < C::NXS * C::XP; s++)
                    bulk_g2s(sB + C::W_MAIN_BYTES + s * C::X_SLAB_BYTES, gx + (size_t)s * C::X_SLAB_BYTES, C::X_SLAB_BYTES, &fullB[0]);
            }
        } else {
            uint32_t b_it = 0;
            for (int tile = bid; tile < ntiles; tile += C::TSTEP * nblk) { // once per pass (pair of tiles)
                // channel split: the weights are packed [split][cin_group][tap]... / [split][x slab]...
                const int split = (p.reverse ? ntiles - 1 - tile : tile) % C::NSPLIT;
                const uint8_t *gws = gw + (size_t)split * (C::NCG * 9 * C::SLAB_BYTES);
                const uint8_t *gxs = gx + (size_t)split * (C::NXS * C::XP * C::X_SLAB_BYTES);
                constexpr int MAIN_SLOTS = C::NCG * (C::NTAPS / C::TPS); // ring slots of the main operand per pass (TPS taps each)
#pragma unroll 1
                for (int s = 0; s < MAIN_SLOTS + C::NXS * C::XP; s++, b_it++) {
                    const uint32_t bs = b_it % C::NBS;
                    mbar_wait(&emptyB[bs], ((b_it / C::NBS) & 1) ^ 1);
                    if (elect_one_sync()) {
                        const bool is_x = s >= MAIN_SLOTS;
                        const uint32_t bytes = is_x ? C::X_SLAB_BYTES : C::RSLOT_BYTES;
                        // packed [cin_group][9 taps]: first slab of slot s = (s * TPS / NTAPS) * 9 + TAP0 + (s * TPS) % NTAPS, TPS consecutive taps
                        const uint8_t *src = is_x ? gxs + (size_t)(s - MAIN_SLOTS) * C::X_SLAB_BYTES
                                                  : gws + (size_t)(((s * C::TPS) / C::NTAPS) * 9 + C::TAP0 + (s * C::TPS) % C::NTAPS) * C::SLAB_BYTES;
                        mbar_arrive_expect_tx(&fullB[bs], bytes);
                        bulk_g2s(sB + bs * C::RSLOT_BYTES, src, bytes, &fullB[bs]);
                    }
                }
            }
        }
    } else if (warp == C::W_ALOAD) {
        // ======================= activation loader: one TMA box per (tile, cin_group) [four for the parity planes of a
        // stride-2 conv]; the halo / zero padding comes from the TMA out-of-bounds fill
        if (lane == 0) { tma_prefetch_desc(&p.in_map); if (C::XC > 0) tma_prefetch_desc(&p.x_map); }
        uint32_t a_it = 0;
        for (int tile0 = bid; tile0 < ((C::RESIDENT && (p.dbg & 8)) ? 0 : ntiles); tile0 += C::TSTEP * nblk) {
            const int np = C::HILO_IN ? 2 : ((C::TP == 2 && tile0 + nblk < ntiles) ? 2 : 1);
            // stage order of a pass: (step 0, tile 0), (step 0, tile 1), (step 1, tile 0), ... -- what the MMA issuer consumes
#pragma unroll 1
            for (int it = 0; it < C::NCG + C::NXS; it++) {
#pragma unroll 1
                for (int h = 0; h < np; h++, a_it++) {
                    const int ltile = C::HILO_IN ? tile0 : tile0 + h * nblk;
                    const int part = C::HILO_IN ? h : 0; // 0 = hi tensor, 1 = lo tensor (stored right behind it: plane index + planes)
                    const int tile = (p.reverse ? ntiles - 1 - ltile : ltile) / C::NSPLIT;
                    int unit, oy0, ox0;
                    if constexpr (C::NB != 1 || C::FLAT) { unit = tile; oy0 = 0; ox0 = 0; }
                    else {
                        unit = tile / C::TILES_PER_IMG;
                        const int rem = tile % C::TILES_PER_IMG;
                        oy0 = (rem / (C::HOUT / 8)) * 16;
                        ox0 = (rem % (C::HOUT / 8)) * 8;
                    }
                    // resident-weight layers: tile t of this CTA belongs to issuer t & 1, which owns half of the ring
                    uint32_t st = a_it % C::NAS, fill = a_it / C::NAS;
                    if constexpr (C::RESIDENT) {
                        const uint32_t t = a_it / C::SPT_, u = (t >> 1) * C::SPT_ + a_it % C::SPT_;
                        st = (t & 1) * C::NAS_HALF + u % C::NAS_HALF;
                        fill = u / C::NAS_HALF;
                    }
                    mbar_wait(&emptyA[st], (fill & 1) ^ 1);
                    if ((p.dbg & 1) && a_it >= (uint32_t)C::NAS) { // debug: stale smem, no TMA traffic
                        if (elect_one_sync()) mbar_arrive(&fullA[st]);
                    } else if (elect_one_sync()) {
                        const uint32_t abase = sA + st * C::A_STAGE_BYTES;
                        if (it < C::NCG) {
                            mbar_arrive_expect_tx(&fullA[st], C::A_TX_BYTES);
                            // strip layout: the image index is a box coordinate (first image of the tile), the last one the plane
                            const int ci = C::STRIP ? unit * C::NB : 0;
                            if constexpr (C::STRIDE == 1) {
                                tma_load_5d(abase, &p.in_map, (ox0 - C::HALO) * 8, ci, oy0 - C::HALO, it * C::CH, C::STRIP ? part : unit, &fullA[st]);
                            } else {
#pragma unroll
                                for (int pl = 0; pl < 4; pl++)
                                    tma_load_5d(abase + pl * C::PLANE_STRIDE, &p.in_map, (ox0 - (pl & 1)) * 8, ci, oy0 - (pl >> 1),
                                                it * C::CH, C::STRIP ? pl + 4 * part : unit * 4 + pl, &fullA[st]);
                            }
                        } else {
                            mbar_arrive_expect_tx(&fullA[st], C::X_STAGE_BYTES);
                            tma_load_5d(abase, &p.x_map, ox0 * 8, C::STRIP ? unit * C::NB : 0, oy0, (it - C::NCG) * (C::GX / 8),
                                        C::STRIP ? part * p.x_unit_mul : unit * p.x_unit_mul, &fullA[st]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (!CHAIN) {
        if (warp == C::W_MMA) {
            tc_fence_after();
            tmem_dealloc(tmem_base, C::TMEM_COLS);
        }
    } else {
        // every wait of this layer has returned (all roles are past their loops): the barrier objects may be destroyed, the next
        // layer lays its own shared-memory carve-up over them
        if (tid == 0)
            for (int i = 0; i < C::NBAR; i++) mbar_inval(&bars[i]);
        __syncthreads();
    }
}

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) conv_umma_kernel(const __grid_constant__ ConvParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    conv_layer_body<C, false>(p, smem, 0u, (int)blockIdx.x, (int)gridDim.x);
}

} // namespace mlt
