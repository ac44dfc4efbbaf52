// conv_umma.cuh -- implicit-GEMM 3x3 convolution (+ fused 1x1 stride-2 shortcut conv, bias, identity
// residual, ReLU) on tcgen05 tensor cores with TMEM accumulators, for the MLT-CNN residual stack
// (reference graph: mlt_ctu_or_pq_arch.py:52-57 BasicBlock, :273-299 forward; BN folded by pack_weights.py).
//
// GEMM view per tile:  D[128 pixels x COUT] = sum over (cin_group, tap)  A_tap[128 x G] * W_tap[COUT x G]^T
//
// Design (B200-first, not an im2col port):
//  * A tile of M = 128 output pixels is a 16x8 patch of one CTU image (or an 8x8 image pair when the
//    feature map is 8x8).  The INPUT PATCH WITH HALO is loaded ONCE per (tile, cin_group) into shared
//    memory, laid out as [G/8 channel chunks][patch pixels][8 ch] fp16 = UMMA "K-major, no-swizzle"
//    core matrices with the 8 pixels of a patch row 16 B apart.  Each of the nine taps is then just a
//    different START ADDRESS of the same smem patch in the tcgen05.mma operand descriptor
//    (SBO = patch-row pitch, LBO = channel-chunk pitch): no im2col copies, no 9x re-reads from L2.
//    Stride-2 convs keep four parity planes of the input patch so every tap is again a dense window.
//  * Folded weights are pre-packed offline in exactly the smem operand layout; each (cin_group, tap)
//    slab is one contiguous cp.async.bulk (TMA engine) into a ring, or the whole layer is resident in
//    smem for the 32/64-channel layers.
//  * Warp-specialised persistent CTA (1 per SM, 14 warps): 2 x 4 epilogue warps taking alternate tiles
//    (TMEM -> regs -> residual/ReLU -> fp16 NHWC), 1 MMA-issuer warp (one thread issues tcgen05.mma),
//    1 weight-copy warp, 4 A-patch producer warps (cp.async completing on mbarriers, up to 8 stages in
//    flight).  2-4 TMEM accumulator stages overlap the epilogue of tile i with the MMAs of tiles i+1...;
//    the bias enters the accumulator through one extra K=16 MMA (ones x [hi(b), lo(b)]).
#pragma once
#include "ptx.cuh"

namespace mlt {

// Activation tensors are NHWC fp16 with a ONE-PIXEL ZERO HALO: [nimg + 1][H + 2][H + 2][C]; pixel (y, x) lives at
// ((y + 1) * (H + 2) + (x + 1)) * C.  The halo is the convolution's zero padding (kernels never write it) and
// image `nimg` is an all-don't-care spare so that a tile spanning two images never reads out of bounds.
// => the A-patch producers need no bounds checks: every 16-byte piece is `tile_base + per-thread constant`.
struct ConvParams {
    const __half *in;    // haloed NHWC, H = HIN, C = CIN
    const __half *w;     // packed [CIN/G][9][G/8][COUT][8]
    const __half *bias;  // tcgen05 bias operand [2][COUT][8] fp16: k=0 -> hi(b), k=1 -> lo(b), rest 0 (pack_weights.py)
    const __half *sc_in; // haloed NHWC, H = 2*HOUT, C = CSC   (block input; 1x1 stride-2 shortcut conv)
    const __half *sc_w;  // packed [CSC/8][COUT][8]
    const __half *res;   // haloed NHWC, H = HOUT, C = COUT identity residual, or nullptr
    __half *out;         // haloed NHWC, H = HOUT, C = COUT
    int nimg;
    int relu;
    int dbg;          // debug timing knobs (MLT_DEBUG_FLAGS; results invalid when != 0): 1 = producers skip the copies,
                      // 2 = epilogue skips global loads/stores, 4 = epilogue skips the TMEM reads too
    long long *trace; // debug (MLT_TRACE_LAYER): [role 0..3][64 tiles][4] clock64() stamps of CTA 0, or nullptr
};

// role: 0 = producer thread 0, 1 = MMA issuer, 2 = epilogue group 0, 3 = epilogue group 1
__device__ __forceinline__ void trace_stamp(long long *trace, int role, uint32_t local_tile, int slot)
{
    if (trace != nullptr && blockIdx.x == 0 && local_tile < 64) trace[(role * 64 + local_tile) * 4 + slot] = clock64();
}

template <int CIN_, int COUT_, int STRIDE_, int HOUT_, int CSC_>
struct ConvCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, STRIDE = STRIDE_, HOUT = HOUT_, CSC = CSC_;
    static constexpr int HIN = HOUT * STRIDE;
    static constexpr int G = (STRIDE == 2) ? 32 : (CIN < 64 ? CIN : 64); // input channels per A stage
    static constexpr int NCG = CIN / G;
    static constexpr int NB = (HOUT == 8) ? 2 : 1; // images per tile
    static constexpr int TR = 128 / (8 * NB);      // tile rows (16 or 8); tile = TR x (NB * 8) pixels
    static constexpr int BLKW = (STRIDE == 1) ? 10 : 9; // patch pixels per 8-wide block (halo included)
    static constexpr int PITCH = BLKW * NB;
    static constexpr int PROWS = (STRIDE == 1) ? TR + 2 : TR + 1;
    static constexpr int NPLANES = (STRIDE == 1) ? 1 : 4;
    static constexpr int PLANE_PX = PROWS * PITCH;
    static constexpr int PATCH_PX = NPLANES * PLANE_PX;
    static constexpr int A_LBO = PATCH_PX * 16;  // bytes between 8-channel chunks
    static constexpr int A_SBO = BLKW * 16;      // bytes between 8-pixel groups (M direction)
    static constexpr int SC_LBO = 128 * 16, SC_SBO = 128;
    static constexpr int A_MAIN_BYTES = (G / 8) * A_LBO;
    static constexpr int A_SC_BYTES = CSC * 256;
    static constexpr int A_STAGE_BYTES = ((A_MAIN_BYTES > A_SC_BYTES ? A_MAIN_BYTES : A_SC_BYTES) + 127) / 128 * 128;
    static constexpr int SLAB_BYTES = G * COUT * 2; // one (cin_group, tap) weight slab
    static constexpr int GS = CSC == 0 ? 16 : (CSC < G ? CSC : G); // shortcut channels per slab
    static constexpr int NSC_SLABS = CSC == 0 ? 0 : CSC / GS;
    static constexpr int SC_SLAB_BYTES = GS * COUT * 2;
    static constexpr int W_MAIN_BYTES = NCG * 9 * SLAB_BYTES;
    static constexpr int W_SC_BYTES = CSC * COUT * 2;
    static constexpr bool RESIDENT = (W_MAIN_BYTES + W_SC_BYTES) <= 80 * 1024;
    static constexpr int NBS = RESIDENT ? 0 : (SLAB_BYTES >= 32 * 1024 ? 4 : (SLAB_BYTES >= 16 * 1024 ? 4 : 6));
    static constexpr int B_BYTES = RESIDENT ? (W_MAIN_BYTES + W_SC_BYTES) : NBS * SLAB_BYTES;
    // A ring: as deep as shared memory allows (<= 8): the producers never block on memory, so NAS stages of
    // cp.async are in flight per SM -- this is what hides the L2/HBM latency.
    static constexpr int BIAS_BYTES = COUT * 32;      // bias as a K=16 B operand
    static constexpr int ONES_BYTES = 2 * 128 * 16;   // matching A operand: k=0,1 -> 1.0, rest 0
    static constexpr int A_BUDGET = 231000 - B_BYTES - BIAS_BYTES - ONES_BYTES;
    static constexpr int NAS = (A_BUDGET / A_STAGE_BYTES) > 8 ? 8 : (A_BUDGET / A_STAGE_BYTES);
    static constexpr int NACC = (COUT <= 128) ? 4 : 2; // TMEM accumulator stages (NACC * COUT <= 512 columns)
    static constexpr int NBAR = 2 * NAS + 2 * (RESIDENT ? 1 : NBS) + 2 * NACC;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + NAS * A_STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_B + B_BYTES;
    static constexpr int OFF_ONES = OFF_BIAS + BIAS_BYTES;
    static constexpr int OFF_BAR = OFF_ONES + ONES_BYTES;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int SMEM_BYTES = OFF_TMEM + 16;
    static constexpr int TMEM_COLS = (NACC * COUT <= 32) ? 32 : (NACC * COUT <= 64 ? 64 : (NACC * COUT <= 128 ? 128 : (NACC * COUT <= 256 ? 256 : 512)));
    // warp roles: 0-3 epilogue group 0, 4-7 epilogue group 1 (alternate tiles), 8 MMA issuer, 9 weight loader, 10-13 A producers
    static constexpr int W_MMA = 8, W_BLOAD = 9, W_PROD = 10;
    static constexpr int TILES_PER_IMG = (NB == 2) ? 1 : (HOUT / 16) * (HOUT / 8);
    static constexpr int NTHREADS = 448;
    static constexpr int HPI = HIN + 2, HPO = HOUT + 2, HPS = 2 * HOUT + 2; // haloed extents (in / out / shortcut in)
    static constexpr int IMG_IN = HPI * HPI * CIN, IMG_OUT = HPO * HPO * COUT, IMG_SC = HPS * HPS * CSC;
    static constexpr int CH = G / 8;                                // 16-byte chunks per pixel per A stage
    static constexpr int KIT = (PATCH_PX * CH + 127) / 128;         // cp.async per producer thread per main stage
    static constexpr int SCH = CSC / 8;
    static constexpr int KIT_SC = SCH;                              // 128 * SCH pieces / 128 threads
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(NAS >= 2, "need at least a double-buffered A ring");
    static_assert(A_LBO / 16 < 16384 && COUT * 16 / 16 < 16384, "descriptor field range");
    static_assert(COUT % 32 == 0 && G % 16 == 0, "shape");

    __host__ __device__ static int num_tiles(int nimg) { return NB == 2 ? (nimg + 1) / 2 : nimg * TILES_PER_IMG; }
};

// patch-pixel offset (in pixels) of tap (kh, kw) inside the A stage
template <class C>
__device__ __forceinline__ int tap_offset_px(int kh, int kw)
{
    if (C::STRIDE == 1) return kh * C::PITCH + kw;
    // stride 2: input row 2*oy + kh - 1  -> parity plane (kh != 1), local row offset (kh == 2)
    const int py = (kh != 1), ro = (kh == 2), px = (kw != 1), co = (kw == 2);
    return (py * 2 + px) * C::PLANE_PX + ro * C::PITCH + co;
}

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) conv_umma_kernel(const ConvParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint64_t *fullA = bars, *emptyA = bars + C::NAS;
    uint64_t *fullB = bars + 2 * C::NAS;
    uint64_t *emptyB = fullB + (C::RESIDENT ? 1 : C::NBS);
    uint64_t *accFull = emptyB + (C::RESIDENT ? 1 : C::NBS), *accEmpty = accFull + C::NACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = C::num_tiles(p.nimg);

    if (tid == 0) {
        for (int i = 0; i < C::NAS; i++) { mbar_init(&fullA[i], 128); mbar_init(&emptyA[i], 1); }
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
    if (warp == C::W_MMA) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sA = smem_u32(smem + C::OFF_A), sB = smem_u32(smem + C::OFF_B);

    if (warp < 8) {
        // ======================= epilogue: TMEM (bias already accumulated) -> regs (+residual) -> fp16, ReLU -> haloed NHWC
        // two groups of 4 warps take alternate tiles, so one group's global stores overlap the other's TMEM reads
        const int grp = warp >> 2, wq = warp & 3;
        const int m = wq * 32 + lane; // accumulator row == TMEM lane == pixel of the tile
        const int r = m / (8 * C::NB), h = (m / 8) % C::NB, c = m % 8;
        uint32_t acc_it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < ntiles; tile += 2 * gridDim.x, acc_it += 2) {
            int img, oy, ox;
            if (C::NB == 2) { img = tile * 2 + h; oy = r; ox = c; }
            else {
                img = tile / C::TILES_PER_IMG;
                const int rem = tile % C::TILES_PER_IMG;
                oy = (rem / (C::HOUT / 8)) * 16 + r;
                ox = (rem % (C::HOUT / 8)) * 8 + c;
            }
            const bool valid = img < p.nimg && !(p.dbg & 2);
            const size_t off = (size_t)img * C::IMG_OUT + (size_t)((oy + 1) * C::HPO + ox + 1) * C::COUT;
            const uint32_t acc = acc_it % C::NACC;
            // identity residual: issue the loads of the first 32 channels BEFORE blocking on the accumulator
            uint4 rv[4];
            const bool has_res = p.res != nullptr && valid;
            if (has_res) {
                const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + off);
#pragma unroll
                for (int q = 0; q < 4; q++) rv[q] = __ldg(rp + q);
            }
            if (wq == 0 && lane == 0) trace_stamp(p.trace, 2 + grp, acc_it, 0);
            mbar_wait(&accFull[acc], (acc_it / C::NACC) & 1);
            tc_fence_after();
            if (wq == 0 && lane == 0) trace_stamp(p.trace, 2 + grp, acc_it, 1);
#pragma unroll 1
            for (int c0 = 0; c0 < ((p.dbg & 4) ? 0 : C::COUT); c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * C::COUT + c0, v);
                uint4 rn[4];
                if (has_res && c0 + 32 < C::COUT) { // next chunk's residual while this one is processed
                    const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + off + c0 + 32);
#pragma unroll
                    for (int q = 0; q < 4; q++) rn[q] = __ldg(rp + q);
                }
                tmem_ld_wait();
                if (valid) {
                    if (has_res) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const __half2 *h2 = reinterpret_cast<const __half2 *>(&rv[q]);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 t = __half22float2(h2[e]);
                                v[q * 8 + e * 2] = __float_as_uint(__uint_as_float(v[q * 8 + e * 2]) + t.x);
                                v[q * 8 + e * 2 + 1] = __float_as_uint(__uint_as_float(v[q * 8 + e * 2 + 1]) + t.y);
                            }
                        }
                    }
                    const __half2 zero2 = __float2half2_rn(0.0f);
                    uint4 *op = reinterpret_cast<uint4 *>(p.out + off + c0);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint4 ov;
                        __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const __half2 t = __floats2half2_rn(__uint_as_float(v[q * 8 + e * 2]), __uint_as_float(v[q * 8 + e * 2 + 1]));
                            h2[e] = p.relu ? __hmax2(t, zero2) : t; // max(round(x), 0) == round(max(x, 0))
                        }
                        op[q] = ov;
                    }
                }
                if (has_res && c0 + 32 < C::COUT) {
#pragma unroll
                    for (int q = 0; q < 4; q++) rv[q] = rn[q];
                }
            }
            tc_fence_before();
            mbar_arrive(&accEmpty[acc]);
            if (wq == 0 && lane == 0) trace_stamp(p.trace, 2 + grp, acc_it, 2);
        }
    } else if (warp == C::W_MMA) {
        // ======================= MMA issuer: the whole warp runs the (uniform) control flow and the waits,
        // one elected lane issues tcgen05.mma / tcgen05.commit
        constexpr uint32_t idesc = umma_idesc_f16(128, C::COUT);
        constexpr uint32_t a_hi = umma_desc_hi(C::A_SBO), b_hi = umma_desc_hi(128), s_hi = umma_desc_hi(C::SC_SBO);
        const uint32_t ones_lo = umma_desc_lo(smem_u32(smem + C::OFF_ONES), 128 * 16);
        const uint32_t bias_lo = umma_desc_lo(smem_u32(smem + C::OFF_BIAS), C::COUT * 16);
        uint32_t a_it = 0, b_it = 0, acc_it = 0;
        if constexpr (C::RESIDENT) { mbar_wait(&fullB[0], 0); tc_fence_after(); }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, acc_it++) {
            const uint32_t acc = acc_it % C::NACC;
            if (lane == 0) trace_stamp(p.trace, 1, acc_it, 0);
            mbar_wait(&accEmpty[acc], ((acc_it / C::NACC) & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) trace_stamp(p.trace, 1, acc_it, 1);
            const uint32_t d_tmem = tmem_base + acc * C::COUT;
            // accumulator := bias  (ones[128 x 16] x biasB[COUT x 16]^T, hi + lo fp16 split => ~fp32-exact bias)
            if (elect_one_sync()) umma_f16(d_tmem, umma_desc_pack(ones_lo, b_hi), umma_desc_pack(bias_lo, b_hi), idesc, 0);
#pragma unroll 1
            for (int cg = 0; cg < C::NCG; cg++, a_it++) {
                const uint32_t st = a_it % C::NAS;
                // cp.async data is published by the copy unit's own mbarrier arrive (cp.async.mbarrier.arrive):
                // observing the phase flip is the ordering point for the tensor-core reads, as in CUTLASS's
                // sm100 cp.async + UMMA mainloop -- no generic->async proxy fence on this critical path
                mbar_wait(&fullA[st], (a_it / C::NAS) & 1);
                tc_fence_after();
                if (cg == 0 && lane == 0) trace_stamp(p.trace, 1, acc_it, 2);
                const uint32_t a_lo0 = umma_desc_lo(sA + st * C::A_STAGE_BYTES, C::A_LBO);
                if constexpr (C::RESIDENT) {
                    if (elect_one_sync()) {
#pragma unroll
                        for (int tap = 0; tap < 9; tap++) {
                            const uint32_t b_lo0 = umma_desc_lo(sB + (cg * 9 + tap) * C::SLAB_BYTES, C::COUT * 16);
                            const uint32_t a_tap = a_lo0 + tap_offset_px<C>(tap / 3, tap % 3); // 16-byte units
#pragma unroll
                            for (int ks = 0; ks < C::G / 16; ks++)
                                umma_f16(d_tmem, umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                         umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, 1);
                        }
                        umma_commit(&emptyA[st]);
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < 9; tap++, b_it++) {
                        const uint32_t bs = b_it % C::NBS;
                        mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint32_t b_lo0 = umma_desc_lo(sB + bs * C::SLAB_BYTES, C::COUT * 16);
                            const uint32_t a_tap = a_lo0 + tap_offset_px<C>(tap / 3, tap % 3);
#pragma unroll
                            for (int ks = 0; ks < C::G / 16; ks++)
                                umma_f16(d_tmem, umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                         umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, 1);
                            umma_commit(&emptyB[bs]);
                            if (tap == 8) umma_commit(&emptyA[st]);
                        }
                    }
                }
            }
            if constexpr (C::CSC > 0) {
                const uint32_t st = a_it % C::NAS;
                mbar_wait(&fullA[st], (a_it / C::NAS) & 1);
                tc_fence_after();
                const uint32_t a_lo0 = umma_desc_lo(sA + st * C::A_STAGE_BYTES, C::SC_LBO);
#pragma unroll
                for (int sl = 0; sl < C::NSC_SLABS; sl++) {
                    uint32_t bs = 0;
                    if constexpr (!C::RESIDENT) {
                        bs = b_it % C::NBS;
                        mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                        tc_fence_after();
                        b_it++;
                    }
                    if (elect_one_sync()) {
                        const uint32_t b_lo0 = C::RESIDENT ? umma_desc_lo(sB + C::W_MAIN_BYTES + sl * C::SC_SLAB_BYTES, C::COUT * 16)
                                                           : umma_desc_lo(sB + bs * C::SLAB_BYTES, C::COUT * 16);
#pragma unroll
                        for (int ks = 0; ks < C::GS / 16; ks++)
                            umma_f16(d_tmem, umma_desc_pack(a_lo0 + (sl * (C::GS / 8) + 2 * ks) * (C::SC_LBO / 16), s_hi),
                                     umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, 1);
                        if constexpr (!C::RESIDENT) umma_commit(&emptyB[bs]);
                        if (sl == C::NSC_SLABS - 1) umma_commit(&emptyA[st]);
                    }
                }
                a_it++;
            }
            if (elect_one_sync()) umma_commit(&accFull[acc]);
            if (lane == 0) trace_stamp(p.trace, 1, acc_it, 3);
        }
    } else if (warp == C::W_BLOAD) {
        // ======================= weight loader (bulk copies on the TMA engine), one elected lane issues
        const uint8_t *gw = reinterpret_cast<const uint8_t *>(p.w);
        const uint8_t *gsc = reinterpret_cast<const uint8_t *>(p.sc_w);
        if constexpr (C::RESIDENT) {
            // whole layer stays in shared memory for the lifetime of this persistent CTA
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(&fullB[0], C::W_MAIN_BYTES + C::W_SC_BYTES);
                for (int s = 0; s < C::NCG * 9; s++)
                    bulk_g2s(sB + s * C::SLAB_BYTES, gw + (size_t)s * C::SLAB_BYTES, C::SLAB_BYTES, &fullB[0]);
                for (int s = 0; s < C::NSC_SLABS; s++)
                    bulk_g2s(sB + C::W_MAIN_BYTES + s * C::SC_SLAB_BYTES, gsc + (size_t)s * C::SC_SLAB_BYTES,
                             C::SC_SLAB_BYTES, &fullB[0]);
            }
        } else {
            uint32_t b_it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
                for (int s = 0; s < C::NCG * 9 + C::NSC_SLABS; s++, b_it++) {
                    const uint32_t bs = b_it % C::NBS;
                    mbar_wait(&emptyB[bs], ((b_it / C::NBS) & 1) ^ 1);
                    if (elect_one_sync()) {
                        const bool is_sc = s >= C::NCG * 9;
                        const uint32_t bytes = is_sc ? C::SC_SLAB_BYTES : C::SLAB_BYTES;
                        const uint8_t *src = is_sc ? gsc + (size_t)(s - C::NCG * 9) * C::SC_SLAB_BYTES : gw + (size_t)s * C::SLAB_BYTES;
                        mbar_arrive_expect_tx(&fullB[bs], bytes);
                        bulk_g2s(sB + bs * C::SLAB_BYTES, src, bytes, &fullB[bs]);
                    }
                }
            }
        }
    } else {
        // ======================= A-patch producers (128 threads, cp.async 16 B)
        // Thread pt owns chunk j = pt % CH of pixels q = pt / CH + k * (128 / CH): the smem destination is affine
        // in k and the global source is `tile base + rel[k]` with rel[] computed ONCE per kernel -- two
        // instructions per 16 bytes in the steady state, no bounds checks (zero halo in global memory).
        const int pt = tid - C::W_PROD * 32;
        constexpr int CH = C::CH, PXSTEP = 128 / CH;
        const int j = pt % CH, q0 = pt / CH;
        int rel[C::KIT];
#pragma unroll
        for (int k = 0; k < C::KIT; k++) {
            const int q = q0 + k * PXSTEP;
            int r = -1;
            if (q < C::PATCH_PX) {
                if (C::STRIDE == 1) {
                    const int pr = q / C::PITCH, rem = q % C::PITCH, hb = rem / C::BLKW, pc = rem % C::BLKW;
                    r = (pr * C::HPI + pc) * C::CIN + hb * C::IMG_IN + j * 8; // (y+1, x+1) = (oy0 + pr, ox0 + pc)
                } else {
                    const int plane = q / C::PLANE_PX, r2 = q % C::PLANE_PX;
                    const int ip = r2 / C::PITCH, rem = r2 % C::PITCH, hb = rem / C::BLKW, jp = rem % C::BLKW;
                    const int py = plane >> 1, px = plane & 1;
                    if (ip < C::TR + py && jp < 8 + px) // y = 2*(oy0+ip-py)+py, x likewise; +1 for the halo
                        r = ((2 * ip - py + 1) * C::HPI + (2 * jp - px + 1)) * C::CIN + hb * C::IMG_IN + j * 8;
                }
            }
            rel[k] = r;
        }
        int rel_sc[C::KIT_SC > 0 ? C::KIT_SC : 1];
        if constexpr (C::CSC > 0) {
#pragma unroll
            for (int k = 0; k < C::KIT_SC; k++) {
                const int s = pt + 128 * k, m = s / C::SCH, jc = s % C::SCH;
                const int r = m / (8 * C::NB), hb = (m / 8) % C::NB, c = m % 8;
                rel_sc[k] = ((2 * r + 1) * C::HPS + 2 * c + 1) * C::CSC + hb * C::IMG_SC + jc * 8;
            }
        }
        const uint32_t dst0 = j * C::A_LBO + q0 * 16;
        uint32_t a_it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int img0, oy0, ox0;
            if (C::NB == 2) { img0 = tile * 2; oy0 = 0; ox0 = 0; }
            else {
                img0 = tile / C::TILES_PER_IMG;
                const int rem = tile % C::TILES_PER_IMG;
                oy0 = (rem / (C::HOUT / 8)) * 16;
                ox0 = (rem % (C::HOUT / 8)) * 8;
            }
            const __half *tin = p.in + (size_t)img0 * C::IMG_IN + (size_t)((C::STRIDE * oy0) * C::HPI + C::STRIDE * ox0) * C::CIN;
            const uint32_t ltile = (uint32_t)((tile - blockIdx.x) / gridDim.x);
            for (int it = 0; it < C::NCG + (C::CSC > 0 ? 1 : 0); it++, a_it++) {
                const uint32_t st = a_it % C::NAS;
                if (pt == 0 && it == 0) trace_stamp(p.trace, 0, ltile, 0);
                mbar_wait(&emptyA[st], ((a_it / C::NAS) & 1) ^ 1);
                if (pt == 0 && it == 0) trace_stamp(p.trace, 0, ltile, 1);
                const uint32_t abase = sA + st * C::A_STAGE_BYTES;
                if (it < C::NCG) {
                    const __half *src = tin + it * C::G;
#pragma unroll
                    for (int k = 0; k < C::KIT; k++)
                        if (rel[k] >= 0 && !(p.dbg & 1)) cp_async16(abase + dst0 + k * (PXSTEP * 16), src + rel[k], true);
                } else if constexpr (C::CSC > 0) {
                    // shortcut operand: block input sampled at (2*oy, 2*ox), rows in accumulator order
                    const __half *src = p.sc_in + (size_t)img0 * C::IMG_SC + (size_t)((2 * oy0) * C::HPS + 2 * ox0) * C::CSC;
#pragma unroll
                    for (int k = 0; k < C::KIT_SC; k++) {
                        const int s = pt + 128 * k;
                        cp_async16(abase + (s % C::SCH) * C::SC_LBO + (s / C::SCH) * 16, src + rel_sc[k], true);
                    }
                }
                // completion is signalled by the copy unit itself: this thread never waits on memory, so up to
                // NAS stages are in flight (the MMA thread issues the generic->async proxy fence after its wait)
                cp_async_mbar_arrive_noinc(&fullA[st]);
                if (pt == 0 && it == 0) trace_stamp(p.trace, 0, ltile, 2);
            }
        }
        cp_async_wait_all(); // nothing may still be landing in shared memory when the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (warp == C::W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

} // namespace mlt
