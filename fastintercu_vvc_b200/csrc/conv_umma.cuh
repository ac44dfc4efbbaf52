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
//  * Warp-specialised persistent CTA (1 per SM): 4 epilogue warps (TMEM -> regs -> bias/residual/ReLU
//    -> fp16 NHWC), 1 MMA-issuer warp (one thread issues tcgen05.mma), 1 weight-copy warp, 4 A-patch
//    producer warps (cp.async with zero-fill = conv padding).  Double-buffered TMEM accumulators
//    overlap the epilogue of tile i with the MMAs of tile i+1.
#pragma once
#include "ptx.cuh"

namespace mlt {

struct ConvParams {
    const __half *in;    // NHWC [nimg][HIN][HIN][CIN]
    const __half *w;     // packed [CIN/G][9][G/8][COUT][8]
    const float *bias;   // [COUT]  (already includes the shortcut's folded BN bias when CSC > 0)
    const __half *sc_in; // NHWC [nimg][2*HOUT][2*HOUT][CSC]   (block input; 1x1 stride-2 shortcut conv)
    const __half *sc_w;  // packed [CSC/8][COUT][8]
    const __half *res;   // NHWC [nimg][HOUT][HOUT][COUT] identity residual, or nullptr
    __half *out;         // NHWC [nimg][HOUT][HOUT][COUT]
    int nimg;
    int relu;
};

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
    static constexpr int NAS = (A_STAGE_BYTES > 36 * 1024) ? 2 : (A_STAGE_BYTES > 16 * 1024 ? 3 : 4); // A stages
    static constexpr int NBS = RESIDENT ? 0 : (SLAB_BYTES >= 32 * 1024 ? 4 : (SLAB_BYTES >= 16 * 1024 ? 6 : 8));
    static constexpr int B_BYTES = RESIDENT ? (W_MAIN_BYTES + W_SC_BYTES) : NBS * SLAB_BYTES;
    static constexpr int NBAR = 2 * NAS + 2 * (RESIDENT ? 1 : NBS) + 4;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + NAS * A_STAGE_BYTES;
    static constexpr int OFF_BIAS = OFF_B + B_BYTES;
    static constexpr int OFF_BAR = OFF_BIAS + COUT * 4;
    static constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
    static constexpr int SMEM_BYTES = OFF_TMEM + 16;
    static constexpr int TMEM_COLS = (2 * COUT <= 32) ? 32 : (2 * COUT <= 64 ? 64 : (2 * COUT <= 128 ? 128 : (2 * COUT <= 256 ? 256 : 512)));
    static constexpr int TILES_PER_IMG = (NB == 2) ? 1 : (HOUT / 16) * (HOUT / 8);
    static constexpr int NTHREADS = 320;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
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
    uint64_t *accFull = emptyB + (C::RESIDENT ? 1 : C::NBS), *accEmpty = accFull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::OFF_TMEM);
    float *s_bias = reinterpret_cast<float *>(smem + C::OFF_BIAS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = C::num_tiles(p.nimg);

    if (tid == 0) {
        for (int i = 0; i < C::NAS; i++) { mbar_init(&fullA[i], 128); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < (C::RESIDENT ? 1 : C::NBS); i++) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&accFull[i], 1); mbar_init(&accEmpty[i], 128); }
        mbar_fence_init();
    }
    for (int i = tid; i < C::COUT; i += C::NTHREADS) s_bias[i] = p.bias[i];
    if (warp == 4) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sA = smem_u32(smem + C::OFF_A), sB = smem_u32(smem + C::OFF_B);

    if (warp < 4) {
        // ======================= epilogue: TMEM -> regs -> bias (+residual) (+ReLU) -> fp16 NHWC
        const int m = warp * 32 + lane; // accumulator row == TMEM lane == pixel of the tile
        const int r = m / (8 * C::NB), h = (m / 8) % C::NB, c = m % 8;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, acc_it++) {
            int img, oy, ox;
            if (C::NB == 2) { img = tile * 2 + h; oy = r; ox = c; }
            else {
                img = tile / C::TILES_PER_IMG;
                const int rem = tile % C::TILES_PER_IMG;
                oy = (rem / (C::HOUT / 8)) * 16 + r;
                ox = (rem % (C::HOUT / 8)) * 8 + c;
            }
            const bool valid = img < p.nimg;
            const size_t off = (((size_t)img * C::HOUT + oy) * C::HOUT + ox) * C::COUT;
            const uint32_t acc = acc_it & 1;
            mbar_wait(&accFull[acc], (acc_it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < C::COUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * C::COUT + c0, v);
                tmem_ld_wait();
                if (valid) {
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) f[i] = __uint_as_float(v[i]) + s_bias[c0 + i];
                    if (p.res != nullptr) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + off + c0);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint4 rv = __ldg(rp + q);
                            const __half2 *h2 = reinterpret_cast<const __half2 *>(&rv);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 t = __half22float2(h2[e]);
                                f[q * 8 + e * 2] += t.x;
                                f[q * 8 + e * 2 + 1] += t.y;
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; i++) f[i] = fmaxf(f[i], 0.0f);
                    }
                    uint4 *op = reinterpret_cast<uint4 *>(p.out + off + c0);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint4 ov;
                        __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                        for (int e = 0; e < 4; e++) h2[e] = __floats2half2_rn(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1]);
                        op[q] = ov;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&accEmpty[acc]);
        }
    } else if (warp == 4) {
        // ======================= MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, C::COUT);
            uint32_t a_it = 0, b_it = 0, acc_it = 0;
            if constexpr (C::RESIDENT) { mbar_wait(&fullB[0], 0); tc_fence_after(); }
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, acc_it++) {
                const uint32_t acc = acc_it & 1;
                mbar_wait(&accEmpty[acc], ((acc_it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::COUT;
                uint32_t accum = 0;
#pragma unroll 1
                for (int cg = 0; cg < C::NCG; cg++, a_it++) {
                    const uint32_t st = a_it % C::NAS;
                    mbar_wait(&fullA[st], (a_it / C::NAS) & 1);
                    tc_fence_after();
                    const uint32_t abase = sA + st * C::A_STAGE_BYTES;
#pragma unroll 1
                    for (int tap = 0; tap < 9; tap++) {
                        uint32_t bbase;
                        uint32_t bs = 0;
                        if constexpr (C::RESIDENT) bbase = sB + (cg * 9 + tap) * C::SLAB_BYTES;
                        else {
                            bs = b_it % C::NBS;
                            mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                            tc_fence_after();
                            bbase = sB + bs * C::SLAB_BYTES;
                        }
                        const uint32_t a_tap = abase + tap_offset_px<C>(tap / 3, tap % 3) * 16;
#pragma unroll
                        for (int ks = 0; ks < C::G / 16; ks++) {
                            const uint64_t ad = umma_desc_kmajor_noswz(a_tap + 2 * ks * C::A_LBO, C::A_LBO, C::A_SBO);
                            const uint64_t bd = umma_desc_kmajor_noswz(bbase + 2 * ks * C::COUT * 16, C::COUT * 16, 128);
                            umma_f16(d_tmem, ad, bd, idesc, accum);
                            accum = 1;
                        }
                        if constexpr (!C::RESIDENT) { umma_commit(&emptyB[bs]); b_it++; }
                    }
                    umma_commit(&emptyA[st]);
                }
                if (C::CSC > 0) {
                    const uint32_t st = a_it % C::NAS;
                    mbar_wait(&fullA[st], (a_it / C::NAS) & 1);
                    tc_fence_after();
                    const uint32_t abase = sA + st * C::A_STAGE_BYTES;
#pragma unroll 1
                    for (int sl = 0; sl < C::NSC_SLABS; sl++) {
                        uint32_t bbase;
                        uint32_t bs = 0;
                        if constexpr (C::RESIDENT) bbase = sB + C::W_MAIN_BYTES + sl * C::SC_SLAB_BYTES;
                        else {
                            bs = b_it % C::NBS;
                            mbar_wait(&fullB[bs], (b_it / C::NBS) & 1);
                            tc_fence_after();
                            bbase = sB + bs * C::SLAB_BYTES;
                        }
#pragma unroll
                        for (int ks = 0; ks < C::GS / 16; ks++) {
                            const uint64_t ad = umma_desc_kmajor_noswz(abase + (sl * (C::GS / 8) + 2 * ks) * C::SC_LBO, C::SC_LBO, C::SC_SBO);
                            const uint64_t bd = umma_desc_kmajor_noswz(bbase + 2 * ks * C::COUT * 16, C::COUT * 16, 128);
                            umma_f16(d_tmem, ad, bd, idesc, accum);
                            accum = 1;
                        }
                        if constexpr (!C::RESIDENT) { umma_commit(&emptyB[bs]); b_it++; }
                    }
                    umma_commit(&emptyA[st]);
                    a_it++;
                }
                umma_commit(&accFull[acc]);
            }
        }
    } else if (warp == 5) {
        // ======================= weight loader (bulk copies on the TMA engine)
        if (lane == 0) {
            const uint8_t *gw = reinterpret_cast<const uint8_t *>(p.w);
            const uint8_t *gsc = reinterpret_cast<const uint8_t *>(p.sc_w);
            if constexpr (C::RESIDENT) {
                // whole layer stays in shared memory for the lifetime of this persistent CTA
                mbar_arrive_expect_tx(&fullB[0], C::W_MAIN_BYTES + C::W_SC_BYTES);
                for (int s = 0; s < C::NCG * 9; s++)
                    bulk_g2s(sB + s * C::SLAB_BYTES, gw + (size_t)s * C::SLAB_BYTES, C::SLAB_BYTES, &fullB[0]);
                for (int s = 0; s < C::NSC_SLABS; s++)
                    bulk_g2s(sB + C::W_MAIN_BYTES + s * C::SC_SLAB_BYTES, gsc + (size_t)s * C::SC_SLAB_BYTES,
                             C::SC_SLAB_BYTES, &fullB[0]);
            } else {
                uint32_t b_it = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    for (int s = 0; s < C::NCG * 9 + C::NSC_SLABS; s++, b_it++) {
                        const uint32_t bs = b_it % C::NBS;
                        mbar_wait(&emptyB[bs], ((b_it / C::NBS) & 1) ^ 1);
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
        // ======================= A-patch producers (128 threads, cp.async 16 B, zero-fill = padding)
        const int pt = tid - 192;
        constexpr int CH = C::G / 8;        // 16-byte channel chunks per pixel in a main stage
        constexpr int PXSTEP = 128 / CH;    // pixels advanced per pass over the 128 producer threads
        const int j = pt % CH, q0 = pt / CH;
        uint32_t a_it = 0;
        int pending = -1; // stage whose copies were committed but not yet published
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int img0, oy0, ox0;
            if (C::NB == 2) { img0 = tile * 2; oy0 = 0; ox0 = 0; }
            else {
                img0 = tile / C::TILES_PER_IMG;
                const int rem = tile % C::TILES_PER_IMG;
                oy0 = (rem / (C::HOUT / 8)) * 16;
                ox0 = (rem % (C::HOUT / 8)) * 8;
            }
            for (int it = 0; it < C::NCG + (C::CSC > 0 ? 1 : 0); it++, a_it++) {
                const uint32_t st = a_it % C::NAS;
                mbar_wait(&emptyA[st], ((a_it / C::NAS) & 1) ^ 1);
                const uint32_t abase = sA + st * C::A_STAGE_BYTES;
                if (it < C::NCG) {
                    const __half *src = p.in + it * C::G + j * 8;
#pragma unroll 1
                    for (int q = q0; q < C::PATCH_PX; q += PXSTEP) {
                        int y, x, hb;
                        bool ok;
                        if (C::STRIDE == 1) {
                            const int pr = q / C::PITCH, rem = q % C::PITCH;
                            hb = rem / C::BLKW;
                            y = oy0 + pr - 1;
                            x = ox0 + (rem % C::BLKW) - 1;
                            ok = true;
                        } else {
                            const int plane = q / C::PLANE_PX, r2 = q % C::PLANE_PX;
                            const int ip = r2 / C::PITCH, rem = r2 % C::PITCH;
                            const int jp = rem % C::BLKW, py = plane >> 1, px = plane & 1;
                            hb = rem / C::BLKW;
                            y = 2 * (oy0 + ip - py) + py;
                            x = 2 * (ox0 + jp - px) + px;
                            ok = (ip < C::TR + py) && (jp < 8 + px);
                        }
                        const int img = img0 + hb;
                        ok = ok && y >= 0 && y < C::HIN && x >= 0 && x < C::HIN && img < p.nimg;
                        const size_t goff = ok ? (((size_t)img * C::HIN + y) * C::HIN + x) * C::CIN : 0;
                        cp_async16(abase + j * C::A_LBO + q * 16, src + goff, ok);
                    }
                } else {
                    // shortcut operand: block input sampled at (2*oy, 2*ox), rows in accumulator order
                    constexpr int HS = 2 * C::HOUT;
                    constexpr int SCH = C::CSC > 0 ? C::CSC / 8 : 1;
#pragma unroll 1
                    for (int s = pt; s < 128 * SCH; s += 128) {
                        const int m = s / SCH, jc = s % SCH;
                        const int r = m / (8 * C::NB), hb = (m / 8) % C::NB, c = m % 8;
                        const int img = img0 + hb;
                        const bool ok = img < p.nimg;
                        const size_t goff = ok ? (((size_t)img * HS + 2 * (oy0 + r)) * HS + 2 * (ox0 + c)) * C::CSC + jc * 8 : 0;
                        cp_async16(abase + jc * C::SC_LBO + m * 16, p.sc_in + goff, ok);
                    }
                }
                cp_async_commit();
                if (pending >= 0) {
                    cp_async_wait<1>();
                    fence_proxy_async_smem();
                    mbar_arrive(&fullA[pending]);
                }
                pending = (int)st;
            }
        }
        if (pending >= 0) {
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(&fullA[pending]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

} // namespace mlt
