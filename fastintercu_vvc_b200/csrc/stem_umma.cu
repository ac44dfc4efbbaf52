// stem_umma.cu -- the network stem as ONE persistent tcgen05 kernel: input staging (EncCu.cpp:810-867) + conv1
// (arch.py:278, 2->32, 3x3, no BN / ReLU) + layer0.0.conv1 (32->32, 3x3, stride 2, folded BN, ReLU; arch.py:52-57).
//
// Why fused: conv1's output is 1 MiB of fp16 per CTU -- written once and read once it costs more HBM time than all
// the math of both layers.  Here it never leaves the SM: conv1 runs on the tensor cores into TMEM, the epilogue warps
// convert it to fp16 straight into the shared-memory operand patch of the stride-2 conv, and only the 64x64x32 result
// (plus the even/even quarter of conv1's output that layer0.0's 1x1 stride-2 shortcut reads) goes to HBM.
//
// Work unit = 16 x 16 outputs of layer0.0.conv1 (two M = 128 MMA tiles) = a 33 x 33 window of conv1's output, held as
// four (row, column) parity planes of 17 x 17 entries (entry (i, j) of plane (py, px) = conv1 pixel
// (2*oy0 - py + 2*i, 2*ox0 - px + 2*j)), 16 units per CTU.
//   * conv1 on tcgen05: a plane's 289 entries are taken LINEARLY (pitch 17) as the M rows of three M = 128 tiles, so
//     that TMEM lane l of tile t is entry 128*t + l and lands at patch address base + entry*16 -- no index math.
//     The A operand is the expanded input patch EP[col parity][row parity][18][17][8 fp16]: entry = {org, res} of the
//     four input pixels x-1..x+2 (K chunk = one kernel row kh; the 4th pixel has zero weights); samples enter as
//     v * 2^-10 (exact in fp16) and the fp16 weights carry (float)(1/1023) * 2^10 (pack_weights.py).
//     Row / column parity splitting makes the stride-2 sampling of every plane a dense window again, and the three
//     kernel rows are start-address / LBO variants of the same arrays (two K=16 MMAs per hi / lo part).
//   * layer0.0.conv1 then reads the patch exactly like conv_umma.cuh reads a TMA-loaded parity patch.
// Pipeline per CTA (1 per SM, 17 warps): 4 stager warps (int16 -> EP), 1 MMA issuer, 8 mid-epilogue warps
// (TMEM -> fp16 patch, zeroing entries outside the picture = the stride-2 conv's padding), 4 final-epilogue warps.
// TMEM: 12 conv1 accumulators (384 columns) + 2 x 2 output accumulators (128 columns).  The issuer interleaves
// conv1(u+1) before conv(u), so the patch conversion of unit u overlaps the MMAs of unit u+1.
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

namespace stem {
constexpr int NE1 = 8;                                      // mid-epilogue warps (2 per TMEM lane quadrant; 16 measured no faster: the stem is smem-port-bound)
constexpr int NTHREADS = (4 + NE1 + 2 + 4) * 32;           // 576
constexpr int W_E2 = 0, W_E1 = 4, W_MMA = 4 + NE1, W_MMA2 = W_MMA + 1, W_STG = W_MMA + 2; // warp roles (first warp of each group)
constexpr int PE = 17, PLANE_ENT = PE * PE;                // 289 entries per parity plane
constexpr int P_LBO = PLANE_ENT * 16;                      // bytes between 8-channel chunks of the patch
constexpr int P_PLANE = 4 * P_LBO;                         // one parity plane (32 channels)
constexpr int PATCH_BYTES = 4 * P_PLANE;                   // 73,984
constexpr int EP_ROWS = 18, EP_ARR = EP_ROWS * PE * 16;    // one (col parity, row parity) array: 306 entries
constexpr int EP_BYTES = 4 * EP_ARR;                       // 19,584
constexpr int RAW_COLS = 48, RAW_ROWS = 35;
constexpr int RAW_BYTES = RAW_ROWS * RAW_COLS * 4;         // 6,720: fp16 {org, res} pair per pixel of the input window
constexpr int W0_BYTES = 9 * 32 * 32 * 2;                  // layer0.0.conv1 weights, resident
constexpr int W1_BYTES = 4 * 1024;                         // conv1 operand variants [py][2 MMAs][2 chunks][32][8]
constexpr int OFF_PATCH = 0;
constexpr int OFF_EP = OFF_PATCH + 2 * PATCH_BYTES;
constexpr int OFF_RAW = OFF_EP + 2 * EP_BYTES;
constexpr int OFF_W0 = (OFF_RAW + RAW_BYTES + 127) / 128 * 128;
constexpr int OFF_W1 = OFF_W0 + W0_BYTES;
constexpr int OFF_BAR = OFF_W1 + W1_BYTES;
constexpr int NBAR = 2 + 2 + 12 + 4 + 2 + 2 + 2 + 2;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int TM_C1 = 0, TM_D = 384; // TMEM columns
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
} // namespace stem

__device__ __forceinline__ uint32_t absdiff16(uint32_t o, uint32_t p) { return o > p ? o - p : p - o; } // cv::absdiff, CV_16U

// S = block size in luma samples: 128 (the CTU network: outputs in the dense per-CTU layout [ctu][4][64][64][8]) or 64 / 32
// (the smaller-CU networks: outputs in the STRIP layout [4 chunks][S/2 rows][cap images][S/2][8], conv_umma.cuh) --
// (S / 32)^2 work units per block.  A 16-px CU is smaller than one work unit and keeps the unfused conv1 kernel.
struct StemParams {
    const CtuDev *ctus;
    const __half *w1;   // conv1 operand variants (SEC_STEM_CONV1)
    const __half *w0;   // layer0.0.conv1 packed [9][4][32][8] (SEC_W_F16 + 0)
    const float *bias;  // layer0.0.conv1 folded-BN bias, fp32 [32] (SEC_BIAS_FUSED + 0), added by the final epilogue
    __half *act0q;      // conv1 output at even rows / even columns: dense chunk-planar [ctu][4][64][64][8]
    __half *act1;       // layer0.0.conv1 output: dense chunk-planar [ctu][4][64][64][8]
    int n;
    int cap;            // strip layouts (S < 128): images per strip
};

template <int S>
__global__ void __launch_bounds__(stem::NTHREADS, 1) stem_umma_kernel(const StemParams p)
{
    using namespace stem;
    constexpr int OH = S / 2, UW = S / 32, UPI = UW * UW; // output map size; work units per row / per block
    // element offset of output pixel (oy, ox), channel chunk 0, of block `b` and the stride between channel chunks
    auto out_off = [&](int b, int oy, int ox) -> size_t {
        return S == 128 ? (size_t)b * (4 * OH * OH * 8) + (size_t)(oy * OH + ox) * 8 : ((size_t)(oy * p.cap + b) * OH + ox) * 8;
    };
    const size_t out_chunk = S == 128 ? (size_t)OH * OH * 8 : (size_t)OH * p.cap * OH * 8;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint64_t *ep_full = bars, *ep_empty = bars + 2, *c1_full = bars + 4, *c1_empty = bars + 16;
    uint64_t *patch_full = bars + 20, *patch_empty = bars + 22, *d_full = bars + 24, *d_empty = bars + 26;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_units = p.n * UPI;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&ep_full[i], 4); mbar_init(&ep_empty[i], 1);
            mbar_init(&patch_full[i], NE1); mbar_init(&patch_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 4);
        }
        for (int i = 0; i < 12; i++) mbar_init(&c1_full[i], 1);
        for (int i = 0; i < 4; i++) mbar_init(&c1_empty[i], 12); // per parity plane: 3 tiles x 4 lane-quadrant warps
        mbar_fence_init();
    }
    for (int i = tid; i < W0_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_W0)[i] = __ldg(reinterpret_cast<const uint4 *>(p.w0) + i);
    for (int i = tid; i < W1_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_W1)[i] = __ldg(reinterpret_cast<const uint4 *>(p.w1) + i);
    // EP rows that no stager ever writes must still hold finite numbers
    for (int i = tid; i < 2 * EP_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_EP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == W_MMA) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sPatch = smem_u32(smem + OFF_PATCH), sEP = smem_u32(smem + OFF_EP);
    griddep_launch_dependents(); // PDL: the prologue above overlapped the previous kernel's tail
    griddep_wait();

    if (warp >= W_STG) {
        // ======================= stagers: int16 window -> fp16 {org, res} pair plane H -> expanded, parity-split operand EP
        // Phase 1 converts every pixel ONCE (integer staging + exact *2^-10) into H[35 rows][48 px] (half2 per pixel);
        // phase 2 only gathers: an EP entry is the 16 bytes H[t][s+6 .. s+9].  Both index tables are unit-invariant and
        // live in registers; the next unit's global loads are issued before this unit's gather (latency off the path).
        const int st = tid - W_STG * 32; // 0..127
        uint32_t *H = reinterpret_cast<uint32_t *>(smem + OFF_RAW);
        constexpr int NV = RAW_ROWS * 6;          // 210 (row, 8-pixel vector) pairs of the window
        constexpr int NE = 4 * EP_ROWS * PE;      // 1224 entries
        constexpr int EPT = (NE + 127) / 128;     // entries per thread
        int src[EPT];                             // H index of the entry's first pixel, or -1 (row 35: stays zero)
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const int e = st + k * 128;
            src[k] = -1;
            if (e < NE) {
                const int arr = e / (EP_ROWS * PE), rem = e % (EP_ROWS * PE), ri = rem / PE, xj = rem % PE;
                const int cpar = arr >> 1, rpar = arr & 1, t = 2 * ri + rpar, sc = 2 * xj + cpar;
                if (t < RAW_ROWS) src[k] = t * RAW_COLS + sc + 6; // centre column x = 2*ox0 - 1 + sc, pixels x-1 .. x+2
            }
        }
        auto load_window = [&](int u, uint4 (&vo)[2], uint4 (&vp)[2]) {
            // rows Y = 2*oy0 - 2 + t (t < 35), columns X = 2*ox0 - 8 + c (c < 48); outside the CTU = conv1's zero padding
            const int ctu = u / UPI, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            const CtuDev d = p.ctus[ctu];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int i = st + k * 128, t = i / 6, vx = i % 6;
                const int Y = 2 * oy0 - 2 + t, X = 2 * ox0 - 8 + vx * 8;
                vo[k] = vp[k] = make_uint4(0, 0, 0, 0);
                if (i < NV && Y >= 0 && Y < S && X >= 0 && X < S) {
                    vo[k] = __ldg(reinterpret_cast<const uint4 *>(d.org + (size_t)Y * d.org_stride + X));
                    vp[k] = __ldg(reinterpret_cast<const uint4 *>(d.pred + (size_t)Y * d.pred_stride + X));
                }
            }
        };
        uint4 vo[2], vp[2];
        if ((int)blockIdx.x < total_units) load_window(blockIdx.x, vo, vp);
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            asm volatile("bar.sync 1, 128;" ::: "memory"); // every stager is done gathering from the previous unit's H
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int i = st + k * 128;
                if (i < NV) {
                    const uint32_t ow[4] = {vo[k].x, vo[k].y, vo[k].z, vo[k].w}, pw[4] = {vp[k].x, vp[k].y, vp[k].z, vp[k].w};
                    uint32_t hv[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t o = (ow[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu, pp = (pw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
                        const uint32_t co = o < 1023u ? o : 1023u; // clamp(v / 1023, 0, 1) == min(v, 1023) / 1023 (EncCu.cpp:848-867)
                        const uint32_t ad = absdiff16(o, pp);      // cv::absdiff on the (uint16_t) casts (EncCu.cpp:816,827,833)
                        const uint32_t cr = ad < 1023u ? ad : 1023u;
                        const __half2 h = __floats2half2_rn((float)co * 0.0009765625f, (float)cr * 0.0009765625f); // exact
                        hv[q] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(H + (size_t)i * 8);
                    dst[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                    dst[1] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (u + (int)gridDim.x < total_units) load_window(u + gridDim.x, vo, vp); // prefetch: lands while we gather
            const uint32_t buf = ul & 1;
            mbar_wait(&ep_empty[buf], ((ul >> 1) & 1) ^ 1);
            uint8_t *ep = smem + OFF_EP + buf * EP_BYTES;
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                if (src[k] >= 0) {
                    // three aligned 8-byte loads (conflict-free: consecutive lanes are 8 bytes apart) + a parity select,
                    // instead of four 4-byte loads at an 8-byte lane stride (2..4-way bank conflicts)
                    const uint2 *hp = reinterpret_cast<const uint2 *>(H + (src[k] & ~1));
                    const uint2 w0 = hp[0], w1 = hp[1], w2 = hp[2];
                    const bool odd = src[k] & 1;
                    *reinterpret_cast<uint4 *>(ep + (size_t)(st + k * 128) * 16) =
                        odd ? make_uint4(w0.y, w1.x, w1.y, w2.x) : make_uint4(w0.x, w0.y, w1.x, w1.y);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ep_full[buf]);
        }
    } else if (warp == W_MMA || warp == W_MMA2) {
        // ======================= MMA issuers: conv1(u) [24 MMAs] / layer0.0.conv1(u) [2 x 18 MMAs]
        constexpr uint32_t idesc = umma_idesc_f16(128, 32);
        constexpr uint32_t e_hi = umma_desc_hi(128), b_hi = umma_desc_hi(128), p_hi = umma_desc_hi(PE * 16);
        const uint32_t sW0 = smem_u32(smem + OFF_W0), sW1 = smem_u32(smem + OFF_W1);
        const int my_units = total_units > (int)blockIdx.x ? (total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        // two issuer warps (one warp cannot issue both MMA streams fast enough): W_MMA runs conv1, W_MMA2 the stride-2 conv;
        // they only meet through the mid-epilogue's barriers, so conv1(u+1) naturally overlaps conv(u)
        if (warp == W_MMA) {
            for (int ul = 0; ul < my_units; ul++) {
                const uint32_t buf = ul & 1;
                mbar_wait(&ep_full[buf], (ul >> 1) & 1);
                tc_fence_after();
                const uint32_t ep = sEP + buf * EP_BYTES;
                // Per tile TWO K=16 MMAs cover the three kernel rows (fp16 weights, error-diffused over the taps by the packer;
                // the 4th chunk slot carries zero weights); the two chunks of an MMA are kernel rows whose EP addresses differ
                // by a positive constant (LBO):
                //   py = 0: rows kh0 @ R1+0, kh1 @ R0+17, kh2 @ R1+17   (R0 / R1 = row-parity arrays, entries)
                //           [kh1 | kh0] x [w1 | w0],  [kh0 | kh2] x [0 | w2]
                //   py = 1: rows kh0 @ R0+0, kh1 @ R1+0, kh2 @ R0+17
                //           [kh0 | kh1] x [w0 | w1],  [kh0 | kh2] x [0 | w2]
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    const int plane = k / 3, t = k % 3, py = plane >> 1, px = plane & 1;
                    if (t == 0) { mbar_wait(&c1_empty[plane], (ul & 1) ^ 1); tc_fence_after(); }
                    if (elect_one_sync()) {
                        const uint32_t a0 = ep + (1 - px) * 2 * EP_ARR + t * 128 * 16; // row-parity-0 array of column parity 1 - px
                        const uint32_t d_tmem = tmem + TM_C1 + k * 32;
                        const uint32_t wb = sW1 + py * 2048;
                        constexpr uint32_t ROW = PE * 16, L289 = (EP_ARR - PE * 16);
                        const uint32_t s0 = py ? a0 : a0 + ROW, s1 = py ? a0 : a0 + EP_ARR;
                        const uint32_t l0 = py ? EP_ARR : L289;
                        umma_f16(d_tmem, umma_desc_pack(umma_desc_lo(s0, l0), e_hi), umma_desc_pack(umma_desc_lo(wb, 32 * 16), b_hi), idesc, 0);
                        umma_f16(d_tmem, umma_desc_pack(umma_desc_lo(s1, ROW), e_hi), umma_desc_pack(umma_desc_lo(wb + 1024, 32 * 16), b_hi), idesc, 1);
                        umma_commit(&c1_full[k]);
                    }
                    __syncwarp();
                }
                if (elect_one_sync()) umma_commit(&ep_empty[buf]);
                __syncwarp();
            }
        } else {
            for (int v = 0; v < my_units; v++) {
                const uint32_t buf = v & 1;
                mbar_wait(&patch_full[buf], (v >> 1) & 1);
                mbar_wait(&d_empty[buf], ((v >> 1) & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t pb = sPatch + buf * PATCH_BYTES;
#pragma unroll
                    for (int half = 0; half < 2; half++) { // left / right 8 output columns
                        const uint32_t d_tmem = tmem + TM_D + buf * 64 + half * 32;
#pragma unroll
                        for (int tap = 0; tap < 9; tap++) {
                            const int kh = tap / 3, kw = tap % 3;
                            const int py = (kh != 1), ro = (kh == 2), px = (kw != 1), co = (kw == 2);
                            const uint32_t a_lo = umma_desc_lo(pb + (py * 2 + px) * P_PLANE + (ro * PE + co + half * 8) * 16, P_LBO);
                            const uint32_t b_lo = umma_desc_lo(sW0 + tap * 2048, 32 * 16);
#pragma unroll
                            for (int ks = 0; ks < 2; ks++)
                                umma_f16(d_tmem, umma_desc_pack(a_lo + ks * (2 * P_LBO / 16), p_hi), umma_desc_pack(b_lo + ks * 64, b_hi), idesc,
                                         (tap | ks) != 0); // bias is added by the final epilogue
                        }
                    }
                    umma_commit(&d_full[buf]);
                    umma_commit(&patch_empty[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= W_E1) {
        // ======================= mid-epilogue: conv1 accumulators -> fp16 -> stride-2 conv's operand patch (+ even/even quarter to HBM)
        const int wq = warp & 3, hsel = (warp - W_E1) >> 2; // TMEM lane quadrant; which of the NE1/4 tile subsets
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            const int ctu = u / UPI, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            const uint32_t buf = ul & 1;
            mbar_wait(&patch_empty[buf], ((ul >> 1) & 1) ^ 1);
            uint8_t *patch = smem + OFF_PATCH + buf * PATCH_BYTES;
#pragma unroll 1
            for (int k = hsel; k < 12; k += NE1 / 4) {
                const int plane = k / 3, t = k % 3, py = plane >> 1, px = plane & 1;
                const int e = t * 128 + wq * 32 + lane, i = e / PE, j = e % PE;
                // the third tile of a plane holds entries 256..288 only: lane quadrants 2 and 3 (and 1, except for the corner
                // entry 288 of plane (1, 1)) have nothing to convert -- skipping their TMEM reads matters, the stem is bound by
                // the TMEM -> register read bandwidth (196 KB of fp32 accumulators per unit)
                const bool live = t < 2 || wq == 0 || (wq == 1 && plane == 3);
                mbar_wait(&c1_full[k], ul & 1);
                tc_fence_after();
                uint32_t v[32];
                if (live) {
                    tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + TM_C1 + k * 32, v);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&c1_empty[plane]);
                if (live && e < PLANE_ENT) {
                    // conv1 pixel (2*oy0 - py + 2*i, 2*ox0 - px + 2*j); outside the picture = zero padding of the stride-2 conv
                    const bool zero = (py && i == 0 && oy0 == 0) || (px && j == 0 && ox0 == 0);
                    uint4 ov[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        __half2 *h2 = reinterpret_cast<__half2 *>(&ov[q]);
#pragma unroll
                        for (int x = 0; x < 4; x++)
                            h2[x] = zero ? __float2half2_rn(0.0f) : __floats2half2_rn(__uint_as_float(v[q * 8 + x * 2]), __uint_as_float(v[q * 8 + x * 2 + 1]));
                        *reinterpret_cast<uint4 *>(patch + plane * P_PLANE + q * P_LBO + e * 16) = ov[q];
                    }
                    if (plane == 0 && i < 16 && j < 16) { // conv1 at (2*(oy0+i), 2*(ox0+j)): input of layer0.0's 1x1 stride-2 shortcut
                        __half *op = p.act0q + out_off(ctu, oy0 + i, ox0 + j);
#pragma unroll
                        for (int q = 0; q < 4; q++) *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov[q];
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&patch_full[buf]);
        }
    } else {
        // ======================= final epilogue: layer0.0.conv1 accumulators (bias included) -> ReLU -> fp16 -> HBM
        const int wq = warp & 3, m = wq * 32 + lane, r = m >> 3, c = m & 7;
        float bias_r[32];
#pragma unroll
        for (int j = 0; j < 32; j++) bias_r[j] = __ldg(p.bias + j);
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            const int ctu = u / UPI, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            const uint32_t buf = ul & 1;
            mbar_wait(&d_full[buf], (ul >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + TM_D + buf * 64 + half * 32, v);
                tmem_ld_wait();
                const __half2 zero2 = __float2half2_rn(0.0f);
                __half *op = p.act1 + out_off(ctu, oy0 + r, ox0 + half * 8 + c);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint4 ov;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                    for (int x = 0; x < 4; x++)
                        h2[x] = __hmax2(__floats2half2_rn(__uint_as_float(v[q * 8 + x * 2]) + bias_r[q * 8 + x * 2],
                                                          __uint_as_float(v[q * 8 + x * 2 + 1]) + bias_r[q * 8 + x * 2 + 1]), zero2);
                    *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

cudaError_t stem_umma_init()
{
    cudaError_t e = cudaFuncSetAttribute(stem_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem::SMEM_BYTES);
    return e;
}

cudaError_t launch_stem_umma(const CtuDev *ctus, int n, const __half *w1, const __half *w0, const float *bias, __half *act0q,
                             __half *act1, int num_sms, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    StemParams p{ctus, w1, w0, bias, act0q, act1, n, 0};
    const int units = n * 16;
    return launch_pdl(stem_umma_kernel<128>, dim3(units < num_sms ? units : num_sms), dim3(stem::NTHREADS), stem::SMEM_BYTES, s, p);
}

// the same stem for a 64- or 32-px CU network: act0q / act1 are strips of `cap` images (S/2 x S/2 x 32, not parity-planar)
cudaError_t launch_cu_stem_umma(int size, const CtuDev *cus, int n, const __half *w1, const __half *w0, const float *bias, __half *act0q,
                                __half *act1, int cap, int num_sms, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    StemParams p{cus, w1, w0, bias, act0q, act1, n, cap};
    const int units = n * (size / 32) * (size / 32);
    const dim3 grid(units < num_sms ? units : num_sms), block(stem::NTHREADS);
    if (size == 64) return launch_pdl(stem_umma_kernel<64>, grid, block, stem::SMEM_BYTES, s, p);
    if (size == 32) return launch_pdl(stem_umma_kernel<32>, grid, block, stem::SMEM_BYTES, s, p);
    return cudaErrorInvalidValue;
}

} // namespace mlt
