// stem5_cu16.cu -- the composed stem (stem5_umma.cu: staging + conv1 o layer0.0.conv1 as ONE 5x5 stride-2 conv + conv1 at the even
// positions; mlt_cu_or_pq_arch.py:100-106, EncCu.cpp:810-867) for the 16-px CU network.
//
// A 16-px CU has 8 x 8 outputs -- half an M = 128 MMA tile -- and every CU is zero-padded on its own, so CUs cannot share an input
// window.  Here a tile is TWO CUs side by side: accumulator row m = (2 i + h) * 8 + j is output (i, j) of CU h of the pair, the
// 8-row group index 2 i + h walks the expanded operand at a constant pitch because an entry row holds both CUs' entries back to back
// (EP[col parity][row parity][pair][10 entry rows][2 CUs][10 entries][8 fp16]: SBO = 10 entries, LBO = 2 entries as in stem5_umma).
// Work unit = 4 CUs = two pairs = two tiles for the composite conv + two for the conv1 quarter: the same 14 MMAs, 128 TMEM columns
// and barrier protocol per unit as the big kernel.  Every CU has a top row and a left column, so every unit carries the fp32
// border terms (pack_weights.stem5_composite) for 4 x (8 + 8) pixels -- evaluated by a border warp of their own, as in stem5_umma.cu.  With this kernel the 16-px network no longer writes and
// re-reads conv1's 16 x 16 x 32 activation (16 KB per CU each way): it goes from 23 to 22 launches like the 64- / 32-px networks.
#include <cstdlib>
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

namespace stem16 {
constexpr int NEPI = 8, NSTG = 4;
constexpr int W_MMA = NEPI, W_STG = NEPI + 1;
constexpr int W_BRD = NEPI + 1 + NSTG;                     // the border warp
constexpr int NTHREADS = (NEPI + 1 + NSTG + 1) * 32;       // 448
constexpr int CUS = 4;                                     // CUs per work unit (two pairs)
constexpr int ER = 10, EC = 10, ROWP = 2 * EC;             // entry rows per pair; entries per CU per row; entries per row (two CUs)
constexpr int EP_ARR = 2 * ER * ROWP * 16;                 // one (col parity, row parity) array: [pair][ER][ROWP] entries = 6,400 B
constexpr int EP_BYTES = 4 * EP_ARR;                       // 25,600
constexpr int HR = 20, HC = 32, HCU = HR * HC;             // H plane per CU: input rows -2..17, columns -8..23 (half2 {org, res} per sample)
constexpr int H_BYTES = CUS * HCU * 4;                     // 10,240
constexpr int W_BYTES = 7 * 2 * 32 * 8 * 2;                // SEC_STEM5_W
constexpr int CORRW_FLOATS = 2 * 5 * 2 * 32 + 2 * 32;      // Wtop, Wleft [e][ch][co], Wc [ch][co]
constexpr int CORR_BYTES = CUS * 16 * 32 * 4;              // per buffer: per CU 8 top + 8 left pixels x 32 channels, fp32
constexpr int OFF_EP = 0;
constexpr int OFF_H = OFF_EP + 2 * EP_BYTES;
constexpr int OFF_W = (OFF_H + 2 * H_BYTES + 127) / 128 * 128; // two H planes (alternate units; border-warp mode)
constexpr int OFF_CORRW = OFF_W + W_BYTES;
constexpr int OFF_CORR = OFF_CORRW + CORRW_FLOATS * 4;
constexpr int OFF_BIAS = OFF_CORR + 2 * CORR_BYTES;
constexpr int OFF_BAR = OFF_BIAS + 32 * 4;
constexpr int NBAR = 14;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int TMEM_COLS = 256;
static_assert(OFF_W % 128 == 0 && OFF_CORRW % 16 == 0 && OFF_CORR % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
} // namespace stem16

struct Stem16Params {
    const CtuDev *cus;
    const __half *w;     // SEC_STEM5_W  [7][2][32][8]
    const float *corrw;  // SEC_STEM5_CORR (border-term weights, then the kernel's bias [32])
    __half *act0q;       // conv1 at even rows / columns: strip [4 chunks][8 rows][cap][8][8]
    __half *act1;        // layer0.0.conv1 output, same layout
    int n, cap;
    int early;           // MLT_STEM16_EARLY=1 (experiment, results valid): hand EP to the issuer before the border terms
    int bw;              // border-warp mode (MLT_STEM16_BW): the fp32 border terms run on their own warp, off the stagers' chain
};

__global__ void __launch_bounds__(stem16::NTHREADS, 2) stem5_cu16_kernel(const Stem16Params p)
{
    using namespace stem16;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint64_t *ep_full = bars, *ep_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6, *corr_full = bars + 8, *h_full = bars + 10, *h_empty = bars + 12;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_units = (p.n + CUS - 1) / CUS;
    const size_t out_chunk = (size_t)8 * p.cap * 8 * 8; // elements between 8-channel chunks of the strip
    auto out_off = [&](int b, int oy, int ox) -> size_t { return ((size_t)(oy * p.cap + b) * 8 + ox) * 8; };

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&ep_full[i], NSTG); mbar_init(&ep_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], NEPI);
            mbar_init(&corr_full[i], p.bw ? 1 : NSTG);
            mbar_init(&h_full[i], NSTG); mbar_init(&h_empty[i], 1);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < W_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_W)[i] = __ldg(reinterpret_cast<const uint4 *>(p.w) + i);
    for (int i = tid; i < CORRW_FLOATS; i += NTHREADS) reinterpret_cast<float *>(smem + OFF_CORRW)[i] = __ldg(p.corrw + i);
    if (tid < 32) reinterpret_cast<float *>(smem + OFF_BIAS)[tid] = __ldg(p.corrw + CORRW_FLOATS + tid);
    // EP slots no stager writes meet zero weights only; the margins of H ARE the CUs' zero padding and are never written again
    for (int i = tid; i < (2 * EP_BYTES + 2 * H_BYTES) / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_EP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == W_MMA) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sEP = smem_u32(smem + OFF_EP);
    griddep_launch_dependents();
    griddep_wait();

    if (warp == W_BRD) {
        // ======================= border warp: the fp32 border terms of every unit (4 CUs x (8 top + 8 left) pixels x 32 channels)
        if (p.bw) {
            const float *cw = reinterpret_cast<const float *>(smem + OFF_CORRW);
            const int my_units = total_units > (int)blockIdx.x ? (total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            const int px = lane >> 2, cg = (lane & 3) * 8;
            for (int v = 0; v < my_units; v++) {
                const uint32_t buf = v & 1;
                const uint32_t *Hp = reinterpret_cast<const uint32_t *>(smem + OFF_H + buf * H_BYTES);
                // both waits every unit: no producer of this pipeline may run more than ONE phase ahead of a parity waiter
                mbar_wait(&h_full[buf], (v >> 1) & 1);
                mbar_wait(&d_empty[buf], ((v >> 1) & 1) ^ 1); // the epilogue two units ago is done with corr[buf]
                float *corr = reinterpret_cast<float *>(smem + OFF_CORR + buf * CORR_BYTES);
#pragma unroll 1
                for (int side = 0; side < 2; side++) {
                    const bool top = side == 0;
                    const float *wv = cw + (top ? 0 : 5 * 2 * 32);
                    float acc[CUS][8];
#pragma unroll
                    for (int c = 0; c < CUS; c++)
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[c][k] = 0.0f;
#pragma unroll
                    for (int e = 0; e < 5; e++) {
                        const float4 *w0 = reinterpret_cast<const float4 *>(wv + (e * 2 + 0) * 32 + cg), *w1 = reinterpret_cast<const float4 *>(wv + (e * 2 + 1) * 32 + cg);
                        const float4 a0 = w0[0], a1 = w0[1], b0 = w1[0], b1 = w1[1];
                        const float wo[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, wr[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        // top: in(0, 2 j - 2 + e) = H[2][2 j + 6 + e];  left: in(2 i - 2 + e, 0) = H[2 i + e][8]
                        const int hidx = top ? 2 * HC + 2 * px + 6 + e : (2 * px + e) * HC + 8;
#pragma unroll
                        for (int c = 0; c < CUS; c++) {
                            const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(Hp + c * HCU + hidx));
#pragma unroll
                            for (int k = 0; k < 8; k++) acc[c][k] = fmaf(wo[k], xv.x, fmaf(wr[k], xv.y, acc[c][k]));
                        }
                    }
                    if (!top && px == 0) { // conv1(-1, -1)'s share sits in both terms of output (0, 0): take it out of the left one
                        const float *wc = cw + 2 * 5 * 2 * 32;
#pragma unroll
                        for (int c = 0; c < CUS; c++) {
                            const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(Hp + c * HCU + 2 * HC + 8));
#pragma unroll
                            for (int k = 0; k < 8; k++) acc[c][k] -= fmaf(wc[cg + k], xv.x, wc[32 + cg + k] * xv.y);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CUS; c++) {
                        float4 *dst = reinterpret_cast<float4 *>(corr + (c * 16 + side * 8 + px) * 32 + cg);
                        dst[0] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                        dst[1] = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
                    }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(&corr_full[buf]); mbar_arrive(&h_empty[buf]); }
            }
        }
    } else if (warp >= W_STG) {
        // ======================= stagers: one 8-sample vector of org / pred per thread and unit -> H -> EP (+ border terms)
        const int st = tid - W_STG * 32; // 0..127
        uint32_t *H = reinterpret_cast<uint32_t *>(smem + OFF_H);
        constexpr int NE = 4 * 2 * ER * ROWP;        // 1600 entry slots
        constexpr int EPT = (NE + 127) / 128;        // 13
        int src[EPT];
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const int e = st + k * 128;
            src[k] = -1;
            if (e < NE) {
                const int arr = e / (2 * ER * ROWP), rem = e % (2 * ER * ROWP), pr = rem / (ER * ROWP), rem2 = rem % (ER * ROWP);
                const int ri = rem2 / ROWP, col = rem2 % ROWP, h = col / EC, xj = col % EC;
                const int cpar = arr >> 1, rpar = arr & 1, t = 2 * ri + rpar, sc = 2 * xj + cpar;
                // first sample: input row t - 2, column sc - 2 of CU (2 pr + h)  <->  H[cu][t][sc + 6]
                if (cpar == 0 || xj < 8) src[k] = (2 * pr + h) * HCU + t * HC + sc + 6;
            }
        }
        const int vcu = st >> 5, vrow = (st >> 1) & 15, vx = st & 1; // this thread's vector: CU of the unit, sample row, left / right half
        auto load_vec = [&](int u, uint4 &vo, uint4 &vp) {
            const int b = u * CUS + vcu;
            vo = vp = make_uint4(0, 0, 0, 0);
            if (b < p.n) {
                const CtuDev d = p.cus[b];
                vo = __ldg(reinterpret_cast<const uint4 *>(d.org + (size_t)vrow * d.org_stride + vx * 8));
                vp = __ldg(reinterpret_cast<const uint4 *>(d.pred + (size_t)vrow * d.pred_stride + vx * 8));
            }
        };
        uint4 vo, vp;
        if ((int)blockIdx.x < total_units) load_vec(blockIdx.x, vo, vp);
        const float *cw = reinterpret_cast<const float *>(smem + OFF_CORRW);
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            if (p.bw) {
                // two H planes: the plane of unit ul - 2 is free once the border warp is done with it (every stager passed the barrier
                // below in unit ul - 1, i.e. finished its own gather of unit ul - 2)
                H = reinterpret_cast<uint32_t *>(smem + OFF_H + (ul & 1) * H_BYTES);
                mbar_wait(&h_empty[ul & 1], ((ul >> 1) & 1) ^ 1);
            } else {
                asm volatile("bar.sync 1, 128;" ::: "memory"); // every stager is done reading the previous unit's H
            }
            {
                const uint32_t ow[4] = {vo.x, vo.y, vo.z, vo.w}, pw[4] = {vp.x, vp.y, vp.z, vp.w};
                uint32_t hv[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t o = (ow[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu, pp = (pw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
                    const uint32_t co = o < 1023u ? o : 1023u;          // clamp(v / 1023, 0, 1) == min(v, 1023) / 1023 (EncCu.cpp:848-867)
                    const uint32_t ad = o > pp ? o - pp : pp - o;        // cv::absdiff on the (uint16_t) casts (EncCu.cpp:816,827,833)
                    const uint32_t cr = ad < 1023u ? ad : 1023u;
                    const __half2 hh = __floats2half2_rn((float)co * 0.0009765625f, (float)cr * 0.0009765625f); // exact
                    hv[q] = *reinterpret_cast<const uint32_t *>(&hh);
                }
                uint4 *dst = reinterpret_cast<uint4 *>(H + vcu * HCU + (vrow + 2) * HC + 8 + vx * 8);
                dst[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                dst[1] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (p.bw && lane == 0) mbar_arrive(&h_full[ul & 1]); // H is complete: the border warp may start
            if (u + (int)gridDim.x < total_units) load_vec(u + gridDim.x, vo, vp); // prefetch: lands while we gather
            const uint32_t buf = ul & 1;
            mbar_wait(&ep_empty[buf], ((ul >> 1) & 1) ^ 1);
            uint8_t *ep = smem + OFF_EP + buf * EP_BYTES;
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                if (src[k] >= 0) {
                    const uint2 *hp = reinterpret_cast<const uint2 *>(H + (src[k] & ~1));
                    const uint2 w0 = hp[0], w1 = hp[1], w2 = hp[2];
                    const bool odd = src[k] & 1;
                    *reinterpret_cast<uint4 *>(ep + (size_t)(st + k * 128) * 16) =
                        odd ? make_uint4(w0.y, w1.x, w1.y, w2.x) : make_uint4(w0.x, w0.y, w1.x, w1.y);
                }
            }
            // a producer may run at most one phase ahead of a parity waiter (stem5_umma.cu): the epilogue of the unit that used this
            // buffer two units ago must be done before corr[buf] is rewritten and corr_full arrives again
            if (p.early) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ep_full[buf]);
            }
            if (p.bw) continue;
            mbar_wait(&d_empty[buf], ((ul >> 1) & 1) ^ 1);
            {
                // border terms: slot = st >> 1 -> CU (slot >> 4), top row (bit 3 clear) / left column (set), pixel slot & 7; 16 channels per thread
                const int slot = st >> 1, cg = (st & 1) * 16, cu = slot >> 4, px = slot & 7;
                const bool top = (slot & 8) == 0;
                const uint32_t *Hc = H + cu * HCU;
                float acc[16];
#pragma unroll
                for (int k = 0; k < 16; k++) acc[k] = 0.0f;
                const float *wv = cw + (top ? 0 : 5 * 2 * 32);
#pragma unroll
                for (int e = 0; e < 5; e++) {
                    // top: in(0, 2 j - 2 + e) = H[2][2 j + 6 + e];  left: in(2 i - 2 + e, 0) = H[2 i + e][8]
                    const int hidx = top ? 2 * HC + 2 * px + 6 + e : (2 * px + e) * HC + 8;
                    const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(Hc + hidx));
#pragma unroll
                    for (int k = 0; k < 16; k++) acc[k] = fmaf(wv[(e * 2 + 0) * 32 + cg + k], xv.x, fmaf(wv[(e * 2 + 1) * 32 + cg + k], xv.y, acc[k]));
                }
                if (!top && px == 0) { // conv1(-1, -1)'s share sits in both terms of output (0, 0): take it out of the left one
                    const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(Hc + 2 * HC + 8));
                    const float *wc = cw + 2 * 5 * 2 * 32;
#pragma unroll
                    for (int k = 0; k < 16; k++) acc[k] -= fmaf(wc[cg + k], xv.x, wc[32 + cg + k] * xv.y);
                }
                float4 *dst = reinterpret_cast<float4 *>(smem + OFF_CORR + buf * CORR_BYTES + (slot * 32 + cg) * 4);
#pragma unroll
                for (int k = 0; k < 4; k++) dst[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
            }
            if (!p.early) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (!p.early) mbar_arrive(&ep_full[buf]);
                mbar_arrive(&corr_full[buf]);
            }
        }
    } else if (warp == W_MMA) {
        // ======================= MMA issuer: per pair of CUs 5 MMAs (composite 5x5 stride-2 conv) + 2 MMAs (conv1 at even positions)
        constexpr uint32_t idesc = umma_idesc_f16(128, 32);
        constexpr uint32_t a_hi = umma_desc_hi(EC * 16), b_hi = umma_desc_hi(128);
        const uint32_t sW = smem_u32(smem + OFF_W);
        const int my_units = total_units > (int)blockIdx.x ? (total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        for (int v = 0; v < my_units; v++) {
            const uint32_t buf = v & 1;
            mbar_wait(&ep_full[buf], (v >> 1) & 1);
            mbar_wait(&d_empty[buf], ((v >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t ep = sEP + buf * EP_BYTES;
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    const uint32_t d5 = tmem + buf * 128 + pr * 32, dq = d5 + 64;
                    const uint32_t prow = pr * ER * ROWP * 16; // this pair's first entry row inside an array
#pragma unroll
                    for (int dy = 0; dy < 5; dy++) {
                        // array (cpar 0, rpar dy & 1), entry row i + dy / 2, entries j and j + 2 of each CU
                        const uint32_t a_lo = umma_desc_lo(ep + (dy & 1) * EP_ARR + prow + (dy >> 1) * ROWP * 16, 32);
                        umma_f16(d5, umma_desc_pack(a_lo, a_hi), umma_desc_pack(umma_desc_lo(sW + dy * 1024, 512), b_hi), idesc, dy != 0);
                    }
                    // conv1 at (2 i, 2 j): kernel rows 0 / 2 = entry rows i / i + 1 of array (1, 1), kernel row 1 = entry row i + 1 of array (1, 0)
                    const uint32_t q0 = umma_desc_lo(ep + 3 * EP_ARR + prow, ROWP * 16);
                    const uint32_t q1 = umma_desc_lo(ep + 2 * EP_ARR + prow + ROWP * 16, ROWP * 16);
                    umma_f16(dq, umma_desc_pack(q0, a_hi), umma_desc_pack(umma_desc_lo(sW + 5 * 1024, 512), b_hi), idesc, 0);
                    umma_f16(dq, umma_desc_pack(q1, a_hi), umma_desc_pack(umma_desc_lo(sW + 6 * 1024, 512), b_hi), idesc, 1);
                }
                umma_commit(&d_full[buf]);
                umma_commit(&ep_empty[buf]);
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue: warp = (lane quadrant, pair); row m = (2 i + h) * 8 + j
        const int wq = warp & 3, pr = warp >> 2, m = wq * 32 + lane, i = m >> 4, h = (m >> 3) & 1, j = m & 7, cu = 2 * pr + h;
        const float *bias_s = reinterpret_cast<const float *>(smem + OFF_BIAS);
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            const int b = u * CUS + cu;
            const bool valid = b < p.n;
            const uint32_t buf = ul & 1;
            mbar_wait(&corr_full[buf], (ul >> 1) & 1);
            mbar_wait(&d_full[buf], (ul >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + buf * 128 + pr * 32;
            uint32_t v[32];
            tmem_ld32(tbase, v);
            tmem_ld_wait();
            const float *cb = reinterpret_cast<const float *>(smem + OFF_CORR + buf * CORR_BYTES) + cu * 16 * 32;
            const float *ct = cb + j * 32, *cl = cb + (8 + i) * 32;
            const bool top = i == 0, left = j == 0;
            const __half2 zero2 = __float2half2_rn(0.0f);
            {
                __half *op = p.act1 + out_off(b, i, j);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float x[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) x[k] = __uint_as_float(v[q * 8 + k]) + bias_s[q * 8 + k];
                    if (top) {
#pragma unroll
                        for (int k = 0; k < 8; k++) x[k] -= ct[q * 8 + k];
                    }
                    if (left) {
#pragma unroll
                        for (int k = 0; k < 8; k++) x[k] -= cl[q * 8 + k];
                    }
                    uint4 ov;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                    for (int k = 0; k < 4; k++) h2[k] = __hmax2(__floats2half2_rn(x[2 * k], x[2 * k + 1]), zero2);
                    if (valid) *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov;
                }
            }
            tmem_ld32(tbase + 64, v);
            tmem_ld_wait();
            {
                __half *op = p.act0q + out_off(b, i, j);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint4 ov;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                    for (int k = 0; k < 4; k++) h2[k] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * k]), __uint_as_float(v[q * 8 + 2 * k + 1]));
                    if (valid) *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov;
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
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

cudaError_t stem5_cu16_init()
{
    return cudaFuncSetAttribute(stem5_cu16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stem16::SMEM_BYTES);
}

cudaError_t launch_cu16_stem5(const CtuDev *cus, int n, const __half *w, const float *corrw, __half *act0q, __half *act1, int cap, int num_sms,
                              cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    static const int early = getenv("MLT_STEM16_EARLY") ? atoi(getenv("MLT_STEM16_EARLY")) : 0;
    static const int bw = getenv("MLT_STEM16_BW") ? atoi(getenv("MLT_STEM16_BW")) != 0 : 1; // default on; 0 = border terms on the stagers (A/B)
    Stem16Params p{cus, w, corrw, act0q, act1, n, cap, early || bw, bw};
    const int units = (n + stem16::CUS - 1) / stem16::CUS, g = 2 * num_sms;
    return launch_pdl(stem5_cu16_kernel, dim3(units < g ? units : g), dim3(stem16::NTHREADS), stem16::SMEM_BYTES, s, p);
}

} // namespace mlt
