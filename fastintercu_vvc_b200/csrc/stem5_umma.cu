// stem5_umma.cu -- the network stem as ONE persistent tcgen05 kernel, conv1 and layer0.0.conv1 COMPOSED:
// input staging (EncCu.cpp:810-867) + [conv1 (arch.py:278) o layer0.0.conv1 (arch.py:52-57)] + conv1 at the even positions.
//
// The reference applies NO BatchNorm and NO activation between conv1 (2 -> 32, 3x3, pad 1) and layer0.0 (arch.py:277-279:
// `out = self.conv1(x); out = self.layer0(out)`, bn1 defined but unused).  Two convolutions in a row are one linear map, so
//     relu(bn(conv3x3_s2(conv3x3(x))))  ==  relu(W5 (*) x + b)      with W5 a 5x5 stride-2 kernel over the TWO input planes
// -- 100 input taps instead of 18 + 288 per output, and conv1's 32 x 128 x 128 output is never formed at all (the round-1 stem
// kept it on chip but still spent 60 MMAs and 196 KB of TMEM reads per 16x16 outputs on it; this kernel spends 14 and 64 KB).
// The only place the composition is not exact is layer0.0.conv1's own zero padding: at output row 0 / column 0 it replaces
// conv1's outputs at row / column -1 by zeros, whereas W5 sees conv1's (non-zero) values there.  Those terms are 5-tap
// 1-D filters over input row 0 / column 0 (pack_weights.stem5_composite: Wtop, Wleft, Wc); a border warp evaluates them on
// the CUDA cores for the 16 + 16 border pixels of a border unit and the epilogue subtracts them (fp32).
// layer0.0's shortcut reads conv1 at the even rows / columns (1x1 stride 2): that quarter is still produced here, as a 3x3
// stride-2 conv of the input (K = 18 -> two K=16 MMAs), and stored for layer0.0.conv2's extra operand as before.
//
// Work unit = 16 x 16 outputs = two M = 128 MMA tiles (left / right 8 columns), 16 units per CTU.  A operand = the expanded
// input patch EP[column parity][row parity][18 rows][18 entries][8 fp16]: entry (ri, xj) of array (cpar, rpar) = {org, res} of
// the four input samples (row 2*oy0 - 2 + 2*ri + rpar, columns 2*ox0 - 2 + 2*xj + cpar .. + 3), samples as v * 2^-10 (exact in
// fp16; the weights carry (float)(1/1023) * 2^10).  Output pixel (i, j) then finds kernel row dy of W5 at entry
// (i + dy/2, j) and (i + dy/2, j + 2) of array (0, dy & 1): SBO = one entry row, LBO = two entries -- five MMAs per tile, no
// index math; conv1-at-even-positions reads arrays (1, *) the same way.
// Pipeline per CTA (2 per SM, 14 warps): 4 stager warps (int16 -> H -> EP), 1 MMA issuer, 8 epilogue warps (lane quadrant x
// tile half), 1 border warp (the fp32 border terms from the unit's H plane -- two planes, alternate units -- into corr[buf]; on the
// stagers they were ~15 % of the chain that bounds the kernel: stem 0.757 -> 0.672 ms per 3840 CTUs; MLT_STEM5_BW=0 puts them back).  TMEM: 2 buffers x (2 x 32 columns W5 result + 2 x 32 columns conv1 quarter) = 256 columns.
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

namespace stem5 {
constexpr int NEPI = 8;
constexpr int W_EPI = 0, W_MMA = NEPI, W_STG = NEPI + 1;
__host__ __device__ constexpr int nthreads(int nstg) { return (NEPI + 1 + nstg + 1) * 32; } // 448 with 4 stager warps: + the border warp
constexpr int PE = 18, EP_ROWS = 18;
constexpr int EP_ARR = EP_ROWS * PE * 16;                  // one (col parity, row parity) array: 324 entries
constexpr int EP_BYTES = 4 * EP_ARR;                       // 20,736
constexpr int RAW_COLS = 48, RAW_ROWS = 35;
constexpr int RAW_BYTES = RAW_ROWS * RAW_COLS * 4;         // fp16 {org, res} pair per sample of the input window
constexpr int W_BYTES = 7 * 2 * 32 * 8 * 2;                // SEC_STEM5_W
constexpr int CORRW_FLOATS = 2 * 5 * 2 * 32 + 2 * 32;      // SEC_STEM5_CORR: Wtop, Wleft [e][ch][co], Wc [ch][co]
constexpr int CORR_BYTES = 32 * 32 * 4;                    // per buffer: 16 top + 16 left border pixels x 32 channels, fp32
constexpr int OFF_EP = 0;
constexpr int OFF_RAW = OFF_EP + 2 * EP_BYTES;
constexpr int NPLANES = 4;                                // H planes: two (alternate units) per stager group
constexpr int OFF_W = (OFF_RAW + NPLANES * RAW_BYTES + 127) / 128 * 128;
constexpr int OFF_CORRW = OFF_W + W_BYTES;
constexpr int OFF_CORR = OFF_CORRW + CORRW_FLOATS * 4;
constexpr int OFF_BIAS = OFF_CORR + 2 * CORR_BYTES;
constexpr int OFF_BAR = OFF_BIAS + 32 * 4;
constexpr int NBAR = 10 + 2 * NPLANES;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int TMEM_COLS = 256;
static_assert(OFF_W % 128 == 0 && OFF_CORRW % 16 == 0 && OFF_CORR % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
} // namespace stem5

constexpr int STEM5_DEFAULT_STAGERS = 4;

struct Stem5Params {
    const CtuDev *ctus;
    const __half *w;     // SEC_STEM5_W  [7][2][32][8]
    const float *corrw;  // SEC_STEM5_CORR
    const float *bias;   // layer0.0.conv1 folded-BN bias, fp32 [32]
    __half *act0q;       // conv1 output at even rows / even columns
    __half *act1;        // layer0.0.conv1 output
    int n;
    int cap;             // strip layouts (S < 128): images per strip
    int bw;              // border-warp mode (MLT_STEM5_BW): the fp32 border terms run on their own warp, off the stagers' chain
    int dbg;             // MLT_STEM5_DBG (timing experiments only, results invalid): 1 no global stores, 2 no global loads, 4 no EP gather
};

__device__ __forceinline__ uint32_t s5_absdiff16(uint32_t o, uint32_t p) { return o > p ? o - p : p - o; } // cv::absdiff, CV_16U

template <int S, int NSTG>
__global__ void __launch_bounds__(stem5::nthreads(NSTG), 2) stem5_umma_kernel(const Stem5Params p)
{
    using namespace stem5;
    // NSTG = 8: TWO stager groups of four warps, each staging alternate units into its own H plane / EP buffer.  A unit's staging is a
    // latency chain (global loads -> bar.sync -> convert -> bar.sync -> gather -> arrive) that more threads do not shorten; two
    // chains in flight do.
    constexpr int NTHREADS = nthreads(NSTG), GROUPS = NSTG == 8 ? 2 : 1, STG_THREADS = NSTG * 32 / GROUPS;
    constexpr int OH = S / 2, UW = S / 32, UPI = UW * UW; // output map size; work units per row / per block
    auto out_off = [&](int b, int oy, int ox) -> size_t {
        return S == 128 ? (size_t)b * (4 * OH * OH * 8) + (size_t)(oy * OH + ox) * 8 : ((size_t)(oy * p.cap + b) * OH + ox) * 8;
    };
    const size_t out_chunk = S == 128 ? (size_t)OH * OH * 8 : (size_t)OH * p.cap * OH * 8;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
    uint64_t *ep_full = bars, *ep_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6, *corr_full = bars + 8;
    uint64_t *h_full = bars + 10, *h_empty = bars + 10 + NPLANES;
    constexpr int W_BRD = NEPI + 1 + NSTG; // the border warp
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_units = p.n * UPI;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&ep_full[i], NSTG / GROUPS); mbar_init(&ep_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], NEPI);
            mbar_init(&corr_full[i], p.bw ? 1 : NSTG / GROUPS);
        }
        for (int i = 0; i < NPLANES; i++) { mbar_init(&h_full[i], NSTG / GROUPS); mbar_init(&h_empty[i], 1); }
        mbar_fence_init();
    }
    for (int i = tid; i < W_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_W)[i] = __ldg(reinterpret_cast<const uint4 *>(p.w) + i);
    for (int i = tid; i < CORRW_FLOATS; i += NTHREADS) reinterpret_cast<float *>(smem + OFF_CORRW)[i] = __ldg(p.corrw + i);
    if (tid < 32) reinterpret_cast<float *>(smem + OFF_BIAS)[tid] = __ldg(p.bias + tid);
    // EP entries that no stager ever writes are still multiplied (by zero weights, or land in unused accumulator rows): keep them finite
    for (int i = tid; i < 2 * EP_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem + OFF_EP)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == W_MMA) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sEP = smem_u32(smem + OFF_EP);
    const bool spin = (p.dbg & 128) != 0; // timing experiment: mbarrier.test_wait polling instead of try_wait
    const bool lane0_poll = (p.dbg & 256) != 0; // timing experiment: one lane per warp polls, the others park at __syncwarp
    auto bwait = [&](uint64_t *bar, uint32_t parity) {
        if (lane0_poll) {
            if (lane == 0) mbar_wait(bar, parity);
            __syncwarp();
        } else if (spin) mbar_wait_spin(bar, parity);
        else mbar_wait(bar, parity);
    };
    griddep_launch_dependents(); // PDL: the prologue above overlapped the previous kernel's tail
    griddep_wait();

    if (warp == W_BRD) {
        // ======================= border warp: the fp32 border terms of every unit of this CTA, from the unit's H plane into corr[buf]
        if (p.bw) {
            const float *cw = reinterpret_cast<const float *>(smem + OFF_CORRW);
            const int my_units = total_units > (int)blockIdx.x ? (total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            const int sb = lane >> 2, cg = (lane & 3) * 8; // pixel slots sb, sb + 8 of the top row / left column; 8 channels
            for (int v = 0; v < my_units; v++) {
                const int u = blockIdx.x + v * gridDim.x, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
                const int it = v / GROUPS, plane = (v % GROUPS) * 2 + (it & 1);
                const uint32_t buf = v & 1;
                const uint32_t *H = reinterpret_cast<const uint32_t *>(smem + OFF_RAW + plane * RAW_BYTES);
                // both waits every unit: no producer of this pipeline may run more than ONE phase ahead of a parity waiter
                mbar_wait(&h_full[plane], (it >> 1) & 1);
                mbar_wait(&d_empty[buf], ((v >> 1) & 1) ^ 1); // the epilogue two units ago is done with corr[buf]
                float *corr = reinterpret_cast<float *>(smem + OFF_CORR + buf * CORR_BYTES);
#pragma unroll 1
                for (int side = 0; side < 2; side++) {
                    const bool top = side == 0;
                    if (top ? oy0 != 0 : ox0 != 0) continue;
                    const float *wv = cw + (top ? 0 : 5 * 2 * 32);
                    // pixel slots sb and sb + 8: output (0, ox0 + px) of the top row / (oy0 + px, 0) of the left column
                    float acc[2][8];
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[0][k] = acc[1][k] = 0.0f;
#pragma unroll
                    for (int e = 0; e < 5; e++) {
                        // top: in(0, 2*(ox0+j) - 2 + e) = H[2][2j + 6 + e];  left: in(2*(oy0+i) - 2 + e, 0) = H[2i + e][8]
                        const float4 *w0 = reinterpret_cast<const float4 *>(wv + (e * 2 + 0) * 32 + cg), *w1 = reinterpret_cast<const float4 *>(wv + (e * 2 + 1) * 32 + cg);
                        const float4 a0 = w0[0], a1 = w0[1], b0 = w1[0], b1 = w1[1];
                        const float wo[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, wr[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const int px = sb + 8 * r;
                            const int hidx = top ? 2 * RAW_COLS + 2 * px + 6 + e : (2 * px + e) * RAW_COLS + 8;
                            const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(H + hidx));
#pragma unroll
                            for (int k = 0; k < 8; k++) acc[r][k] = fmaf(wo[k], xv.x, fmaf(wr[k], xv.y, acc[r][k]));
                        }
                    }
                    if (!top && sb == 0 && oy0 == 0) { // conv1(-1, -1)'s share sits in both terms of output (0, 0): take it out of this one
                        const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(H + 2 * RAW_COLS + 8));
                        const float *wc = cw + 2 * 5 * 2 * 32;
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[0][k] -= fmaf(wc[cg + k], xv.x, wc[32 + cg + k] * xv.y);
                    }
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        float4 *dst = reinterpret_cast<float4 *>(corr + ((top ? 0 : 16) + sb + 8 * r) * 32 + cg);
                        dst[0] = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                        dst[1] = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
                    }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(&corr_full[buf]); mbar_arrive(&h_empty[plane]); }
            }
        }
    } else if (warp >= W_STG) {
        // ======================= stagers: int16 window -> fp16 {org, res} plane H -> expanded, parity-split operand EP (+ border terms)
        const int grp = (tid - W_STG * 32) / STG_THREADS, st = (tid - W_STG * 32) % STG_THREADS; // stager group; thread within the group
        uint32_t *H = reinterpret_cast<uint32_t *>(smem + OFF_RAW + grp * 2 * RAW_BYTES);
        const int u_first = blockIdx.x + grp * gridDim.x, u_step = GROUPS * gridDim.x; // group g stages this CTA's units g, g + GROUPS, ...
        auto group_sync = [&]() { // named barrier of this group (immediate ids: a register id makes ptxas reserve all 16 barriers)
            if (p.dbg & 1024) return; // timing experiment
            if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(STG_THREADS) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"(STG_THREADS) : "memory");
        };
        constexpr int NV = RAW_ROWS * 6;          // 210 (row, 8-sample vector) pairs of the window
        constexpr int NE = 4 * EP_ROWS * PE;      // 1296 entry slots
        constexpr int EPT = (NE + STG_THREADS - 1) / STG_THREADS; // entry slots per thread
        constexpr int VPT = (NV + STG_THREADS - 1) / STG_THREADS; // window vectors per thread
        int src[EPT];                             // H index of the entry's first sample, or -1 (never read with non-zero weights)
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const int e = st + k * STG_THREADS;
            src[k] = -1;
            if (e < NE) {
                const int arr = e / (EP_ROWS * PE), rem = e % (EP_ROWS * PE), ri = rem / PE, xj = rem % PE;
                const int cpar = arr >> 1, rpar = arr & 1, t = 2 * ri + rpar, sc = 2 * xj + cpar;
                // rows the MMAs read: array (0, 0) ri <= 17, (0, 1) and (1, 1) ri <= 16, (1, 0) ri in 1..16; columns: cpar 0 xj <= 17, cpar 1 xj <= 15
                const bool used = cpar == 0 ? (rpar == 0 || ri <= 16) : (xj <= 15 && ri <= 16 && (rpar == 1 || ri >= 1));
                if (used && t < RAW_ROWS) src[k] = t * RAW_COLS + sc + 6; // first sample X = 2*ox0 - 2 + sc  <->  H column sc + 6
            }
        }
        auto load_window = [&](int u, uint4 (&vo)[VPT], uint4 (&vp)[VPT]) {
            // rows Y = 2*oy0 - 2 + t (t < 35), columns X = 2*ox0 - 8 + c (c < 48); outside the block = zero padding
            const int ctu = u / UPI, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            const CtuDev d = p.ctus[ctu];
#pragma unroll
            for (int k = 0; k < VPT; k++) {
                const int i = st + k * STG_THREADS, t = i / 6, vx = i % 6;
                const int Y = 2 * oy0 - 2 + t, X = 2 * ox0 - 8 + vx * 8;
                vo[k] = vp[k] = make_uint4(0, 0, 0, 0);
                if (i < NV && Y >= 0 && Y < S && X >= 0 && X < S && !(p.dbg & 2)) {
                    vo[k] = __ldg(reinterpret_cast<const uint4 *>(d.org + (size_t)Y * d.org_stride + X));
                    vp[k] = __ldg(reinterpret_cast<const uint4 *>(d.pred + (size_t)Y * d.pred_stride + X));
                }
            }
        };
        uint4 vo[VPT], vp[VPT];
        if (u_first < total_units) load_window(u_first, vo, vp);
        const float *cw = reinterpret_cast<const float *>(smem + OFF_CORRW);
        uint32_t it = 0;
        for (int u = u_first; u < total_units; u += u_step, it++) {
            const int oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            // this CTA's unit index ul = it * GROUPS + grp uses buffer ul & 1 for the (ul >> 1)-th time
            const uint32_t ul = it * GROUPS + grp;
            if (p.bw) {
                // two H planes per group: the plane of unit it - 2 is free once the border warp is done with it (every stager passed
                // the barrier below in unit it - 1, i.e. finished its own gather of unit it - 2)
                H = reinterpret_cast<uint32_t *>(smem + OFF_RAW + (grp * 2 + (it & 1)) * RAW_BYTES);
                bwait(&h_empty[grp * 2 + (it & 1)], ((it >> 1) & 1) ^ 1);
            } else {
                group_sync(); // every stager of the group is done reading the previous unit's H
            }
#pragma unroll
            for (int k = 0; k < VPT; k++) {
                const int i = st + k * STG_THREADS;
                if (i < NV && !(p.dbg & 32)) {
                    const uint32_t ow[4] = {vo[k].x, vo[k].y, vo[k].z, vo[k].w}, pw[4] = {vp[k].x, vp[k].y, vp[k].z, vp[k].w};
                    uint32_t hv[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t o = (ow[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu, pp = (pw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
                        const uint32_t co = o < 1023u ? o : 1023u;   // clamp(v / 1023, 0, 1) == min(v, 1023) / 1023 (EncCu.cpp:848-867)
                        const uint32_t ad = s5_absdiff16(o, pp);     // cv::absdiff on the (uint16_t) casts (EncCu.cpp:816,827,833)
                        const uint32_t cr = ad < 1023u ? ad : 1023u;
                        const __half2 h = __floats2half2_rn((float)co * 0.0009765625f, (float)cr * 0.0009765625f); // exact
                        hv[q] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(H + (size_t)i * 8);
                    dst[0] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
                    dst[1] = make_uint4(hv[4], hv[5], hv[6], hv[7]);
                }
            }
            group_sync();
            if (p.bw && lane == 0) mbar_arrive(&h_full[grp * 2 + (it & 1)]); // H is complete: the border warp may start
            if (u + u_step < total_units) load_window(u + u_step, vo, vp); // prefetch: lands while we gather
            const uint32_t buf = ul & 1;
            bwait(&ep_empty[buf], ((ul >> 1) & 1) ^ 1);
            uint8_t *ep = smem + OFF_EP + buf * EP_BYTES;
#pragma unroll
            for (int k = 0; k < EPT; k++) {
                if (src[k] >= 0 && !(p.dbg & 4)) {
                    const uint2 *hp = reinterpret_cast<const uint2 *>(H + (src[k] & ~1));
                    const uint2 w0 = hp[0], w1 = hp[1], w2 = hp[2];
                    const bool odd = src[k] & 1;
                    *reinterpret_cast<uint4 *>(ep + (size_t)(st + k * STG_THREADS) * 16) =
                        odd ? make_uint4(w0.y, w1.x, w1.y, w2.x) : make_uint4(w0.x, w0.y, w1.x, w1.y);
                }
            }
            // The epilogue of the unit that used this buffer two units ago must be done before this unit's corr_full arrival (and
            // before corr[buf] is rewritten): corr_full may run at most ONE phase ahead of its waiter.  Without this wait the
            // arrival for use k can land while the epilogue still waits for use k - 1 -- the barrier is then in phase k + 1, whose
            // parity equals that of k - 1, the epilogue's parity wait never returns and the pipeline deadlocks (seen with two stager
            // groups from ~25 units per CTA on; with one group the stagers were always the slowest role, so it stayed latent).
            const bool early = p.bw || (p.dbg & 2048) != 0; // dbg 2048: experiment (results valid), EP goes to the issuer BEFORE the border terms
            if (early) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ep_full[buf]);
            }
            if (p.bw) continue;
            if (!(p.dbg & 512)) bwait(&d_empty[buf], ((ul >> 1) & 1) ^ 1);
            if ((oy0 == 0 || ox0 == 0) && st < 128 && !(p.dbg & 512)) {
                // border terms (see the header): slot ps < 16 = output (0, ox0 + ps), top; ps >= 16 = output (oy0 + ps - 16, 0), left
                const int ps = st >> 2, cg = (st & 3) * 8;
                const bool top = ps < 16;
                if (top ? oy0 == 0 : ox0 == 0) {
                    float acc[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) acc[k] = 0.0f;
                    const float *wv = cw + (top ? 0 : 5 * 2 * 32);
#pragma unroll
                    for (int e = 0; e < 5; e++) {
                        // top: in(0, 2*(ox0+j) - 2 + e) = H[2][2j + 6 + e];  left: in(2*(oy0+i) - 2 + e, 0) = H[2i + e][8]
                        const int hidx = top ? 2 * RAW_COLS + 2 * ps + 6 + e : (2 * (ps - 16) + e) * RAW_COLS + 8;
                        const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(H + hidx));
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[k] = fmaf(wv[(e * 2 + 0) * 32 + cg + k], xv.x, fmaf(wv[(e * 2 + 1) * 32 + cg + k], xv.y, acc[k]));
                    }
                    if (!top && ps == 16 && oy0 == 0) { // conv1(-1, -1)'s share sits in both terms of output (0, 0): take it out of this one
                        const float2 xv = __half22float2(*reinterpret_cast<const __half2 *>(H + 2 * RAW_COLS + 8));
                        const float *wc = cw + 2 * 5 * 2 * 32;
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[k] -= fmaf(wc[cg + k], xv.x, wc[32 + cg + k] * xv.y);
                    }
                    float4 *dst = reinterpret_cast<float4 *>(smem + OFF_CORR + buf * CORR_BYTES + (ps * 32 + cg) * 4);
                    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            }
            if (!early) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (!early) mbar_arrive(&ep_full[buf]);
                mbar_arrive(&corr_full[buf]);
            }
        }
    } else if (warp == W_MMA) {
        // ======================= MMA issuer: 2 halves x (5 MMAs of the composite 5x5 stride-2 conv + 2 MMAs of conv1 at even positions)
        constexpr uint32_t idesc = umma_idesc_f16(128, 32);
        constexpr uint32_t a_hi = umma_desc_hi(PE * 16), b_hi = umma_desc_hi(128);
        const uint32_t sW = smem_u32(smem + OFF_W);
        const int my_units = total_units > (int)blockIdx.x ? (total_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        for (int v = 0; v < my_units; v++) {
            const uint32_t buf = v & 1;
            bwait(&ep_full[buf], (v >> 1) & 1);
            bwait(&d_empty[buf], ((v >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t ep = sEP + buf * EP_BYTES;
#pragma unroll
                for (int half = 0; half < ((p.dbg & 8) ? 0 : 2); half++) { // left / right 8 output columns
                    const uint32_t d5 = tmem + buf * 128 + half * 32, dq = d5 + 64;
#pragma unroll
                    for (int dy = 0; dy < 5; dy++) {
                        // array (cpar 0, rpar dy & 1), entry row i + dy / 2, entries j and j + 2 (LBO = 2 entries)
                        const uint32_t a_lo = umma_desc_lo(ep + (dy & 1) * EP_ARR + ((dy >> 1) * PE + half * 8) * 16, 32);
                        umma_f16(d5, umma_desc_pack(a_lo, a_hi), umma_desc_pack(umma_desc_lo(sW + dy * 1024, 512), b_hi), idesc, dy != 0);
                    }
                    // conv1 at (2 oy, 2 ox): kernel rows kh = 0 / 2 are entry rows i / i + 1 of array (1, 1), kh = 1 is row i + 1 of array (1, 0)
                    const uint32_t q0 = umma_desc_lo(ep + 3 * EP_ARR + (half * 8) * 16, PE * 16);
                    const uint32_t q1 = umma_desc_lo(ep + 2 * EP_ARR + (PE + half * 8) * 16, PE * 16);
                    umma_f16(dq, umma_desc_pack(q0, a_hi), umma_desc_pack(umma_desc_lo(sW + 5 * 1024, 512), b_hi), idesc, 0);
                    umma_f16(dq, umma_desc_pack(q1, a_hi), umma_desc_pack(umma_desc_lo(sW + 6 * 1024, 512), b_hi), idesc, 1);
                }
                if (p.dbg & 64) { mbar_arrive(&d_full[buf]); mbar_arrive(&ep_empty[buf]); } // timing experiment: plain arrives instead of commits
                else {
                    umma_commit(&d_full[buf]);
                    umma_commit(&ep_empty[buf]);
                }
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue: (W5 result + bias - border terms) -> ReLU -> fp16 -> act1;  conv1 quarter -> fp16 -> act0q
        const int wq = warp & 3, half = warp >> 2, m = wq * 32 + lane, r = m >> 3, c = m & 7, j = half * 8 + c;
        const float *bias_s = reinterpret_cast<const float *>(smem + OFF_BIAS);
        uint32_t ul = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, ul++) {
            const int ctu = u / UPI, oy0 = ((u % UPI) / UW) * 16, ox0 = ((u % UPI) % UW) * 16;
            const uint32_t buf = ul & 1;
            bwait(&corr_full[buf], (ul >> 1) & 1);
            bwait(&d_full[buf], (ul >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(wq * 32) << 16) + buf * 128 + half * 32;
            uint32_t v[32];
            if (p.dbg & 16) { // timing experiment: no TMEM reads, no math, no stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[buf]);
                continue;
            }
            tmem_ld32(tbase, v);
            tmem_ld_wait();
            const float *ct = reinterpret_cast<const float *>(smem + OFF_CORR + buf * CORR_BYTES) + j * 32;
            const float *cl = reinterpret_cast<const float *>(smem + OFF_CORR + buf * CORR_BYTES) + (16 + r) * 32;
            const bool top = oy0 == 0 && r == 0, left = ox0 == 0 && j == 0;
            const __half2 zero2 = __float2half2_rn(0.0f);
            {
                __half *op = p.act1 + out_off(ctu, oy0 + r, ox0 + j);
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
                    if (!(p.dbg & 1)) *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov;
                }
            }
            tmem_ld32(tbase + 64, v);
            tmem_ld_wait();
            {
                __half *op = p.act0q + out_off(ctu, oy0 + r, ox0 + j);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint4 ov;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                    for (int k = 0; k < 4; k++) h2[k] = __floats2half2_rn(__uint_as_float(v[q * 8 + 2 * k]), __uint_as_float(v[q * 8 + 2 * k + 1]));
                    if (!(p.dbg & 1)) *reinterpret_cast<uint4 *>(op + (size_t)q * out_chunk) = ov;
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

static int stem5_dbg()
{
    static const int v = getenv("MLT_STEM5_DBG") ? atoi(getenv("MLT_STEM5_DBG")) : 0;
    return v;
}

static int stem5_bw()
{
    static const int v = getenv("MLT_STEM5_BW") ? atoi(getenv("MLT_STEM5_BW")) : 1; // default on; 0 = border terms on the stagers (A/B)
    return v != 0;
}

static int stem5_nstg()
{
    static const int v = getenv("MLT_STEM5_STAGERS") ? atoi(getenv("MLT_STEM5_STAGERS")) : STEM5_DEFAULT_STAGERS; // measurement override: 4 / 6 / 8
    return v == 4 || v == 6 || v == 8 ? v : STEM5_DEFAULT_STAGERS;
}

template <int S>
static cudaError_t stem5_launch(const Stem5Params &p, int grid, cudaStream_t s)
{
    switch (stem5_nstg()) {
    case 4: return launch_pdl(stem5_umma_kernel<S, 4>, dim3(grid), dim3(stem5::nthreads(4)), stem5::SMEM_BYTES, s, p);
    case 6: return launch_pdl(stem5_umma_kernel<S, 6>, dim3(grid), dim3(stem5::nthreads(6)), stem5::SMEM_BYTES, s, p);
    default: return launch_pdl(stem5_umma_kernel<S, 8>, dim3(grid), dim3(stem5::nthreads(8)), stem5::SMEM_BYTES, s, p);
    }
}

template <int S>
static cudaError_t stem5_attr()
{
    cudaError_t e = cudaFuncSetAttribute(stem5_umma_kernel<S, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem5::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem5_umma_kernel<S, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem5::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem5_umma_kernel<S, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, stem5::SMEM_BYTES);
    return e;
}

cudaError_t stem5_umma_init()
{
    cudaError_t e = stem5_attr<128>();
    if (e == cudaSuccess) e = stem5_attr<64>();
    if (e == cudaSuccess) e = stem5_attr<32>();
    return e;
}

static int stem5_grid(int units, int num_sms)
{
    static const int per_sm = getenv("MLT_STEM5_CTAS") ? atoi(getenv("MLT_STEM5_CTAS")) : 2; // measurement override (1 or 2 CTAs per SM)
    const int g = num_sms * (per_sm == 1 ? 1 : 2);
    return units < g ? units : g;
}

cudaError_t launch_stem5_umma(const CtuDev *ctus, int n, const __half *w, const float *corrw, const float *bias, __half *act0q, __half *act1,
                              int num_sms, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    Stem5Params p{ctus, w, corrw, bias, act0q, act1, n, 0, stem5_bw(), stem5_dbg()};
    return stem5_launch<128>(p, stem5_grid(n * 16, num_sms), s);
}

// the same stem for a 64- or 32-px CU network: act0q / act1 are strips of `cap` images (S/2 x S/2 x 32, not parity-planar)
cudaError_t launch_cu_stem5_umma(int size, const CtuDev *cus, int n, const __half *w, const float *corrw, const float *bias, __half *act0q,
                                 __half *act1, int cap, int num_sms, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    Stem5Params p{cus, w, corrw, bias, act0q, act1, n, cap, stem5_bw(), stem5_dbg()};
    const int grid = stem5_grid(n * (size / 32) * (size / 32), num_sms);
    if (size == 64) return stem5_launch<64>(p, grid, s);
    if (size == 32) return stem5_launch<32>(p, grid, s);
    return cudaErrorInvalidValue;
}

} // namespace mlt
