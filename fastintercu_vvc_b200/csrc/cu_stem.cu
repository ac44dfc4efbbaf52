// cu_stem.cu -- input staging (EncCu.cpp:810-867 at cuw = cuh = 64 / 32 / 16) + the first convolution of the smaller-CU
// networks (mlt_cu_or_pq_arch.py:105: conv1, 2 -> 32 channels, 3x3, stride 1, pad 1, no BN / activation) on tcgen05.
//
// Same formulation as conv1_umma_kernel (stage_conv1.cu) -- the integer part of the staging is exact ((uint16) cast,
// |org - pred|, clamp at 1023 == clamp of v / 1023 to [0, 1]); the A operand is v * 2^-10 (exact in fp16), the B operand
// the hi / lo fp16 split of w * (float)(1/1023) * 2^10, fp32 accumulation in TMEM -- generalised over the CU size:
// one CTA = 4 MMA tiles of 16 x 8 pixels = 32 / SW sub-strips of 16 rows x SW columns (SW = min(size, 32)), i.e. half a
// row-strip of a 64-px CU, one row-strip of a 32-px CU or two whole 16-px CUs.  Small CTAs on purpose: the kernel is a
// chain of latencies (global loads -> expansion -> 16 MMAs -> TMEM reads -> stores), so throughput comes from CTAs in
// flight, and 128 TMEM columns per CTA let four of them overlap on an SM (8 tiles / 256 columns: two; measured 342 us
// for 3840 64-px CUs, 288 us with 4 tiles; 2 tiles / 64 columns / 128 threads measured no faster: what remains is the
// scattered 64-byte store segments of the strip layout on a 64x64 map, see profiles/r01/README.md).
// Output: activation 0, fp16, parity-planar STRIP layout [plane][4 chunks][size/2][cap][size/2][8] (conv_umma.cuh).
#include "conv_umma.cuh"
#include "mlt_internal.h"

namespace mlt {

template <int S>
__global__ void __launch_bounds__(256) cu_conv1_umma_kernel(const CtuDev *__restrict__ cus, int n, const __half *__restrict__ wop,
                                                           __half *__restrict__ out, int cap)
{
    constexpr int NT = 4;               // MMA tiles per CTA
    constexpr int SW = S < 32 ? S : 32; // sub-strip width
    constexpr int NSUB = 32 / SW;       // sub-strips per CTA
    constexpr int TPS = SW / 8;         // MMA tiles per sub-strip
    constexpr int CST = S / SW;         // column strips per row strip
    constexpr int SPI = (S / 16) * CST; // sub-strips (16 rows x SW columns) per CU
    constexpr int INC = SW + 16;        // staged input columns x = -8 .. SW + 7 (16-byte aligned loads)
    constexpr int E_BYTES = 19 * SW * 16;
    static_assert(S == 64 || S == 32 || S == 16, "CU size");
    __shared__ __align__(128) uint8_t s_e[NSUB * E_BYTES + 256]; // expanded patches (+ slack: tile reads never leave the array)
    __shared__ __align__(128) uint8_t s_w[2 * 4 * 32 * 16];      // [hi, lo][4 chunks][32 cout][8]
    __shared__ __align__(16) int16_t s_in[NSUB][2][18][INC];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, NT * 32); tmem_relinquish(); }
    reinterpret_cast<uint4 *>(s_w)[tid] = __ldg(reinterpret_cast<const uint4 *>(wop) + tid); // 256 x 16 B
    // raw samples of every sub-strip: 2 planes x 18 rows x INC/8 vectors; outside the CU = conv zero padding (org = pred = 0)
    constexpr int VPR = INC / 8, PER_SUB = 2 * 18 * VPR;
    for (int i = tid; i < NSUB * PER_SUB; i += 256) {
        const int sub = i / PER_SUB, rem = i % PER_SUB;
        const int plane = rem / (18 * VPR), row = (rem / VPR) % 18, v = rem % VPR;
        const int g = blockIdx.x * NSUB + sub, cu = g / SPI, y0 = ((g % SPI) / CST) * 16, x0 = ((g % SPI) % CST) * SW;
        const int y = y0 - 1 + row, x = x0 - 8 + v * 8;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (cu < n && y >= 0 && y < S && x >= 0 && x < S) {
            const CtuDev d = cus[cu];
            const int16_t *src = plane ? d.pred + (size_t)y * d.pred_stride : d.org + (size_t)y * d.org_stride;
            val = __ldg(reinterpret_cast<const uint4 *>(src + x));
        }
        *reinterpret_cast<uint4 *>(&s_in[sub][plane][row][v * 8]) = val;
    }
    __syncthreads();
    // expanded patch: entry (row i, px x) = fp16 {org, res} of input pixels x-1 .. x+2 of input row y0-1+i; row 18 = 0
    constexpr int QPR = SW / 4;
    for (int i = tid; i < NSUB * 19 * QPR; i += 256) {
        const int sub = i / (19 * QPR), rem = i % (19 * QPR);
        const int row = rem / QPR, xq = (rem % QPR) * 4;
        uint4 *dst = reinterpret_cast<uint4 *>(s_e + (size_t)sub * E_BYTES + ((size_t)row * SW + xq) * 16);
        if (row == 18) {
#pragma unroll
            for (int j = 0; j < 4; j++) dst[j] = make_uint4(0, 0, 0, 0);
            continue;
        }
        __half2 px[7]; // (org, res) * 2^-10 of input columns xq-1 .. xq+5
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const uint32_t o = (uint16_t)s_in[sub][0][row][8 + xq - 1 + j], pp = (uint16_t)s_in[sub][1][row][8 + xq - 1 + j];
            const uint32_t vo = o < 1023u ? o : 1023u; // clamp(v / 1023, 0, 1) == min(v, 1023) / 1023 (EncCu.cpp:848-867)
            const uint32_t ad = o > pp ? o - pp : pp - o; // cv::absdiff on CV_16U (EncCu.cpp:833)
            const uint32_t vr = ad < 1023u ? ad : 1023u;
            px[j] = __floats2half2_rn((float)vo * 0.0009765625f, (float)vr * 0.0009765625f); // exact: <= 10 significant bits
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint4 v;
            __half2 *h2 = reinterpret_cast<__half2 *>(&v);
            h2[0] = px[j]; h2[1] = px[j + 1]; h2[2] = px[j + 2]; h2[3] = px[j + 3];
            dst[j] = v;
        }
    }
    if (tid < 16) reinterpret_cast<uint4 *>(s_e + NSUB * E_BYTES)[tid] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 32);
            constexpr uint32_t a_hi = umma_desc_hi(SW * 16), b_hi = umma_desc_hi(128);
            const uint32_t sE = smem_u32(s_e), sW = smem_u32(s_w);
#pragma unroll
            for (int tile = 0; tile < NT; tile++) {
                const uint32_t base = sE + (tile / TPS) * E_BYTES + (tile % TPS) * 128;
                const uint32_t a0 = umma_desc_lo(base, SW * 16);               // chunks kh = 0, 1 (patch rows r, r + 1)
                const uint32_t a2 = umma_desc_lo(base + 2 * SW * 16, SW * 16); // chunks kh = 2, (3: zero weights)
#pragma unroll
                for (int part = 0; part < 2; part++) { // hi, lo halves of the weights
                    const uint32_t b0 = umma_desc_lo(sW + part * 2048, 32 * 16), b2 = umma_desc_lo(sW + part * 2048 + 1024, 32 * 16);
                    umma_f16(tmem + tile * 32, umma_desc_pack(a0, a_hi), umma_desc_pack(b0, b_hi), idesc, part);
                    umma_f16(tmem + tile * 32, umma_desc_pack(a2, a_hi), umma_desc_pack(b2, b_hi), idesc, 1);
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    // epilogue: warp w reads TMEM lane quadrant w % 4 (pixels), tiles (w / 4) * 2, + 1
    {
        const int wq = warp & 3, m = wq * 32 + lane, r = m >> 3, c = m & 7;
        constexpr int HP = S / 2;
        const size_t chunk = (size_t)HP * cap * HP * 8, plane_stride = chunk * 4; // [plane][4 chunks][HP][cap][HP][8]
#pragma unroll 1
        for (int tile = (warp >> 2) * (NT / 2); tile < (warp >> 2) * (NT / 2) + NT / 2; tile++) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + tile * 32, v);
            tmem_ld_wait();
            const int g = blockIdx.x * NSUB + tile / TPS, cu = g / SPI;
            const int oy = ((g % SPI) / CST) * 16 + r, ox = ((g % SPI) % CST) * SW + (tile % TPS) * 8 + c;
            if (cu < n) {
                __half *op = out + (size_t)((oy & 1) * 2 + (ox & 1)) * plane_stride + ((size_t)((oy >> 1) * cap + cu) * HP + (ox >> 1)) * 8;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint4 ov;
                    __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                    for (int e = 0; e < 4; e++) h2[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + e * 2]), __uint_as_float(v[q * 8 + e * 2 + 1]));
                    *reinterpret_cast<uint4 *>(op + (size_t)q * chunk) = ov;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, NT * 32); }
}

cudaError_t launch_cu_conv1(int size, const CtuDev *cus, int n, const __half *wop, __half *out, int cap, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const int sw = size < 32 ? size : 32, spi = (size / 16) * (size / sw), nsub = 32 / sw;
    const int grid = (n * spi + nsub - 1) / nsub;
    switch (size) {
    case 64: cu_conv1_umma_kernel<64><<<grid, 256, 0, s>>>(cus, n, wop, out, cap); break;
    case 32: cu_conv1_umma_kernel<32><<<grid, 256, 0, s>>>(cus, n, wop, out, cap); break;
    case 16: cu_conv1_umma_kernel<16><<<grid, 256, 0, s>>>(cus, n, wop, out, cap); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace mlt
