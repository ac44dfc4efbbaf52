// stage_conv1.cu -- input staging (EncCu.cpp:810-877) and the first convolution (arch.py:278).
//
//   staging : Pel (int16) org / pred  ->  (uint16) cast, |org - pred|, * (float)(1/1023), clamp [0,1]
//   conv1   : 3x3, 2 -> 32 channels, stride 1, pad 1, no BN, no activation
//
// The product path never materialises the fp32 [2][128][128] tensor the reference ships to the GPU
// (128 KiB H2D per CTU, EncCu.cpp:875):
//   * `conv1_umma_kernel` (product) reads the int16 samples with 128-bit loads, does the integer part of the
//     staging exactly (uint16 cast, |org - pred|, clamp at 1023 == clamp of v/1023 to [0,1]) and runs conv1 on the
//     tensor cores: A = v * 2^-10 (exact in fp16), B = hi/lo fp16 split of w * (float)(1/1023) * 2^10, fp32
//     accumulation in TMEM -- i.e. conv1 at ~fp32 weight precision on exact inputs, output fp16 chunk-planar.
//   * `stage_kernel` writes the normalised fp32 tensor out for the bit-exactness probe (mlt_debug_stage) with
//     exactly the reference's fp32 arithmetic; `stage_conv1_kernel` is the fp32 CUDA-core cross-check engine's conv1.
#include "conv_umma.cuh"
#include "mlt_internal.h"

namespace mlt {

__device__ __forceinline__ float norm1023(uint32_t v)
{
    // cv::Mat::convertTo(CV_32FC1, 1.0/1023, 0): (float)v * (float)alpha, one rounding; then the clamp loops
    const float f = __fmul_rn((float)v, __uint_as_float(MLT_ALPHA_BITS));
    return fminf(fmaxf(f, 0.0f), 1.0f);
}

__device__ __forceinline__ uint32_t absdiff_u16(uint32_t o, uint32_t p) { return o > p ? o - p : p - o; } // cv::absdiff, CV_16U

// ---------------------------------------------------------------------------------------------------
// Probe kernel: one block per CTU, each thread converts 8 samples per pass (one 128-bit load per plane).
__global__ void __launch_bounds__(256) stage_kernel(const CtuDev *__restrict__ ctus, float *__restrict__ out)
{
    const CtuDev d = ctus[blockIdx.x];
    float *o0 = out + (size_t)blockIdx.x * 2 * 128 * 128, *o1 = o0 + 128 * 128;
    for (int v = threadIdx.x; v < 128 * 16; v += 256) {
        const int y = v >> 4, x = (v & 15) * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(d.org + (size_t)y * d.org_stride + x));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(d.pred + (size_t)y * d.pred_stride + x));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        float fo[8], fr[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t o_lo = aw[k] & 0xFFFFu, o_hi = aw[k] >> 16, p_lo = bw[k] & 0xFFFFu, p_hi = bw[k] >> 16;
            fo[2 * k] = norm1023(o_lo);
            fo[2 * k + 1] = norm1023(o_hi);
            fr[2 * k] = norm1023(absdiff_u16(o_lo, p_lo));
            fr[2 * k + 1] = norm1023(absdiff_u16(o_hi, p_hi));
        }
        float4 *q0 = reinterpret_cast<float4 *>(o0 + y * 128 + x), *q1 = reinterpret_cast<float4 *>(o1 + y * 128 + x);
        q0[0] = make_float4(fo[0], fo[1], fo[2], fo[3]);
        q0[1] = make_float4(fo[4], fo[5], fo[6], fo[7]);
        q1[0] = make_float4(fr[0], fr[1], fr[2], fr[3]);
        q1[1] = make_float4(fr[4], fr[5], fr[6], fr[7]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Fused staging + conv1.  One block = 16 x 32 output pixels of one CTU (32 blocks per CTU).
// Thread = 4 consecutive pixels x 16 output channels (64 fp32 accumulators); the two channel halves are
// split across warps so every weight read from shared memory is a warp-wide broadcast.
constexpr int SROW = 48; // int16 per staged smem row: [7] left halo, [8..39] interior (16 B aligned), [40] right halo

template <typename OutT>
__global__ void __launch_bounds__(256) stage_conv1_kernel(const CtuDev *__restrict__ ctus, const float *__restrict__ w,
                                                         OutT *__restrict__ out)
{
    __shared__ __align__(16) int16_t s_org[18][SROW];
    __shared__ __align__(16) int16_t s_pred[18][SROW];
    __shared__ __align__(16) float s_w[9 * 2 * 32];

    const int ctu = blockIdx.x >> 5, t = blockIdx.x & 31;
    const int y0 = (t >> 2) * 16, x0 = (t & 3) * 32;
    const CtuDev d = ctus[ctu];
    const int tid = threadIdx.x;

    for (int i = tid; i < 9 * 2 * 32; i += 256) s_w[i] = w[i];
    if (tid < 144) { // interior: 18 rows x 4 vectors x 2 planes, 128-bit loads
        const int plane = tid / 72, rem = tid % 72, row = rem >> 2, v = rem & 3;
        const int y = y0 - 1 + row;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < 128) {
            const int16_t *src = plane ? d.pred + (size_t)y * d.pred_stride : d.org + (size_t)y * d.org_stride;
            val = __ldg(reinterpret_cast<const uint4 *>(src + x0 + v * 8));
        }
        int16_t *dst = plane ? &s_pred[row][8 + v * 8] : &s_org[row][8 + v * 8];
        *reinterpret_cast<uint4 *>(dst) = val;
    } else if (tid < 216) { // halo columns
        const int k = tid - 144, plane = k / 36, rem = k % 36, row = rem >> 1, side = rem & 1;
        const int y = y0 - 1 + row, x = side ? x0 + 32 : x0 - 1;
        int16_t val = 0;
        if (y >= 0 && y < 128 && x >= 0 && x < 128)
            val = plane ? d.pred[(size_t)y * d.pred_stride + x] : d.org[(size_t)y * d.org_stride + x];
        (plane ? s_pred : s_org)[row][side ? 40 : 7] = val;
    }
    __syncthreads();

    const int half = tid >> 7, strip = tid & 127;
    const int row = strip >> 3, sx = (strip & 7) * 4;
    float acc[4][16];
#pragma unroll
    for (int px = 0; px < 4; px++)
#pragma unroll
        for (int co = 0; co < 16; co++) acc[px][co] = 0.0f;

#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
        float in0[6], in1[6];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const uint32_t o = (uint16_t)s_org[row + kh][7 + sx + c], p = (uint16_t)s_pred[row + kh][7 + sx + c];
            in0[c] = norm1023(o);                 // channel 0: org      (EncCu.cpp:838)
            in1[c] = norm1023(absdiff_u16(o, p)); // channel 1: residual (EncCu.cpp:833-836)
        }
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            const float4 *w0 = reinterpret_cast<const float4 *>(&s_w[((kh * 3 + kw) * 2 + 0) * 32 + half * 16]);
            const float4 *w1 = reinterpret_cast<const float4 *>(&s_w[((kh * 3 + kw) * 2 + 1) * 32 + half * 16]);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 a = w0[q], b = w1[q];
#pragma unroll
                for (int px = 0; px < 4; px++) {
                    const float v0 = in0[px + kw], v1 = in1[px + kw];
                    acc[px][q * 4 + 0] = fmaf(v1, b.x, fmaf(v0, a.x, acc[px][q * 4 + 0]));
                    acc[px][q * 4 + 1] = fmaf(v1, b.y, fmaf(v0, a.y, acc[px][q * 4 + 1]));
                    acc[px][q * 4 + 2] = fmaf(v1, b.z, fmaf(v0, a.z, acc[px][q * 4 + 2]));
                    acc[px][q * 4 + 3] = fmaf(v1, b.w, fmaf(v0, a.w, acc[px][q * 4 + 3]));
                }
            }
        }
    }

    // dense NHWC fp32 (cross-check engine only)
    const size_t pix0 = ((size_t)ctu * 128 + (y0 + row)) * 128 + (x0 + sx);
#pragma unroll
    for (int px = 0; px < 4; px++) {
        OutT *o = out + (pix0 + px) * 32 + half * 16;
#pragma unroll
        for (int q = 0; q < 4; q++)
            reinterpret_cast<float4 *>(o)[q] = make_float4(acc[px][4 * q], acc[px][4 * q + 1], acc[px][4 * q + 2], acc[px][4 * q + 3]);
    }
}

cudaError_t launch_stage(const CtuDev *ctus, int n, float *out, cudaStream_t s)
{
    stage_kernel<<<n, 256, 0, s>>>(ctus, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_conv1_f(const CtuDev *ctus, int n, const float *w, float *out, cudaStream_t s)
{
    stage_conv1_kernel<float><<<n * 32, 256, 0, s>>>(ctus, w, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Product path: staging + conv1 on tcgen05.  One CTA = a 16-row x 64-column strip of one CTU = 8 MMA tiles of
// 16 x 8 pixels (M = 128), all 8 accumulators (8 x 32 columns) live in TMEM at once.
//   K layout per tile (K = 32, two K=16 MMAs): chunk kh (kh = 0..2) = 8 fp16 = (org, res) of input pixels
//   x-1 .. x+2 of input row y+kh-1 (the 4th pixel has zero weights); chunk 3 = zero weights.
//   The expanded patch E[19 rows][64 px][8] sits in shared memory once; the three vertical taps are row-shifted
//   windows of it (operand descriptor start address), LBO = one patch row.
// Output: activation 0 in the parity-planar layout (conv_umma.cuh), no bias / BN / ReLU (arch.py:278).
constexpr int C1_SW = 64, C1_ROWS = 19, C1_IN_COLS = 80; // staged input columns x0-8 .. x0+71 (16-byte aligned loads)

__global__ void __launch_bounds__(256) conv1_umma_kernel(const CtuDev *__restrict__ ctus, const __half *__restrict__ wop,
                                                        __half *__restrict__ out)
{
    __shared__ __align__(128) uint8_t s_e[C1_ROWS * C1_SW * 16];
    __shared__ __align__(128) uint8_t s_w[2 * 4 * 32 * 16]; // [hi, lo][4 chunks][32 cout][8]
    __shared__ __align__(16) int16_t s_in[2][18][C1_IN_COLS];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ctu = blockIdx.x >> 4, t = blockIdx.x & 15;
    const int y0 = (t >> 1) * 16, x0 = (t & 1) * C1_SW;
    const CtuDev d = ctus[ctu];

    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
    reinterpret_cast<uint4 *>(s_w)[tid] = __ldg(reinterpret_cast<const uint4 *>(wop) + tid); // 256 x 16 B
    // raw samples: 2 planes x 18 rows x 10 vectors of 8 int16; outside the CTU = conv zero padding (org = pred = 0)
    for (int i = tid; i < 2 * 18 * 10; i += 256) {
        const int plane = i / 180, rem = i % 180, row = rem / 10, v = rem % 10;
        const int y = y0 - 1 + row, x = x0 - 8 + v * 8;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < 128 && x >= 0 && x < 128) {
            const int16_t *src = plane ? d.pred + (size_t)y * d.pred_stride : d.org + (size_t)y * d.org_stride;
            val = __ldg(reinterpret_cast<const uint4 *>(src + x));
        }
        *reinterpret_cast<uint4 *>(&s_in[plane][row][v * 8]) = val;
    }
    __syncthreads();
    // expanded patch: entry (row i, px x) = fp16 {org, res} of input pixels x-1 .. x+2 of input row y0-1+i; row 18 = 0
    for (int i = tid; i < C1_ROWS * (C1_SW / 4); i += 256) {
        const int row = i / (C1_SW / 4), xq = (i % (C1_SW / 4)) * 4;
        uint4 *dst = reinterpret_cast<uint4 *>(s_e + ((size_t)row * C1_SW + xq) * 16);
        if (row == 18) {
#pragma unroll
            for (int j = 0; j < 4; j++) dst[j] = make_uint4(0, 0, 0, 0);
            continue;
        }
        __half2 px[7]; // (org, res) * 2^-10 of input columns x0+xq-1 .. x0+xq+5
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const uint32_t o = (uint16_t)s_in[0][row][8 + xq - 1 + j], pp = (uint16_t)s_in[1][row][8 + xq - 1 + j];
            const uint32_t vo = o < 1023u ? o : 1023u;        // clamp(v / 1023, 0, 1) == min(v, 1023) / 1023 (EncCu.cpp:848-867)
            const uint32_t ad = absdiff_u16(o, pp);
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
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 32);
            constexpr uint32_t a_hi = umma_desc_hi(C1_SW * 16), b_hi = umma_desc_hi(128);
            const uint32_t sE = smem_u32(s_e), sW = smem_u32(s_w);
#pragma unroll
            for (int tile = 0; tile < 8; tile++) {
                const uint32_t a0 = umma_desc_lo(sE + tile * 128, C1_SW * 16);                    // chunks kh = 0, 1
                const uint32_t a2 = umma_desc_lo(sE + 2 * C1_SW * 16 + tile * 128, C1_SW * 16);   // chunks kh = 2, (3: zero weights)
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
    // epilogue: warp w reads TMEM lane quadrant w % 4 (pixels), tiles (w / 4) * 4 .. + 3
    {
        const int wq = warp & 3, m = wq * 32 + lane, r = m >> 3, c = m & 7;
        const int oy = y0 + r;
        constexpr size_t CHUNK = 64 * 64 * 8, PLANE = CHUNK * 4, UNIT = PLANE * 4; // a0: [ctu][4 planes][4 chunks][64][64][8]
#pragma unroll 1
        for (int tile = (warp >> 2) * 4; tile < (warp >> 2) * 4 + 4; tile++) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + tile * 32, v);
            tmem_ld_wait();
            const int ox = x0 + tile * 8 + c;
            __half *op = out + (size_t)ctu * UNIT + (size_t)((oy & 1) * 2 + (ox & 1)) * PLANE + (size_t)((oy >> 1) * 64 + (ox >> 1)) * 8;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 ov;
                __half2 *h2 = reinterpret_cast<__half2 *>(&ov);
#pragma unroll
                for (int e = 0; e < 4; e++) h2[e] = __floats2half2_rn(__uint_as_float(v[q * 8 + e * 2]), __uint_as_float(v[q * 8 + e * 2 + 1]));
                *reinterpret_cast<uint4 *>(op + (size_t)q * CHUNK) = ov;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

cudaError_t launch_conv1_umma(const CtuDev *ctus, int n, const __half *wop, __half *out, cudaStream_t s)
{
    conv1_umma_kernel<<<n * 16, 256, 0, s>>>(ctus, wop, out);
    return cudaGetLastError();
}

// debug: chunk-planar fp16 activation (conv_umma.cuh ActLayout) -> dense NHWC fp32 [nimg][h][h][c]
__global__ void unpack_act_kernel(const __half *__restrict__ in, float *__restrict__ out, ActLayout L, size_t n)
{
    const int h = L.H, c = L.C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const size_t pix = i / c;
        const int x = (int)(pix % h), y = (int)((pix / h) % h);
        const size_t img = pix / ((size_t)h * h);
        const size_t unit = L.strip ? 0 : (L.pair ? img >> 1 : img);
        const int sub = L.strip ? (int)img : (L.pair ? (int)(img & 1) : 0), plane = L.par ? (y & 1) * 2 + (x & 1) : 0;
        const int yy = L.par ? y >> 1 : y, xx = L.par ? x >> 1 : x, hp = L.hp();
        const size_t off = (unit * L.npl() + plane) * (size_t)(c / 8) * L.chunk_stride() + (size_t)(ch / 8) * L.chunk_stride() +
                           (size_t)((yy * L.nimg() + sub) * hp + xx) * 8 + (ch & 7);
        out[i] = __half2float(in[off]);
    }
}
cudaError_t launch_unpack_act(const __half *in, float *out, int nimg, const ActLayout &L, cudaStream_t s)
{
    const size_t n = (size_t)nimg * L.H * L.H * L.C;
    const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    unpack_act_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(in, out, L, n);
    return cudaGetLastError();
}

} // namespace mlt
