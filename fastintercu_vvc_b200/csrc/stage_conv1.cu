// stage_conv1.cu -- input staging (EncCu.cpp:810-877) and the first convolution (arch.py:278).
//
//   staging : Pel (int16) org / pred  ->  (uint16) cast, |org - pred|, * (float)(1/1023), clamp [0,1]
//   conv1   : 3x3, 2 -> 32 channels, stride 1, pad 1, no BN, no activation
//
// The product path never materialises the fp32 [2][128][128] tensor the reference ships to the GPU
// (128 KiB H2D per CTU, EncCu.cpp:875): `stage_conv1_kernel` reads the int16 samples with 128-bit
// loads, normalises them in registers with exactly the reference's fp32 arithmetic and feeds conv1
// directly.  `stage_kernel` writes that tensor out only for the bit-exactness probe (mlt_debug_stage).
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

    // fp16 product path: haloed NHWC [n+1][130][130][32] (conv_umma.cuh); fp32 cross-check path: dense NHWC
    const size_t pix0 = sizeof(OutT) == 2 ? ((size_t)ctu * 130 + (y0 + row + 1)) * 130 + (x0 + sx + 1)
                                          : ((size_t)ctu * 128 + (y0 + row)) * 128 + (x0 + sx);
#pragma unroll
    for (int px = 0; px < 4; px++) {
        OutT *o = out + (pix0 + px) * 32 + half * 16;
        if constexpr (sizeof(OutT) == 2) {
            uint4 v[2];
            __half2 *h2 = reinterpret_cast<__half2 *>(v);
#pragma unroll
            for (int e = 0; e < 8; e++) h2[e] = __floats2half2_rn(acc[px][2 * e], acc[px][2 * e + 1]);
            reinterpret_cast<uint4 *>(o)[0] = v[0];
            reinterpret_cast<uint4 *>(o)[1] = v[1];
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                reinterpret_cast<float4 *>(o)[q] = make_float4(acc[px][4 * q], acc[px][4 * q + 1], acc[px][4 * q + 2], acc[px][4 * q + 3]);
        }
    }
}

cudaError_t launch_stage(const CtuDev *ctus, int n, float *out, cudaStream_t s)
{
    stage_kernel<<<n, 256, 0, s>>>(ctus, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_conv1_h(const CtuDev *ctus, int n, const float *w, __half *out, cudaStream_t s)
{
    stage_conv1_kernel<__half><<<n * 32, 256, 0, s>>>(ctus, w, out);
    return cudaGetLastError();
}
cudaError_t launch_stage_conv1_f(const CtuDev *ctus, int n, const float *w, float *out, cudaStream_t s)
{
    stage_conv1_kernel<float><<<n * 32, 256, 0, s>>>(ctus, w, out);
    return cudaGetLastError();
}

// debug: haloed NHWC fp16 [nimg][h+2][h+2][c] -> dense NHWC fp32 [nimg][h][h][c]
__global__ void unhalo_to_float_kernel(const __half *__restrict__ in, float *__restrict__ out, int h, int c, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const size_t pix = i / c;
        const int x = (int)(pix % h), y = (int)((pix / h) % h);
        const size_t img = pix / ((size_t)h * h);
        out[i] = __half2float(in[((img * (h + 2) + y + 1) * (h + 2) + x + 1) * c + ch]);
    }
}
cudaError_t launch_unhalo_to_float(const __half *in, float *out, int nimg, int h, int c, cudaStream_t s)
{
    const size_t n = (size_t)nimg * h * h * c;
    const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    unhalo_to_float_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(in, out, h, c, n);
    return cudaGetLastError();
}

} // namespace mlt
