// conv_simt.cu -- fp32 CUDA-core convolution: the GPU-side CROSS-CHECK engine (mlt_set_engine(ctx, 1)).
// It pins layer semantics (BN folding, shortcut, residual order; arch.py:52-57) to ~1e-5 of the fp32 CPU restatement used by the tests
// and lets tests compare the tcgen05 engine layer by layer ON the device.  It is never selected implicitly
// and is not a fallback: the product path is conv_umma.cuh.
#include "mlt_internal.h"

namespace mlt {

// One thread = one output pixel x 8 consecutive output channels.  NHWC fp32, weights [k*k][cin][cout].
template <int KS>
__global__ void __launch_bounds__(256) conv_simt_kernel(const float *__restrict__ in, const float *__restrict__ w,
                                                        const float *__restrict__ bias, const float *__restrict__ res,
                                                        float *__restrict__ out, int nimg, int hin, int cin, int cout,
                                                        int stride, int relu)
{
    const int hout = hin / stride, groups = cout / 8;
    const size_t total = (size_t)nimg * hout * hout * groups;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const size_t pix = idx / groups;
    const int ox = (int)(pix % hout), oy = (int)((pix / hout) % hout), img = (int)(pix / ((size_t)hout * hout));
    constexpr int PAD = KS / 2;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = bias[g * 8 + i];
    for (int ky = 0; ky < KS; ky++) {
        const int iy = oy * stride + ky - PAD;
        if (iy < 0 || iy >= hin) continue;
        for (int kx = 0; kx < KS; kx++) {
            const int ix = ox * stride + kx - PAD;
            if (ix < 0 || ix >= hin) continue;
            const float *ip = in + (((size_t)img * hin + iy) * hin + ix) * cin;
            const float *wp = w + ((size_t)(ky * KS + kx) * cin) * cout + g * 8;
            for (int c = 0; c < cin; c++) {
                const float v = __ldg(ip + c);
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)c * cout));
                const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)c * cout + 4));
                acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
                acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
                acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
            }
        }
    }
    const size_t o = pix * cout + g * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float v = acc[i];
        if (res) v += res[o + i];
        if (relu) v = fmaxf(v, 0.0f);
        out[o + i] = v;
    }
}

cudaError_t launch_conv_simt(const float *in, const float *w, const float *bias, const float *res, float *out, int nimg,
                             int hin, int cin, int cout, int ksize, int stride, int relu, cudaStream_t s)
{
    const int hout = hin / stride;
    const size_t total = (size_t)nimg * hout * hout * (cout / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (ksize == 3)
        conv_simt_kernel<3><<<blocks, 256, 0, s>>>(in, w, bias, res, out, nimg, hin, cin, cout, stride, relu);
    else
        conv_simt_kernel<1><<<blocks, 256, 0, s>>>(in, w, bias, res, out, nimg, hin, cin, cout, stride, relu);
    return cudaGetLastError();
}

} // namespace mlt
