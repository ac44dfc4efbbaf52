// head.cu -- prediction head: three global average pools (arch.py:282,288,294), cat(poc, qp) + Linear
// (arch.py:284-297), per-level softmax, argmax of every level (EncCu.cpp:913-921 uses level 3) and the
// per-level split flags.  One block per CTU; all reductions run in a fixed order (smem partials +
// warp shuffles), so results are bit-reproducible run to run -- an encoder must be deterministic.
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&f)[8]);

template <>
__device__ __forceinline__ void load8<__half>(const __half *p, float (&f)[8])
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    const __half2 *h2 = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const float2 t = __half22float2(h2[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&f)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// mean over the H x H pixels of one image -> feat[C]; 256 threads, deterministic (fixed summation order).
// fp32 cross-check engine: dense NHWC.  Product path: chunk-planar fp16 (conv_umma.cuh ActLayout) -- the thread
// group of one 8-channel chunk walks that chunk's pixels (16 B apart) in order.
template <int H, int C>
__device__ __forceinline__ void gap_nhwc(const float *act, float *partial /*[256/(C/8)][C]*/, float *feat)
{
    constexpr int P = H * H, CH = C / 8, PH = 256 / CH;
    const int cj = threadIdx.x % CH, pp = threadIdx.x / CH;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = pp; p < P; p += PH) {
        float f[8];
        load8<float>(act + (size_t)p * C + cj * 8, f);
#pragma unroll
        for (int e = 0; e < 8; e++) s[e] += f[e];
    }
#pragma unroll
    for (int e = 0; e < 8; e++) partial[pp * C + cj * 8 + e] = s[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.0f;
        for (int k = 0; k < PH; k++) t += partial[k * C + c];
        feat[c] = t * (1.0f / (float)(H * H));
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) head_kernel(const HeadParams p)
{
    __shared__ float partial[2048];
    __shared__ float feat[64 + 128 + 256];
    __shared__ float logits[9];
    const int n = blockIdx.x;
    griddep_launch_dependents();
    griddep_wait(); // PDL: activations come from the previous kernel
    if constexpr (sizeof(T) == 2) {
        // product path: the epilogues of the last conv of each stage (conv_umma.cuh GAP) left per-tile partial sums
        // [tile * NB + sub-image][4 lane quadrants][C]; add the few partials of this image in a fixed order
        for (int c = threadIdx.x; c < 448; c += 256) {
            float t = 0.0f;
            if (c < 64) { // layer1: 32x32 = 8 tiles of 16x8 per image
                const float *g = p.gap_part[0] + (size_t)n * 8 * 4 * 64 + c;
                for (int k = 0; k < 32; k++) t += g[k * 64];
                t *= 1.0f / (32 * 32);
            } else if (c < 192) { // layer2: 16x16 = 2 tiles per image
                const float *g = p.gap_part[1] + (size_t)n * 2 * 4 * 128 + (c - 64);
                for (int k = 0; k < 8; k++) t += g[k * 128];
                t *= 1.0f / (16 * 16);
            } else { // layer3: one tile per image pair, sub-image n & 1
                const float *g = p.gap_part[2] + (size_t)n * 4 * 256 + (c - 192);
                for (int k = 0; k < 4; k++) t += g[k * 256];
                t *= 1.0f / (8 * 8);
            }
            feat[c] = t;
        }
        __syncthreads();
    } else {
        gap_nhwc<32, 64>(static_cast<const float *>(p.act[0]) + (size_t)n * 32 * 32 * 64, partial, feat);
        gap_nhwc<16, 128>(static_cast<const float *>(p.act[1]) + (size_t)n * 16 * 16 * 128, partial, feat + 64);
        gap_nhwc<8, 256>(static_cast<const float *>(p.act[2]) + (size_t)n * 8 * 8 * 256, partial, feat + 192);
    }

    const float poc = (float)p.ctus[n].poc, qp = (float)p.ctus[n].qp; // raw ints promoted by torch.cat (arch.py:274-275)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < 9; o += 8) {
        const int lvl = o < 2 ? 0 : (o < 5 ? 1 : 2);
        const int row = o - (lvl == 0 ? 0 : (lvl == 1 ? 2 : 5));
        const int C = 64 << lvl;
        const float *f = feat + (lvl == 0 ? 0 : (lvl == 1 ? 64 : 192));
        const float *wr = p.fc_w[lvl] + (size_t)row * (C + 2);
        float s = 0.0f;
        for (int k = lane; k < C; k += 32) s = fmaf(wr[k], f[k], s);
        if (lane == 0) s = fmaf(wr[C], poc, s);
        if (lane == 1) s = fmaf(wr[C + 1], qp, s);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) logits[o] = s + p.fc_b[lvl][row];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mlt_result r;
        int arg[3];
        const int base[3] = {0, 2, 5}, cnt[3] = {2, 3, 4};
        for (int l = 0; l < 3; l++) {
            float mx = logits[base[l]];
            int am = 0;
            for (int k = 1; k < cnt[l]; k++)
                if (logits[base[l] + k] > mx) { mx = logits[base[l] + k]; am = k; } // first maximum wins (torch.argmax)
            float e[4], sum = 0.0f;
            for (int k = 0; k < cnt[l]; k++) { e[k] = expf(logits[base[l] + k] - mx); sum += e[k]; }
            for (int k = 0; k < cnt[l]; k++) r.probs[base[l] + k] = e[k] / sum;
            arg[l] = am;
        }
        for (int k = 0; k < 9; k++) r.logits[k] = logits[k];
        r.split_l1 = arg[0];
        r.split_l2 = arg[1];
        r.split_l3 = arg[2];
        uint32_t fl = 0;
        if (arg[0] == 1) fl |= MLT_FLAG_L1_SPLIT;
        if (arg[1] == 1) fl |= MLT_FLAG_L2_QT;
        if (arg[1] == 2) fl |= MLT_FLAG_L2_MTT;
        if (arg[2] == 1) fl |= MLT_FLAG_L3_QT;
        if (arg[2] == 2) fl |= MLT_FLAG_L3_BT_H;
        if (arg[2] == 3) fl |= MLT_FLAG_L3_BT_V;
        const bool cons = (arg[0] == 0 && arg[1] == 0 && arg[2] == 0) || (arg[0] == 1 && arg[1] == 1 && arg[2] == 1) ||
                          (arg[0] == 1 && arg[1] == 2 && arg[2] >= 2);
        if (cons) fl |= MLT_FLAG_CONSISTENT;
        r.flags = fl;
        p.out[n] = r;
    }
}

cudaError_t launch_head_h(const HeadParams &p, cudaStream_t s)
{
    return launch_pdl(head_kernel<__half>, dim3(p.n), dim3(256), 0, s, p);
}
cudaError_t launch_head_f(const HeadParams &p, cudaStream_t s)
{
    head_kernel<float><<<p.n, 256, 0, s>>>(p);
    return cudaGetLastError();
}

} // namespace mlt
