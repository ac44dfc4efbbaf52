// cu_head.cu -- prediction head of the smaller-CU networks: four global average pools (after layer1..layer4,
// mlt_cu_or_pq_arch.py:108-127), cat(poc, qp) + Linear (66->2, 98->3, 130->4, 258->6), per-level softmax and argmax
// (the hook uses level 1 below 128x128: EncCu.cpp:916-921).  One block per CU; every reduction runs in a fixed order, so
// results are bit-reproducible run to run and independent of the batch a CU is in.
#include "mlt_internal.h"
#include "ptx.cuh"

namespace mlt {

namespace {
constexpr int NT = 256;
constexpr int FEAT_TOTAL = 64 + 96 + 128 + 256;
__device__ __forceinline__ int feat_off(int h) { return h == 0 ? 0 : (h == 1 ? 64 : (h == 2 ? 160 : 288)); }
__device__ __forceinline__ int logit_off(int l) { return l == 0 ? 0 : (l == 1 ? 2 : (l == 2 ? 5 : (l == 3 ? 9 : 15))); }
} // namespace

// mean over all pixels of image `img` of a strip-layout tensor -> feat[C].  Thread group (chunk, pixel slice) walks its
// pixels in order with 128-bit loads; slices are then added in order.
__device__ __forceinline__ void cu_gap(const __half *act, const ActLayout L, int img, float *partial, float *feat)
{
    const int hp = L.hp(), npl = L.npl(), nch = L.C / 8, npix = npl * hp * hp;
    const int slices = NT / nch; // C in {64, 96, 128, 256} -> nch in {8, 12, 16, 32}: 32 / 21 / 16 / 8 slices
    const int cj = threadIdx.x % nch, sl = threadIdx.x / nch;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sl < slices) {
        for (int p = sl; p < npix; p += slices) {
            const int pl = p / (hp * hp), y = (p / hp) % hp, x = p % hp;
            const size_t off = ((((size_t)pl * nch + cj) * hp + y) * L.strip + img) * hp * 8 + (size_t)x * 8;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(act + off));
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 t = __half22float2(h2[e]);
                s[2 * e] += t.x;
                s[2 * e + 1] += t.y;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; e++) partial[sl * L.C + cj * 8 + e] = s[e];
    }
    __syncthreads();
    const int used = slices < npix ? slices : npix; // slices beyond the pixel count hold zeros
    for (int c = threadIdx.x; c < L.C; c += NT) {
        float t = 0.0f;
        for (int k = 0; k < used; k++) t += partial[k * L.C + c];
        feat[c] = t / (float)(L.H * L.H);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT) cu_head_kernel(const CuHeadParams p)
{
    __shared__ float partial[32 * 64 + 64]; // max over heads of slices * C = 2048 (+ slack for 21 * 96 = 2016)
    __shared__ float feat[FEAT_TOTAL];
    __shared__ float logits[CU_NLOGIT];
    const int n = blockIdx.x;
    griddep_launch_dependents();
    griddep_wait(); // PDL: activations come from the previous kernels of the stream
    for (int h = 0; h < CU_NHEAD; h++) cu_gap(p.act[h], p.lay[h], n, partial, feat + feat_off(h));

    const float poc = (float)p.cus[n].poc, qp = (float)p.cus[n].qp; // raw ints promoted by torch.cat (mlt_cu_or_pq_arch.py:100-101,110)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o = warp; o < CU_NLOGIT; o += NT / 32) {
        int lvl = 0;
        while (o >= logit_off(lvl + 1)) lvl++;
        const int row = o - logit_off(lvl);
        const int C = p.lay[lvl].C;
        const float *f = feat + feat_off(lvl);
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
        mlt_cu_result r;
        for (int l = 0; l < CU_NHEAD; l++) {
            const int b = logit_off(l), cnt = logit_off(l + 1) - b;
            float mx = logits[b];
            int am = 0;
            for (int k = 1; k < cnt; k++)
                if (logits[b + k] > mx) { mx = logits[b + k]; am = k; } // first maximum wins (torch.argmax)
            float e[6], sum = 0.0f;
            for (int k = 0; k < cnt; k++) { e[k] = expf(logits[b + k] - mx); sum += e[k]; }
            for (int k = 0; k < cnt; k++) r.probs[b + k] = e[k] / sum;
            r.split[l] = am;
        }
        for (int k = 0; k < CU_NLOGIT; k++) r.logits[k] = logits[k];
        p.out[n] = r;
    }
}

cudaError_t launch_cu_head(const CuHeadParams &p, cudaStream_t s)
{
    if (p.n <= 0) return cudaSuccess;
    return launch_pdl(cu_head_kernel, dim3(p.n), dim3(NT), 0, s, p);
}

} // namespace mlt
