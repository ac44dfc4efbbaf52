// cu_head.cu -- prediction head of the smaller-CU networks: four global average pools (after layer1..layer4,
// mlt_cu_or_pq_arch.py:108-127), cat(poc, qp) + Linear (66->2, 98->3, 130->4, 258->6), per-level softmax and argmax
// (the hook uses level 1 below 128x128: EncCu.cpp:916-921).  One warp per CU; every reduction runs in a fixed order, so
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

// Eight CUs per block.  Pooling: maps >= 8x8 come as fp32 partial sums left by the conv epilogues (conv_umma.cuh GAP; one
// warp per CU adds them, lanes across channels); maps <= 4x4 are pooled from the stored fp16 strip-layout activation
// (+ its lo tensor where the layer writes hi + lo pairs) with thread = (CU, 8-channel chunk), CU fastest, so that a warp
// reads runs of consecutive CUs (the strip layout keeps the same pixel of consecutive CUs hp * 16 B apart).  Then one
// warp per CU for the FC heads.  A block per CU (first version) or a warp per CU with lanes across chunks (second) were
// all launch overhead / uncoalesced 16-byte reads: 1.07 and 0.43 ms of a 61,440-CU step.
__global__ void __launch_bounds__(NT) cu_head_kernel(const CuHeadParams p)
{
    constexpr int CPB = NT / 32; // CUs per block
    __shared__ float s_feat[CPB][FEAT_TOTAL];
    __shared__ float s_logits[CPB][CU_NLOGIT + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * CPB;
    griddep_launch_dependents();
    griddep_wait(); // PDL: activations come from the previous kernels of the stream
    for (int h = 0; h < CU_NHEAD; h++) {
        const ActLayout L = p.lay[h];
        const float inv = 1.0f / (float)(L.H * L.H);
        if (p.gap_part[h] != nullptr) {
            // add the per-(tile, lane quadrant) fp32 partial sums of CU n0 + warp in a fixed order
            const int n = n0 + warp, cnt = p.gap_count[h];
            if (n < p.n) {
                const float *g = p.gap_part[h] + (size_t)n * cnt * L.C;
                for (int c = lane; c < L.C; c += 32) {
                    float t = 0.0f;
                    for (int k = 0; k < cnt; k++) t += g[k * L.C + c];
                    s_feat[warp][feat_off(h) + c] = t * inv;
                }
            }
        } else {
            const int hp = L.hp(), npix = L.npl() * hp * hp, nch = L.C / 8;
            const int cu = threadIdx.x % CPB, n = n0 + cu;
            for (int cj = threadIdx.x / CPB; cj < nch; cj += NT / CPB) {
                float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (n < p.n) {
                    for (int q = 0; q < npix; q++) { // pixels in a fixed order
                        const int pl = q / (hp * hp), y = (q / hp) % hp, x = q % hp;
                        const size_t off = ((((size_t)pl * nch + cj) * hp + y) * L.strip + n) * hp * 8 + (size_t)x * 8;
                        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.act[h] + off));
                        const __half2 *h2 = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float2 t = __half22float2(h2[e]);
                            s[2 * e] += t.x;
                            s[2 * e + 1] += t.y;
                        }
                        if (p.hilo[h]) { // fp16 hi + lo pair: the lo tensor sits one whole tensor further on
                            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(p.act[h] + L.unit_elems() + off));
                            const __half2 *l2 = reinterpret_cast<const __half2 *>(&w);
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float2 t = __half22float2(l2[e]);
                                s[2 * e] += t.x;
                                s[2 * e + 1] += t.y;
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; e++) s_feat[cu][feat_off(h) + cj * 8 + e] = s[e] * inv;
            }
        }
    }
    __syncthreads();
    const int n = n0 + warp;
    if (n >= p.n) return; // (no block-level barrier below: warps are independent)
    float *feat = s_feat[warp], *logits = s_logits[warp];
    __syncwarp();
    const float poc = (float)p.cus[n].poc, qp = (float)p.cus[n].qp; // raw ints promoted by torch.cat (mlt_cu_or_pq_arch.py:100-101,110)
    for (int o = 0; o < CU_NLOGIT; o++) {
        int lvl = 0;
        while (o >= logit_off(lvl + 1)) lvl++;
        const int row = o - logit_off(lvl);
        const int C = p.lay[lvl].C;
        const float *f = feat + feat_off(lvl);
        const float *wr = p.fc_w[lvl] + (size_t)row * (C + 2);
        float s = 0.0f;
        for (int k = lane; k < C; k += 32) s = fmaf(__ldg(wr + k), f[k], s);
        if (lane == 0) s = fmaf(__ldg(wr + C), poc, s);
        if (lane == 1) s = fmaf(__ldg(wr + C + 1), qp, s);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) logits[o] = s + __ldg(p.fc_b[lvl] + row);
    }
    __syncwarp();
    if (lane == 0) {
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
    return launch_pdl(cu_head_kernel, dim3((p.n + NT / 32 - 1) / (NT / 32)), dim3(NT), 0, s, p);
}

} // namespace mlt
