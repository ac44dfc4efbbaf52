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

// One WARP per CU (8 CUs per block): the pooled inputs are tiny -- fp32 partial sums left by the conv epilogues for maps
// >= 8x8 (conv_umma.cuh GAP), or the stored fp16 activation of a map <= 4x4 -- so a block per CU would be all launch and
// scheduling overhead (measured: 1.07 ms of 3.9 ms for 61,440 16-px CUs).
__global__ void __launch_bounds__(NT) cu_head_kernel(const CuHeadParams p)
{
    __shared__ float s_feat[NT / 32][FEAT_TOTAL];
    __shared__ float s_logits[NT / 32][CU_NLOGIT + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * (NT / 32) + warp;
    griddep_launch_dependents();
    griddep_wait(); // PDL: activations come from the previous kernels of the stream
    if (n >= p.n) return; // (no block-level barrier below: warps are independent)
    float *feat = s_feat[warp], *logits = s_logits[warp];
    for (int h = 0; h < CU_NHEAD; h++) {
        const ActLayout L = p.lay[h];
        const float inv = 1.0f / (float)(L.H * L.H);
        float *f = feat + feat_off(h);
        if (p.gap_part[h] != nullptr) {
            // maps >= 8x8: add the per-(tile, lane quadrant) fp32 partial sums in a fixed order
            const int cnt = p.gap_count[h];
            const float *g = p.gap_part[h] + (size_t)n * cnt * L.C;
            for (int c = lane; c < L.C; c += 32) {
                float t = 0.0f;
                for (int k = 0; k < cnt; k++) t += g[k * L.C + c];
                f[c] = t * inv;
            }
        } else {
            // maps <= 4x4: pool the stored strip-layout activation; lane = 8-channel chunk, pixels in order
            const int hp = L.hp(), npix = L.npl() * hp * hp, nch = L.C / 8;
            for (int cj = lane; cj < nch; cj += 32) {
                float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int q = 0; q < npix; q++) {
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
#pragma unroll
                for (int e = 0; e < 8; e++) f[cj * 8 + e] = s[e] * inv;
            }
        }
    }
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
