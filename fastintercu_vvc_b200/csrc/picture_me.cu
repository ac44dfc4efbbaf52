// picture_me.cu -- integer full-search block matching for the frame-level pre-pass: one
// integer MV per eligible CTU, so that the neighbour-independent prediction of mlt_predict_picture follows the picture's
// motion instead of being the co-located block.  The reference has no such stage (its `pred` is the RDO loop's own best
// merge candidate, EncCu.cpp:820-830); this is the cheapest stand-in that needs nothing from neighbouring CTUs:
//   cost(dx, dy) = sum over the 128x128 block of |org(x, y) - ref(clamp(x + dx), clamp(y + dy))|,  dx, dy in [-R, R]
// on the reference plane with replicated borders (Picture.cpp:1117); the winner is the candidate with the smallest
// (cost, preference) pair, preference 0 = the zero MV, then raster order of (dy, dx) -- fully deterministic.
//
// Integer, shared-memory-bound work (D^2 candidates x 16,384 samples per CTU).  CTA = (CTU, band of 16 rows): the band of
// the original block and the matching (16 + 2R) x (128 + 2R) reference window are staged in shared memory once, thread =
// candidate (consecutive lanes read consecutive reference samples, the original sample is a broadcast), partial costs
// are added with integer atomics (order-independent), and a second small kernel takes the arg-min per CTU.
#include "mlt_internal.h"

namespace mlt {
namespace {

constexpr int CTU = MLT_CTU_SIZE;
constexpr int BAND = 16; // rows per CTA

__global__ void __launch_bounds__(1024) picture_me_cost_kernel(const int16_t *__restrict__ org, const int16_t *__restrict__ ref, int pitch, int w, int h,
                                                              const PicCtu *__restrict__ ctus, int R, unsigned *__restrict__ cost /*[n][D*D]*/)
{
    extern __shared__ int16_t sm[];
    const int D = 2 * R + 1, P = CTU + 2 * R + 2, WR = BAND + 2 * R; // window pitch (bank-staggered) and rows
    int16_t *so = sm, *sr = sm + BAND * CTU;
    const PicCtu c = ctus[blockIdx.y];
    const int y0 = c.y + blockIdx.x * BAND;
    for (int i = threadIdx.x; i < BAND * CTU / 8; i += blockDim.x) { // 128-bit loads: CTU columns are 16-byte aligned
        const int r = i / (CTU / 8), v = i % (CTU / 8);
        reinterpret_cast<uint4 *>(so)[i] = __ldg(reinterpret_cast<const uint4 *>(org + (size_t)(y0 + r) * pitch + c.x) + v);
    }
    for (int i = threadIdx.x; i < WR * (CTU + 2 * R); i += blockDim.x) {
        const int r = i / (CTU + 2 * R), x = i % (CTU + 2 * R);
        int sy = y0 - R + r, sx = c.x - R + x;
        sy = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy);
        sx = sx < 0 ? 0 : (sx > w - 1 ? w - 1 : sx);
        sr[r * P + x] = __ldg(ref + (size_t)sy * pitch + sx);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < D * D; k += blockDim.x) {
        const int dy = k / D, dx = k % D; // offsets into the window (MV + R)
        unsigned acc = 0;
        for (int r = 0; r < BAND; r++) {
            const int16_t *o = so + r * CTU, *q = sr + (r + dy) * P + dx;
#pragma unroll 8
            for (int x = 0; x < CTU; x++) acc = __sad((int)o[x], (int)q[x], acc);
        }
        atomicAdd(cost + (size_t)blockIdx.y * D * D + k, acc);
    }
}

// one warp per CTU: arg-min of (cost, preference); preference 0 = zero MV, then raster order of (dy, dx)
__global__ void picture_me_argmin_kernel(const unsigned *__restrict__ cost, int n, int R, int16_t *__restrict__ mv /*[n][2]*/, unsigned *__restrict__ best_cost)
{
    const int ctu = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ctu >= n) return;
    const int D = 2 * R + 1, centre = R * D + R;
    unsigned long long best = ~0ull;
    for (int k = lane; k < D * D; k += 32) {
        const unsigned pref = k == centre ? 0u : (k < centre ? (unsigned)k + 1u : (unsigned)k);
        const unsigned long long key = ((unsigned long long)cost[(size_t)ctu * D * D + k] << 16) | pref;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) {
        const unsigned pref = (unsigned)(best & 0xffffu);
        const int k = pref == 0 ? centre : ((int)pref <= centre ? (int)pref - 1 : (int)pref);
        mv[2 * ctu] = (int16_t)(k % D - R);
        mv[2 * ctu + 1] = (int16_t)(k / D - R);
        if (best_cost) best_cost[ctu] = (unsigned)(best >> 16);
    }
}

} // namespace

cudaError_t launch_picture_me(const int16_t *org, const int16_t *ref, int pitch, int w, int h, const PicCtu *ctus, int n, int R, unsigned *cost,
                              int16_t *mv, unsigned *best_cost, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const int D = 2 * R + 1;
    cudaError_t e = cudaMemsetAsync(cost, 0, (size_t)n * D * D * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    // one thread per candidate when they fit one block (R = 8: 289 candidates -> 320 threads), else strided over 1024
    const int threads = (D * D + 31) / 32 * 32 < 1024 ? (D * D + 31) / 32 * 32 : 1024;
    picture_me_cost_kernel<<<dim3(CTU / BAND, n), threads < 128 ? 128 : threads, picture_me_smem_bytes(R), s>>>(org, ref, pitch, w, h, ctus, R, cost);
    picture_me_argmin_kernel<<<(n + 7) / 8, 256, 0, s>>>(cost, n, R, mv, best_cost);
    return cudaGetLastError();
}

} // namespace mlt
