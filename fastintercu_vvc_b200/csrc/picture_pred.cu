// picture_pred.cu -- frame-level pre-pass: a neighbour-independent prediction block for
// EVERY eligible CTU of a picture, built on the device from one reference-picture luma plane, so that the whole frame is
// inferred in one batch before EncSlice::encodeCtus (EncSlice.cpp:1479) walks the CTUs.
//
// The reference's hook takes `pred` from the RDO loop itself (best of AFFINE + MERGE_SKIP so far, EncCu.cpp:820-830),
// which ties every inference call to the CTU before it.  Here pred(x, y) = ref(clamp(x + mvx), clamp(y + mvy)): integer-
// sample motion compensation (no interpolation filter) from a reference picture whose border is extended by sample
// replication, which is what Picture::extendPicBorder (Picture.cpp:1117) does to every reference picture.  One integer
// MV per CTU (zero MV when none is given).
//
// HBM-bound byte work: 32 KiB read + 32 KiB written per CTU.  One thread = 8 samples of one row = one 128-bit store;
// interior threads read two aligned 128-bit words and funnel-shift by the MV's sub-word offset, threads whose window
// touches the picture border fall back to clamped scalar loads.
#include "mlt_internal.h"

namespace mlt {
namespace {

constexpr int CTU = MLT_CTU_SIZE;
constexpr int ROWS_PER_CTA = 16; // 16 threads per row x 16 rows = 256 threads

// 8 samples of row `row` starting at sample sx of a plane with replicated borders -- interior windows: two aligned 128-bit words, funnel-shifted;
// windows touching the border: clamped sample by sample
__device__ __forceinline__ uint4 load8_clamped(const int16_t *__restrict__ row, int sx, int w)
{
    uint32_t o[4];
    if (sx >= 0 && sx + 8 <= w) {
        const int sh = sx & 7;
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(row + (sx - sh)));
        if (sh == 0) return lo;
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(row + (sx - sh) + 8));
        const uint32_t wd[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const int ws = sh >> 1;
        uint32_t v[5];
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] = ws == 0 ? wd[j] : (ws == 1 ? wd[j + 1] : (ws == 2 ? wd[j + 2] : wd[j + 3]));
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = (sh & 1) ? __funnelshift_r(v[j], v[j + 1], 16) : v[j];
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int x0 = sx + 2 * j, x1 = x0 + 1;
            x0 = x0 < 0 ? 0 : (x0 > w - 1 ? w - 1 : x0);
            x1 = x1 < 0 ? 0 : (x1 > w - 1 ? w - 1 : x1);
            o[j] = (uint32_t)(uint16_t)__ldg(row + x0) | ((uint32_t)(uint16_t)__ldg(row + x1) << 16);
        }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) picture_pred_kernel(const int16_t *__restrict__ ref, int pitch, int w, int h, const PicCtu *__restrict__ ctus,
                                                           int16_t *__restrict__ out /*[n][2][128][128], plane 1 written*/)
{
    const PicCtu c = ctus[blockIdx.y];
    const int t = threadIdx.x & 15, r = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 4);
    int sy = c.y + r + c.mvy;
    sy = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy);
    const int sx = c.x + 8 * t + c.mvx;
    const int16_t *row = ref + (size_t)sy * pitch;
    int16_t *dst = out + ((size_t)blockIdx.y * 2 + 1) * CTU * CTU + (size_t)r * CTU + 8 * t;
    // interior: the 8 samples lie inside [a, a + 16) with a = sx rounded down to 8 samples (rows are 16-byte aligned and the
    // pitch is a multiple of 8, so the second word exists whenever the shift is non-zero)
    *reinterpret_cast<uint4 *>(dst) = load8_clamped(row, sx, w);
}

// The smaller-CU form: every S x S block of the picture's CU raster (cols blocks per row, fully inside the picture) as a
// dense [n][2][S][S] batch -- plane 0 = the original block, plane 1 = the integer-MV prediction out of the reference plane
// -- plus its (poc, qp) pair.  One thread = 8 samples of one row of both planes.
__global__ void __launch_bounds__(256) picture_cu_gather_kernel(const int16_t *__restrict__ org, const int16_t *__restrict__ ref, int pitch, int w, int h,
                                                                int S, int cols, int n, const int16_t *__restrict__ mv /*[n][2] or null*/, int poc, int qp,
                                                                int16_t *__restrict__ out, int32_t *__restrict__ pocqp)
{
    const int tpr = S >> 3, tpc = S * tpr; // threads per row / per CU
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * tpc) return;
    const int cu = (int)(g / tpc), k = (int)(g % tpc), r = k / tpr, t = k % tpr;
    const int x = (cu % cols) * S + 8 * t, y = (cu / cols) * S + r;
    const int mvx = mv ? mv[2 * cu] : 0, mvy = mv ? mv[2 * cu + 1] : 0;
    int sy = y + mvy;
    sy = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy);
    int16_t *dst = out + (size_t)cu * 2 * S * S + (size_t)r * S + 8 * t;
    *reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(org + (size_t)y * pitch + x));
    *reinterpret_cast<uint4 *>(dst + S * S) = load8_clamped(ref + (size_t)sy * pitch, x + mvx, w);
    if (k == 0) { pocqp[2 * cu] = poc; pocqp[2 * cu + 1] = qp; }
}

} // namespace

cudaError_t launch_picture_cu_gather(const int16_t *org, const int16_t *ref, int pitch, int w, int h, int size, int n, const int16_t *mv, int poc,
                                     int qp, int16_t *out, int32_t *pocqp, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const long long threads = (long long)n * size * (size / 8);
    picture_cu_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(org, ref, pitch, w, h, size, w / size, n, mv, poc, qp, out, pocqp);
    return cudaGetLastError();
}

cudaError_t launch_picture_pred(const int16_t *ref, int pitch, int w, int h, const PicCtu *ctus, int n, int16_t *out, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    picture_pred_kernel<<<dim3(CTU / ROWS_PER_CTA, n), 256, 0, s>>>(ref, pitch, w, h, ctus, out);
    return cudaGetLastError();
}

} // namespace mlt
