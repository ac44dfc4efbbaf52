// tma_probe.cu -- pins the TMA conventions conv_umma.cuh relies on: a 5-D tiled tensor map over a chunk-planar fp16
// activation tensor [unit][chunk][row][img][x][8ch], SWIZZLE_NONE, box inner extent 80 elements (10 px x 8 ch = 160 B),
// negative / out-of-range start coordinates zero-filled, box written densely to shared memory in dimension order.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mlt;

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int U = 3, CHK = 4, H = 6, NIMG = 2, W = 8;
constexpr int BX = 10, BI = 2, BR = 5, BC = 2; // box: 10 px, 2 imgs, 5 rows, 2 chunks

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, __half *out, int c0, int c1, int c2, int c3, int c4)
{
    __shared__ __align__(128) __half buf[BC * BR * BI * BX * 8];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < BC * BR * BI * BX * 8; i += blockDim.x) buf[i] = __float2half(-1.0f);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, sizeof(buf));
        tma_load_5d(smem_u32(buf), &tm, c0, c1, c2, c3, c4, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BC * BR * BI * BX * 8; i += blockDim.x) out[i] = buf[i];
}

int main()
{
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    const size_t n = (size_t)U * CHK * H * NIMG * W * 8;
    std::vector<__half> h(n);
    for (size_t i = 0; i < n; i++) h[i] = __float2half((float)(i % 2039) + 1.0f); // exact in fp16, never 0 or -1
    __half *d, *dout;
    cudaMalloc(&d, n * 2);
    cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
    const int nb = BC * BR * BI * BX * 8;
    cudaMalloc(&dout, nb * 2);
    CUtensorMap tm;
    cuuint64_t dims[5] = {W * 8, NIMG, H, CHK, U};
    cuuint64_t strides[4] = {W * 16, (cuuint64_t)NIMG * W * 16, (cuuint64_t)H * NIMG * W * 16, (cuuint64_t)CHK * H * NIMG * W * 16};
    cuuint32_t box[5] = {BX * 8, BI, BR, BC, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    int bad_total = 0;
    const int cases[4][5] = {{-8, 0, -1, 1, 2}, {0, 0, 0, 0, 0}, {-8, 0, 3, 2, 1}, {8, 0, 2, 2, 2}};
    for (int t = 0; t < 4; t++) {
        const int *c = cases[t];
        probe_kernel<<<1, 128>>>(tm, dout, c[0], c[1], c[2], c[3], c[4]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<__half> o(nb);
        cudaMemcpy(o.data(), dout, nb * 2, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int k = 0; k < BC; k++) for (int rr = 0; rr < BR; rr++) for (int im = 0; im < BI; im++) for (int x = 0; x < BX; x++) for (int e8 = 0; e8 < 8; e8++) {
            const int gx8 = c[0] + x * 8 + e8, gi = c[1] + im, gr = c[2] + rr, gk = c[3] + k, gu = c[4];
            float want = 0.0f;
            if (gx8 >= 0 && gx8 < W * 8 && gi >= 0 && gi < NIMG && gr >= 0 && gr < H && gk >= 0 && gk < CHK && gu >= 0 && gu < U) {
                const size_t idx = ((((size_t)gu * CHK + gk) * H + gr) * NIMG + gi) * W * 8 + gx8;
                want = __half2float(h[idx]);
            }
            const float got = __half2float(o[(((k * BR + rr) * BI + im) * BX + x) * 8 + e8]);
            if (got != want) { if (bad < 4) printf("  case %d mismatch k%d r%d i%d x%d e%d got %g want %g\n", t, k, rr, im, x, e8, got, want); bad++; }
        }
        printf("case %d start (%d,%d,%d,%d,%d): %s (%d bad)\n", t, c[0], c[1], c[2], c[3], c[4], bad ? "FAIL" : "OK", bad);
        bad_total += bad;
    }
    printf(bad_total ? "tma_probe: FAILED\n" : "tma_probe: ALL OK\n");
    return bad_total != 0;
}
