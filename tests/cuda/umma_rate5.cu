// umma_rate5.cu -- which property of the production MMA stream costs ~14 cycles per N=32 MMA relative to umma_rate4?
// Variants: operand data (ones / random bits / small random fp16), A stage rotation, TMEM allocation size.
#include <cstdio>
#include "ptx.cuh"
using namespace mlt;

// DATA 0: ones+zeros pattern; 1: random bits (NaN/Inf/denormals included); 2: finite small fp16
// ROT  1: rotate the A base over 8 stages of 11520 B (production L0c ring), B at 92160 + tap * 2048
template <int DATA, int ROT, int TCOLS, int FENCE>
__global__ void __launch_bounds__(352, 1) rate_kernel(long long *out, int iters)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rng = 1234567u + tid * 7919u + blockIdx.x * 104729u;
    for (int i = tid; i < 116 * 1024 / 4; i += blockDim.x) {
        rng = rng * 1664525u + 1013904223u;
        uint32_t v = 0x3C003C00u;
        if (DATA == 1) v = rng;
        if (DATA == 2) v = (rng & 0x83FF83FFu) | 0x30003000u; // +-[0.125, 0.25)
        reinterpret_cast<uint32_t *>(smem)[i] = v;
    }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, TCOLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 8) {
        constexpr uint32_t idesc = umma_idesc_f16(128, 32);
        constexpr uint32_t a_hi = umma_desc_hi(160), b_hi = umma_desc_hi(128);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + (ROT ? 92160 : 64 * 1024));
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (FENCE & 1) tc_fence_after();
            if (FENCE & 2) tc_fence_before();
            const uint32_t a_lo = umma_desc_lo(sA + (ROT ? (it & 7) * 11520 : 0), 2880);
            if (elect_one_sync()) {
#pragma unroll
                for (int tap = 0; tap < 9; tap++) {
                    const uint32_t b_lo = umma_desc_lo(sB + tap * 2048, 32 * 16);
                    const uint32_t a_tap = a_lo + (tap / 3) * 10 + tap % 3;
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
                        umma_f16(tmem + (it & 3) * 32, umma_desc_pack(a_tap + ks * 360, a_hi), umma_desc_pack(b_lo + ks * 64, b_hi), idesc, (tap | ks) != 0);
                }
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if ((tid & 31) == 0) out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, TCOLS); }
}

template <int DATA, int ROT, int TCOLS, int FENCE = 0>
void run(const char *name, int iters)
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof h);
    const int smem = 118 * 1024;
    cudaFuncSetAttribute(rate_kernel<DATA, ROT, TCOLS, FENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<DATA, ROT, TCOLS, FENCE><<<148, 352, smem>>>(d, iters / 10);
    cudaEventRecord(e0);
    rate_kernel<DATA, ROT, TCOLS, FENCE><<<148, 352, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long worst = 0; for (int b = 0; b < 148; b++) if (h[b] > worst) worst = h[b];
    const double mmas = 18.0 * iters;
    printf("%-52s %6.1f cyc/MMA  %7.3f ms  => %5.0f MHz implied, %6.1f ns/MMA\n", name, worst / mmas, ms, worst / (ms * 1e3), ms * 1e6 / mmas);
    cudaFree(d);
}

int main()
{
    run<0, 1, 128>("baseline (rotating A ring, TMEM 128)", 30000);
    run<0, 1, 128, 1>("+ tcgen05.fence::after_thread_sync per tile", 30000);
    run<0, 1, 128, 2>("+ tcgen05.fence::before_thread_sync per tile", 30000);
    run<0, 1, 128, 3>("+ both fences per tile", 30000);
    return 0;
}
