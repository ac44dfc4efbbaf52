// umma_rate4.cu -- sustained tcgen05.mma rate under full-chip load: do long runs (power-capped steady state) execute more
// SM cycles per MMA than the short bursts of umma_rate3.cu?  Reports clock64 cycles per MMA and the clock implied by
// cycles / wall time.
#include <cstdio>
#include "ptx.cuh"
using namespace mlt;

template <int N, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *out, int iters)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, nobody[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) for (int c = 0; c < 4; c++) mbar_init(&nobody[c], 1);
    for (int i = tid; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 2880), b_lo = umma_desc_lo(smem_u32(smem + 64 * 1024), N * 16);
        constexpr uint32_t a_hi = umma_desc_hi(160), b_hi = umma_desc_hi(128);
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < iters; it++) {
            if (elect_one_sync()) {
#pragma unroll
                for (int i = 0; i < 18; i++)
                    umma_f16(tmem + (it & 3) * N, umma_desc_pack(a_lo + (i % 9) * 11 + (i % 2) * 360, a_hi),
                             umma_desc_pack(b_lo + (i % 9) * 2 * N, b_hi), idesc, i > 0);
                if (COMMIT_EVERY > 0 && (it % (COMMIT_EVERY > 0 ? COMMIT_EVERY : 1)) == COMMIT_EVERY - 1) umma_commit(&bar);
                if (COMMIT_EVERY < 0) {
#pragma unroll
                    for (int c = 0; c < -COMMIT_EVERY; c++) umma_commit(&nobody[c]);
                }
            }
            __syncwarp();
            if (COMMIT_EVERY > 0 && (it % (COMMIT_EVERY > 0 ? COMMIT_EVERY : 1)) == COMMIT_EVERY - 1) { mbar_wait(&bar, phase); phase ^= 1; tc_fence_after(); }
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, phase);
        if (tid == 0) out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int CE>
void run(const char *name, int iters)
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof h);
    const int smem = 128 * 1024;
    cudaFuncSetAttribute(rate_kernel<N, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<N, CE><<<148, 128, smem>>>(d, iters / 10);
    cudaEventRecord(e0);
    rate_kernel<N, CE><<<148, 128, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long worst = 0; for (int b = 0; b < 148; b++) if (h[b] > worst) worst = h[b];
    const double mmas = 18.0 * iters;
    printf("%-44s iters %6d: %6.1f cyc/MMA (clock64, slowest CTA)  %7.3f ms  => %5.0f MHz implied, %6.1f ns/MMA\n", name, iters,
           worst / mmas, ms, worst / (ms * 1e3), ms * 1e6 / mmas);
    cudaFree(d);
}

int main()
{
    run<32, 0>("N=32  no commits, burst", 40);
    run<32, 0>("N=32  no commits, 0.5 ms", 1200);
    run<32, 0>("N=32  no commits, 20 ms", 50000);
    run<32, 1>("N=32  commit+wait every tile, 20 ms", 50000);
    run<32, 4>("N=32  commit+wait every 4 tiles, 20 ms", 50000);
    run<32, -1>("N=32  1 commit per tile, nobody waits", 50000);
    run<32, -2>("N=32  2 commits per tile, nobody waits", 50000);
    run<32, -3>("N=32  3 commits per tile, nobody waits", 50000);
    run<64, -2>("N=64  2 commits per tile (18 MMAs), nobody waits", 40000);
    run<128, -2>("N=128 2 commits per tile (18 MMAs), nobody waits", 30000);
    run<64, 0>("N=64  no commits, 20 ms", 40000);
    run<128, 0>("N=128 no commits, 20 ms", 30000);
    return 0;
}
