// umma_rate.cu -- tcgen05.mma issue / completion rate microbenchmark (M=128, K=16, kind::f16), fully unrolled with
// compile-time descriptor offsets like the production issue loop.  Variants add the production kernel's
// concurrent activity one by one: full grid, cp.async writers into shared memory, warps spinning on an mbarrier.
#include <cstdio>
#include "ptx.cuh"
using namespace mlt;

// MODE bit0: 4 warps stream cp.async into smem; bit1: 8 warps spin on an mbarrier; bit2: 4 warps do tcgen05.ld
template <int N, int NACC, int NMMA, int MODE>
__global__ void __launch_bounds__(128 + 512, 1) rate_kernel(long long *out, const uint4 *gsrc)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, never;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); mbar_fence_init(); done = 0; }
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 2880), b_lo = umma_desc_lo(smem_u32(smem + 48 * 1024), N * 16);
        constexpr uint32_t a_hi = umma_desc_hi(160), b_hi = umma_desc_hi(128);
        t0 = clock64();
        if (elect_one_sync()) {
#pragma unroll
            for (int i = 0; i < NMMA; i++)
                umma_f16(tmem + (i % NACC) * N, umma_desc_pack(a_lo + (i % 9) * 11 + (i % 2) * 360, a_hi),
                         umma_desc_pack(b_lo + (i % 9) * 2 * N, b_hi), idesc, i >= NACC);
            umma_commit(&bar);
        }
        t1 = clock64();
        mbar_wait(&bar, 0);
        tc_fence_after();
        if (tid == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = clock64() - t0; done = 1; mbar_arrive(&never); }
    } else if (warp < 4) {
        // idle
    } else if (warp < 8) {
        if (MODE & 1) { // cp.async writers into the upper smem region (not read by the MMAs)
            const uint32_t dst = smem_u32(smem + 96 * 1024) + (tid - 128) * 16;
            int it = 0;
            while (!done) {
#pragma unroll
                for (int k = 0; k < 6; k++) cp_async16(dst + k * 2048, gsrc + (size_t)blockIdx.x * 4096 + ((it * 6 + k) % 32) * 128 + (tid - 128), true);
                cp_async_commit();
                cp_async_wait<2>();
                it++;
            }
            cp_async_wait_all();
        }
    } else if (warp < 16) {
        if (MODE & 2) mbar_wait(&never, 0); // spinning waiters, like epilogue / producer warps blocked on a barrier
    } else {
        if (MODE & 4) {
            uint32_t v[32];
            uint32_t acc = 0;
            while (!done) {
                tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256, v);
                tmem_ld_wait();
                acc += v[0];
            }
            if (acc == 0x12345678u) out[0] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, int NMMA, int MODE>
void run(const char *name, int grid, const uint4 *gsrc)
{
    long long *d, h[2 * 148], best[2] = {1LL << 60, 1LL << 60};
    cudaMalloc(&d, sizeof h);
    const int smem = 96 * 1024 + 16 * 1024;
    cudaFuncSetAttribute(rate_kernel<N, NACC, NMMA, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int r = 0; r < 5; r++) {
        rate_kernel<N, NACC, NMMA, MODE><<<grid, 640, smem>>>(d, gsrc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        long long worst[2] = {0, 0};
        for (int b = 0; b < grid; b++) for (int k = 0; k < 2; k++) if (h[2 * b + k] > worst[k]) worst[k] = h[2 * b + k];
        for (int k = 0; k < 2; k++) if (worst[k] < best[k]) best[k] = worst[k];
    }
    printf("%-52s issue %6.1f  complete %6.1f cyc/MMA (slowest CTA; math floor %d)\n", name, (double)best[0] / NMMA, (double)best[1] / NMMA, 128 * N / 256);
    cudaFree(d);
}

int main()
{
    uint4 *gsrc;
    cudaMalloc(&gsrc, 148 * 4096 * 16);
    cudaMemset(gsrc, 0, 148 * 4096 * 16);
    run<32, 1, 72, 0>("N=32 1 CTA   alone", 1, gsrc);
    run<32, 1, 72, 0>("N=32 148 CTAs alone", 148, gsrc);
    run<32, 1, 72, 1>("N=32 148 CTAs + cp.async writers", 148, gsrc);
    run<32, 1, 72, 2>("N=32 148 CTAs + 8 warps spinning on mbarrier", 148, gsrc);
    run<32, 1, 72, 4>("N=32 148 CTAs + 4 warps tcgen05.ld", 148, gsrc);
    run<32, 1, 72, 7>("N=32 148 CTAs + all three", 148, gsrc);
    run<32, 4, 72, 7>("N=32 148 CTAs + all three, 4 accumulators", 148, gsrc);
    run<64, 1, 72, 0>("N=64 148 CTAs alone", 148, gsrc);
    run<64, 1, 72, 7>("N=64 148 CTAs + all three", 148, gsrc);
    run<128, 1, 36, 0>("N=128 148 CTAs alone", 148, gsrc);
    run<128, 1, 36, 7>("N=128 148 CTAs + all three", 148, gsrc);
    run<256, 1, 18, 0>("N=256 148 CTAs alone", 148, gsrc);
    run<256, 1, 18, 7>("N=256 148 CTAs + all three", 148, gsrc);
    return 0;
}
