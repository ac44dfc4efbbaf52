// umma_rate2.cu -- replicate the production MMA issue loop of conv_umma.cuh (same ConvCfg constants, same
// descriptor math) in isolation: how many cycles does one tile's worth of MMAs take when nothing else runs?
#include <cstdio>
#include "conv_umma.cuh"
using namespace mlt;

template <class C, int NT>
__global__ void __launch_bounds__(128, 1) issue_kernel(long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[NT];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < C::SMEM_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0, 0);
    fence_proxy_async_smem();
    if (tid == 0) { for (int i = 0; i < NT; i++) mbar_init(&bar[i], 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t sA = smem_u32(smem + C::OFF_A), sB = smem_u32(smem + C::OFF_B);
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(128, C::COUT);
        constexpr uint32_t a_hi = umma_desc_hi(C::A_SBO), b_hi = umma_desc_hi(128);
        const uint32_t ones_lo = umma_desc_lo(smem_u32(smem + C::OFF_ONES), 128 * 16);
        const uint32_t bias_lo = umma_desc_lo(smem_u32(smem + C::OFF_BIAS), C::COUT * 16);
        long long t[NT + 1];
        uint32_t a_it = 0;
        t[0] = clock64();
#pragma unroll 1
        for (int tile = 0; tile < NT; tile++) {
            const uint32_t acc = tile % C::NACC;
            const uint32_t d_tmem = tmem_base + acc * C::COUT;
            if (elect_one_sync()) umma_f16(d_tmem, umma_desc_pack(ones_lo, b_hi), umma_desc_pack(bias_lo, b_hi), idesc, 0);
#pragma unroll 1
            for (int cg = 0; cg < C::NCG; cg++, a_it++) {
                const uint32_t st = a_it % C::NAS;
                const uint32_t a_lo0 = umma_desc_lo(sA + st * C::A_STAGE_BYTES, C::A_LBO);
                if (elect_one_sync()) {
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const uint32_t b_lo0 = umma_desc_lo(sB + ((cg * 9 + tap) % 9) * C::SLAB_BYTES, C::COUT * 16);
                        const uint32_t a_tap = a_lo0 + tap_offset_px<C>(tap / 3, tap % 3);
#pragma unroll
                        for (int ks = 0; ks < C::G / 16; ks++)
                            umma_f16(d_tmem, umma_desc_pack(a_tap + ks * (2 * C::A_LBO / 16), a_hi),
                                     umma_desc_pack(b_lo0 + ks * (2 * C::COUT), b_hi), idesc, 1);
                    }
                }
            }
            if (elect_one_sync()) umma_commit(&bar[tile]);
            t[tile + 1] = clock64();
        }
        mbar_wait(&bar[NT - 1], 0);
        tc_fence_after();
        const long long tend = clock64();
        if (lane == 0) {
            for (int i = 0; i <= NT; i++) out[i] = t[i] - t[0];
            out[NT + 1] = tend - t[0];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, C::TMEM_COLS); }
}

template <class C>
void run(const char *name)
{
    constexpr int NT = 8;
    long long *d, h[NT + 2];
    cudaMalloc(&d, sizeof h);
    cudaFuncSetAttribute(issue_kernel<C, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    for (int r = 0; r < 2; r++) {
        issue_kernel<C, NT><<<1, 128, C::SMEM_BYTES>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const int mmas = 1 + C::NCG * 9 * (C::G / 16);
    printf("%-28s %3d MMAs/tile: issue-done stamps per tile:", name, mmas);
    for (int i = 1; i <= NT; i++) printf(" %lld", h[i] - h[i - 1]);
    printf(" | all complete at %lld => %.1f cyc/MMA (floor %d)\n", h[NT + 1], (double)h[NT + 1] / (NT * mmas), 128 * C::COUT / 256);
    cudaFree(d);
}

int main()
{
    run<ConvCfg<32, 32, 1, 64, 0>>("L0c 32->32 s1");
    run<ConvCfg<32, 32, 2, 64, 0>>("L0a 32->32 s2");
    run<ConvCfg<64, 64, 1, 32, 0>>("L1c 64->64 s1");
    run<ConvCfg<32, 64, 2, 32, 0>>("L1a 32->64 s2");
    return 0;
}
