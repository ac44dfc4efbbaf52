// umma_probe.cu -- hardware probe for the tcgen05 shared-memory operand descriptor conventions that
// csrc/conv_umma.cuh relies on (K-major, SWIZZLE_NONE, arbitrary 16-byte-aligned start addresses,
// SBO = patch-row pitch that is NOT a multiple of 128 B).  Run on the B200 box:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I fastintercu_vvc_b200/csrc tests/cuda/umma_probe.cu -o tests/cuda/umma_probe.bin
// Each case fills shared memory with a known pattern, issues tcgen05.mma (M=128, N, K=16*nk) and compares
// the TMEM result with the CPU expectation under hypothesis H1 (LBO = K-chunk step, SBO = 8-row-group step,
// the convention of the kernels) and H2 (the two swapped).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ptx.cuh"

using namespace mlt;

constexpr int SMEM_A = 64 * 1024, SMEM_B = 64 * 1024;

struct Case {
    const char *name;
    uint32_t a_start, a_lbo, a_sbo, a_kstep; // byte offsets inside the A region
    uint32_t b_start, b_lbo, b_sbo, b_kstep;
    int n, nk;
    int swz = 0;      // 0 = SWIZZLE_NONE, 2 = SWIZZLE_128B (descriptor layout_type), applies to A and B
    int a_boff = 0;   // descriptor base_offset field for A
    int reps = 1;     // timing: issue the nk MMAs `reps` times
    int nacc = 1;     // timing: round-robin over this many independent TMEM accumulators (nacc * n <= 256 columns)
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int swz, int boff)
{
    return umma_desc_kmajor_noswz(addr, lbo, sbo) | ((uint64_t)(boff & 7) << 49) | ((uint64_t)(swz & 7) << 61);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __half *a_img, const __half *b_img, Case cs, float *D, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < SMEM_A / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(a_img)[i];
    for (int i = tid; i < SMEM_B / 16; i += 128) reinterpret_cast<uint4 *>(smem + SMEM_A)[i] = reinterpret_cast<const uint4 *>(b_img)[i];
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_f16(128, cs.n);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + SMEM_A);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int r = 0; r < cs.reps; r++)
                for (int k = 0; k < cs.nk; k++) {
                    const uint64_t ad = make_desc(sa + cs.a_start + k * cs.a_kstep, cs.a_lbo, cs.a_sbo, cs.swz, cs.a_boff);
                    const uint64_t bd = make_desc(sb + cs.b_start + k * cs.b_kstep, cs.b_lbo, cs.b_sbo, cs.swz, 0);
                    const int i = r * cs.nk + k;
                    umma_f16(tmem + (i % cs.nacc) * cs.n, ad, bd, idesc, i >= cs.nacc);
                }
            umma_commit(&bar);
        }
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (tid == 0) *cycles = clock64() - t0;
    for (int c0 = 0; c0 < cs.n; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; i++) D[(size_t)tid * cs.n + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

static float h2f(__half h) { return __half2float(h); }

// byte address of logical element (row r, k index kk of K-step ks) under a (lbo, sbo) interpretation.
// swz == 2: 128-byte rows, the 16-byte chunk index is XORed with ADDRESS bits [7:9] (how TMA / our cp.async
// producers would write a pixel-major [pixel][64 ch] tile), i.e. "absolute-address swizzle".
static size_t elem_addr(uint32_t start, uint32_t lbo, uint32_t sbo, uint32_t kstep, int r, int ks, int kk, int swz)
{
    if (swz == 0)
        return (size_t)start + (size_t)ks * kstep + (size_t)(r / 8) * sbo + (size_t)(r % 8) * 16 + (size_t)(kk / 8) * lbo + (size_t)(kk % 8) * 2;
    const size_t lin = (size_t)start + (size_t)ks * kstep + (size_t)(r / 8) * sbo + (size_t)(r % 8) * 128 + (size_t)kk * 2; // un-swizzled address
    const size_t chunk = (lin >> 4) & 7, row = (lin >> 7) & 7;
    return (lin & ~(size_t)0x70) | ((chunk ^ row) << 4);
}
static float elem(const std::vector<__half> &img, uint32_t start, uint32_t lbo, uint32_t sbo, uint32_t kstep, int r, int ks, int kk, int swz)
{
    const size_t byte = elem_addr(start, lbo, sbo, kstep, r, ks, kk, swz);
    if (byte / 2 >= img.size()) return 0.f;
    return h2f(img[byte / 2]);
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 2; }
    printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    std::vector<__half> a(SMEM_A / 2), b(SMEM_B / 2);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 16) % 9) - 4; };
    for (auto &x : a) x = __float2half((float)rnd());
    for (auto &x : b) x = __float2half((float)rnd());
    __half *da, *db;
    float *dD;
    long long *dcyc;
    cudaMalloc(&da, SMEM_A); cudaMalloc(&db, SMEM_B); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dcyc, 8);
    cudaMemcpy(da, a.data(), SMEM_A, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), SMEM_B, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_A + SMEM_B);

    const Case cases[] = {
        // name                         a_start a_lbo  a_sbo a_kstep   b_start b_lbo b_sbo b_kstep  n   nk swz boff
        {"dense canonical N=32 K=16",        0, 2048,  128,  4096,       0,  512,  128, 1024,  32, 1},
        {"dense canonical N=32 K=64",        0, 2048,  128,  4096,       0,  512,  128, 1024,  32, 4},
        {"dense canonical N=256 K=32",       0, 2048,  128,  4096,       0, 4096,  128, 8192, 256, 2},
        {"A sbo=160 (s1 patch), start 0",    0, 2880,  160,  5760,       0,  512,  128, 1024,  32, 2},
        {"A sbo=160, start +16B",           16, 2880,  160,  5760,       0,  512,  128, 1024,  32, 2},
        {"A sbo=160, start +176B (tap 1,1)", 176, 2880,  160,  5760,       0,  512,  128, 1024,  32, 2},
        {"A sbo=160, start +352B (tap 2,2)", 352, 2880,  160,  5760,       0,  512,  128, 1024,  32, 2},
        {"A sbo=144 (s2 patch) lbo=9792",   2448 + 160, 9792, 144, 19584, 0, 1024,  128, 2048,  64, 2},
        {"A sbo=160 lbo=3200 (2-img patch)", 16 + 320, 3200,  160,  6400,     0, 4096,  128, 8192, 256, 2},
        {"B start +16B (unaligned slab)",    0, 2048,  128,  4096,      16,  512,  128, 1024,  32, 2},
        // ---- SWIZZLE_128B, K-major: rows of 128 B (64 fp16), 8-row groups SBO apart, K-step = +32 B
        {"SW128 dense N=32 K=64",            0,   16, 1024,    32,       0,   16, 1024,   32,  32, 4, 2, 0},
        {"SW128 dense N=256 K=64",           0,   16, 1024,    32,       0,   16, 1024,   32, 256, 4, 2, 0},
        {"SW128 A sbo=1280 start 0",         0,   16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 0},
        {"SW128 A sbo=1280 start +128 b0",  128,   16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 0},
        {"SW128 A sbo=1280 start +128 b1",  128,   16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 1},
        {"SW128 A sbo=1280 start +1408 b0", 1408,  16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 0},
        {"SW128 A sbo=1280 start +1408 b3", 1408,  16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 3},
        {"SW128 A sbo=1280 start +2816 b0", 2816,  16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 0},
        {"SW128 A sbo=1280 start +2816 b6", 2816,  16, 1280,    32,       0,   16, 1024,   32,  32, 4, 2, 6},
    };
    int bad = 0;
    for (const Case &cs : cases) {
        cudaMemset(dD, 0xff, 128 * 256 * 4);
        probe_kernel<<<1, 128, SMEM_A + SMEM_B>>>(da, db, cs, dD, dcyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-36s CUDA ERROR %s\n", cs.name, cudaGetErrorString(e)); return 3; }
        std::vector<float> D((size_t)128 * cs.n);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double err[2] = {0, 0};
        for (int hyp = 0; hyp < 2; hyp++) {
            const uint32_t al = hyp ? cs.a_sbo : cs.a_lbo, as = hyp ? cs.a_lbo : cs.a_sbo;
            const uint32_t bl = hyp ? cs.b_sbo : cs.b_lbo, bs = hyp ? cs.b_lbo : cs.b_sbo;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < cs.n; n++) {
                    float ref = 0;
                    for (int ks = 0; ks < cs.nk; ks++)
                        for (int kk = 0; kk < 16; kk++)
                            ref += elem(a, cs.a_start, al, as, cs.a_kstep, m, ks, kk, cs.swz) * elem(b, cs.b_start, bl, bs, cs.b_kstep, n, ks, kk, cs.swz);
                    const double d = fabs((double)ref - (double)D[(size_t)m * cs.n + n]);
                    if (d > err[hyp]) err[hyp] = d;
                }
        }
        const bool ok = err[0] == 0.0;
        const bool required = cs.swz == 0; // the SW128 rows are exploratory (which base_offset convention holds)
        bad += (!ok && required);
        printf("%-36s H1(LBO=K,SBO=MN) maxerr=%-8g H2(swapped) maxerr=%-8g %s\n", cs.name, err[0], err[1], ok ? "OK" : (required ? "MISMATCH" : "mismatch (exploratory)"));
    }
    // ---- throughput: cycles per MMA (M=128, K=16) issued back to back by one thread, per operand layout
    printf("\n%-44s %8s %8s\n", "timing (256 MMAs back to back)", "cyc/MMA", "floor");
    struct T { const char *name; Case c; };
    const T timing[] = {
        {"noswz dense   N=32 ", {"", 0, 2048, 128, 4096, 0, 512, 128, 1024, 32, 4, 0, 0, 64}},
        {"noswz sbo=160 N=32 ", {"", 176, 2880, 160, 5760, 0, 512, 128, 1024, 32, 4, 0, 0, 64}},
        {"SW128 dense   N=32 ", {"", 0, 16, 1024, 32, 0, 16, 1024, 32, 32, 4, 2, 0, 64}},
        {"SW128 sbo=1280 N=32", {"", 0, 16, 1280, 32, 0, 16, 1024, 32, 32, 4, 2, 0, 64}},
        {"noswz dense   N=64 ", {"", 0, 2048, 128, 4096, 0, 1024, 128, 2048, 64, 4, 0, 0, 64}},
        {"SW128 dense   N=64 ", {"", 0, 16, 1024, 32, 0, 16, 1024, 32, 64, 4, 2, 0, 64}},
        {"noswz dense   N=128", {"", 0, 2048, 128, 4096, 0, 2048, 128, 4096, 128, 4, 0, 0, 64}},
        {"SW128 dense   N=128", {"", 0, 16, 1024, 32, 0, 16, 1024, 32, 128, 4, 2, 0, 64}},
        {"noswz dense   N=256", {"", 0, 2048, 128, 4096, 0, 4096, 128, 8192, 256, 4, 0, 0, 64}},
        {"SW128 dense   N=256", {"", 0, 16, 1024, 32, 0, 16, 1024, 32, 256, 4, 2, 0, 64}},
        {"noswz N=32  2 independent accumulators", {"", 0, 2048, 128, 4096, 0, 512, 128, 1024, 32, 4, 0, 0, 64, 2}},
        {"noswz N=32  4 independent accumulators", {"", 0, 2048, 128, 4096, 0, 512, 128, 1024, 32, 4, 0, 0, 64, 4}},
        {"noswz N=32  8 independent accumulators", {"", 0, 2048, 128, 4096, 0, 512, 128, 1024, 32, 4, 0, 0, 64, 8}},
        {"noswz N=64  2 independent accumulators", {"", 0, 2048, 128, 4096, 0, 1024, 128, 2048, 64, 4, 0, 0, 64, 2}},
        {"noswz N=64  4 independent accumulators", {"", 0, 2048, 128, 4096, 0, 1024, 128, 2048, 64, 4, 0, 0, 64, 4}},
        {"noswz N=128 2 independent accumulators", {"", 0, 2048, 128, 4096, 0, 2048, 128, 4096, 128, 4, 0, 0, 64, 2}},
        {"noswz N=16  1 accumulator", {"", 0, 2048, 128, 4096, 0, 256, 128, 512, 16, 4, 0, 0, 64, 1}},
        {"noswz N=8*? N=96 1 accumulator", {"", 0, 2048, 128, 4096, 0, 1536, 128, 3072, 96, 4, 0, 0, 64, 1}},
    };
    for (const T &t : timing) {
        long long best = 1LL << 60;
        for (int rep = 0; rep < 5; rep++) {
            probe_kernel<<<1, 128, SMEM_A + SMEM_B>>>(da, db, t.c, dD, dcyc);
            cudaDeviceSynchronize();
            long long c;
            cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
            if (c < best) best = c;
        }
        printf("%-44s %8.1f %8d\n", t.name, (double)best / (t.c.reps * t.c.nk), 128 * t.c.n / 256);
    }
    printf("umma_probe: %s\n", bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
