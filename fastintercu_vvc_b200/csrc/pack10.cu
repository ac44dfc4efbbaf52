// pack10.cu -- 10-bit packed transport of Pel samples (host side: pack, device side: unpack).
//
// The path's only host<->device traffic that matters is the int16 org / pred luma (64 KiB per CTU, EncCu.cpp:810-830 reads
// `Pel = int16_t`, TypeDef.h:277).  With InternalBitDepth 10 (cfg/encoder_randomaccess_vtm.cfg) every sample is in [0, 1023], so
// six of the sixteen bits are zero: the packed form is a little-endian bit stream, sample i in bits [10 i, 10 i + 10), i.e. 4
// samples per 5 bytes, 40 KiB per CTU (-37.5 %).  On a multi-GPU host the aggregate H2D rate is what bounds the end-to-end
// throughput of the batch API (profiles/r01: 29-35 GB/s per rank at 4-8 ranks), so fewer bytes per CTU is the only lever.
// The producers (one encoder process per encode) pack their own blocks with mlt_pack10 -- AVX2 where the host has it (5.4 us per CTU on one core of
// the GPU box's host against ~14 us for the scalar loop; 5.8 against 29 us in the build container), against
// seconds of RDO per CTU -- and the device unpacks into the dense int16 batch the stem kernel reads (one HBM-bound pass:
// 40 KiB in + 64 KiB out per CTU, ~1 % of a step).  Samples outside [0, 1023] cannot be packed: mlt_pack10 counts them and the
// caller falls back to the int16 entry points (the reference's staging treats such values through the (uint16_t) cast,
// EncCu.cpp:816,827 -- kept exactly by the int16 path only).
#include "mlt_internal.h"

namespace mlt {

// one thread = 64 samples = 80 packed bytes (five 16-byte loads) -> 128 bytes out (eight 16-byte stores)
__global__ void __launch_bounds__(256) unpack10_kernel(const uint4 *__restrict__ packed, uint4 *__restrict__ out, size_t groups)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    uint32_t w[21];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint4 v = __ldg(packed + g * 5 + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    w[20] = 0;
#pragma unroll
    for (int o = 0; o < 8; o++) {
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int s0 = o * 8 + 2 * k, b0 = 10 * s0, b1 = b0 + 10;
            const uint32_t v0 = __funnelshift_r(w[b0 >> 5], w[(b0 >> 5) + 1], b0 & 31) & 0x3FFu;
            const uint32_t v1 = __funnelshift_r(w[b1 >> 5], w[(b1 >> 5) + 1], b1 & 31) & 0x3FFu;
            r[k] = v0 | (v1 << 16);
        }
        out[g * 8 + o] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

cudaError_t launch_unpack10(const uint8_t *packed, int16_t *out, size_t samples, cudaStream_t s)
{
    const size_t groups = samples / 64; // callers pass whole CTUs (32768 samples)
    if (groups == 0) return cudaSuccess;
    unpack10_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4 *>(packed), reinterpret_cast<uint4 *>(out), groups);
    return cudaGetLastError();
}

} // namespace mlt

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define MLT_PACK10_AVX2 1
// 16 samples -> 20 bytes per step: pmaddwd pairs two 10-bit fields into 20 bits, a 64-bit shift / or pairs those into 40 bits, one
// byte shuffle per 128-bit lane compacts 2 x 5 bytes; the two 16-byte stores overlap (10 valid bytes each), so the caller leaves
// the last 24 samples to the scalar loop.  Returns the OR of every raw sample it saw (range check).
__attribute__((target("avx2"))) static uint32_t pack10_avx2(const int16_t *src, uint64_t blocks, uint8_t *dst)
{
    const __m256i m10 = _mm256_set1_epi16(0x03FF), mul = _mm256_set1_epi32(0x04000001), m20 = _mm256_set1_epi64x(0xFFFFF);
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, 3, 4, 8, 9, 10, 11, 12, -1, -1, -1, -1, -1, -1, 0, 1, 2, 3, 4, 8, 9, 10, 11, 12, -1, -1, -1, -1, -1, -1);
    __m256i any = _mm256_setzero_si256();
    for (uint64_t b = 0; b < blocks; b++, src += 16, dst += 20) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
        any = _mm256_or_si256(any, x);
        const __m256i y = _mm256_madd_epi16(_mm256_and_si256(x, m10), mul);                     // 8 x (a | b << 10)
        const __m256i z = _mm256_or_si256(_mm256_and_si256(y, m20), _mm256_slli_epi64(_mm256_srli_epi64(y, 32), 20)); // 4 x 40 bits
        const __m256i c = _mm256_shuffle_epi8(z, shuf);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst), _mm256_castsi256_si128(c));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 10), _mm256_extracti128_si256(c, 1));
    }
    alignas(32) uint16_t lanes[16];
    _mm256_store_si256(reinterpret_cast<__m256i *>(lanes), any);
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r |= lanes[i];
    return r;
}
#endif

extern "C" {

// Host: pack `count` int16 samples (count % 4 == 0) into count * 10 / 8 bytes; returns the number of samples outside [0, 1023]
// (their low 10 bits are stored; a non-zero return means this block must go through the int16 entry points instead).
MLT_API uint64_t mlt_pack10(const int16_t *src, uint64_t count, uint8_t *dst)
{
    uint64_t bad = 0, i0 = 0;
    uint32_t any = 0;
#ifdef MLT_PACK10_AVX2
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && count >= 40) {
        const uint64_t blocks = (count - 24) / 16; // the vector stores write 6 bytes past their 20: keep >= 24 samples for the scalar tail
        any = pack10_avx2(src, blocks, dst);
        i0 = blocks * 16;
        dst += blocks * 20;
    }
#endif
    for (uint64_t i = i0; i + 4 <= count; i += 4, dst += 5) {
        const uint32_t a = (uint16_t)src[i], b = (uint16_t)src[i + 1], c = (uint16_t)src[i + 2], d = (uint16_t)src[i + 3];
        any |= a | b | c | d;
        const uint64_t v = (uint64_t)(a & 0x3FFu) | ((uint64_t)(b & 0x3FFu) << 10) | ((uint64_t)(c & 0x3FFu) << 20) | ((uint64_t)(d & 0x3FFu) << 30);
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); dst[4] = (uint8_t)(v >> 32);
    }
    if (any & ~0x3FFu)
        for (uint64_t i = 0; i < count; i++) bad += ((uint16_t)src[i] & ~0x3FFu) != 0;
    return bad;
}

} // extern "C"
