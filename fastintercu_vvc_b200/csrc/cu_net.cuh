// cu_net.cuh -- the 20 3x3 convolutions of the smaller-CU networks (64 / 32 / 16-px `GapBigMltCuORPQ`,
// mlt_cu_or_pq_arch.py:59-130: BasicBlock [2,2,2,2,2], planes 32 / 64 / 96 / 128 / 256, every stage stride 2) as
// instantiations of the tcgen05 implicit-GEMM kernel of conv_umma.cuh in its STRIP mode.  One table per CU size,
// derived at compile time from the size: stage L works on (size >> (L + 1))^2 maps, down to 1x1.
#pragma once
#include "conv_umma.cuh"

namespace mlt {

constexpr int cu_planes(int L) { return L == 0 ? 32 : (L == 1 ? 64 : (L == 2 ? 96 : (L == 3 ? 128 : 256))); }
constexpr int cu_map(int size, int L) { return (size >> (L + 1)) > 1 ? (size >> (L + 1)) : 1; } // output map of stage L

// conv LI (0..19) = stage LI / 4, position LI % 4: 0 = block0.conv1 (stride 2), 1 = block0.conv2 + 1x1 stride-2 shortcut,
// 2 = block1.conv1, 3 = block1.conv2 + identity
template <int S, int LI>
struct CuSel {
    static constexpr int L = LI / 4, K = LI % 4;
    static constexpr int P = cu_planes(L), PIN = L == 0 ? 32 : cu_planes(L - 1);
    static constexpr int HIN = L == 0 ? S : cu_map(S, L - 1), H = cu_map(S, L);
    // a 3x3 stride-2 pad-1 conv (and the 1x1 stride-2 shortcut) on a 1x1 map reads exactly what the stride-1 conv reads
    static constexpr bool FAKE_S2 = HIN == 1;
    static constexpr int CIN = K == 0 ? PIN : P;
    static constexpr int STRIDE = (K == 0 && !FAKE_S2) ? 2 : 1;
    static constexpr int XC = K == 1 ? PIN : (K == 3 ? P : 0);
    static constexpr int XLO = K == 1 ? 1 : 0;
    static constexpr int OUT_PAR = (K == 3 && L < 4 && H >= 2) ? 1 : 0; // feeds the next stage's stride-2 block
    static constexpr int X_PAR = (K == 1 && !FAKE_S2) ? 1 : 0;          // shortcut input = plane (even, even) of the stage input
    // the last two stages (and the tensor that feeds them) carry fp16 hi + lo activation pairs (ConvCfg::HILO_*)
    static constexpr int HILO_OUT = LI >= cu_hilo_from(S) - 1 ? 1 : 0, HILO_IN = LI >= cu_hilo_from(S) ? 1 : 0;
    using type = ConvCfg<CIN, P, STRIDE, H, XC, OUT_PAR, XLO, 1 | (HILO_IN << 1) | (HILO_OUT << 2)>;
};

template <int S, int LI = 0, class F>
inline cudaError_t cu_dispatch(int li, F &&f)
{
    if constexpr (LI >= CU_NCONV) {
        return cudaErrorInvalidValue;
    } else {
        if (li == LI) return f(CuSel<S, LI>{});
        return cu_dispatch<S, LI + 1>(li, f);
    }
}

template <int S>
struct CuNetOps {
    static cudaError_t init()
    {
        for (int li = 0; li < CU_NCONV; li++) {
            const cudaError_t e = cu_dispatch<S>(li, [](auto sel) {
                using C = typename decltype(sel)::type;
                return cudaFuncSetAttribute(conv_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
            });
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    static cudaError_t info(int li, CuLayerInfo *o)
    {
        return cu_dispatch<S>(li, [o](auto sel) {
            using C = typename decltype(sel)::type;
            *o = CuLayerInfo{C::CIN, C::COUT, C::STRIDE, C::HOUT, C::XC, C::OUT_PAR, C::NB, C::FLAT ? 1 : 0, C::G, C::XC > 0 ? C::GX : 0,
                             C::GAP ? (C::NB == 2 ? 4 : C::TILES_PER_IMG * 4) : 0, C::HILO_OUT ? 1 : 0};
            return cudaSuccess;
        });
    }
    static cudaError_t prepare(int li, ConvParams *p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l)
    {
        return cu_dispatch<S>(li, [&](auto sel) {
            using Sel = decltype(sel);
            using C = typename Sel::type;
            if (in_l.strip <= 0 || in_l.C != C::CIN || in_l.H != C::HOUT * C::STRIDE || in_l.par != (C::STRIDE == 2)) return cudaErrorInvalidValue;
            cudaError_t e = make_act_map(&p->in_map, in, in_l, C::HILO_IN ? 2 : 1, C::BLKW, C::NB, C::PROWS, C::CH);
            if (e != cudaSuccess) return e;
            if constexpr (C::XC > 0) {
                if (!x || !x_l || x_l->strip != in_l.strip || x_l->C != C::XC || x_l->hp() != C::HOUT || (x_l->par != 0 && x_l->par != Sel::X_PAR)) return cudaErrorInvalidValue; // (par: plane 0 is read)
                e = make_act_map(&p->x_map, x, *x_l, C::HILO_IN ? 2 : 1, C::XBOXW, C::NB, C::TR, C::GX / 8);
                if (e != cudaSuccess) return e;
            } else {
                p->x_map = p->in_map;
            }
            p->x_unit_mul = (C::XC > 0 && x_l) ? x_l->npl() : 1; // planes between the hi and the lo copy of the extra operand
            p->strip_cap = in_l.strip;
            return cudaSuccess;
        });
    }
    static cudaError_t launch(int li, const ConvParams &p, int num_sms, cudaStream_t s)
    {
        return cu_dispatch<S>(li, [&](auto sel) {
            using C = typename decltype(sel)::type;
            const int ntiles = C::num_tiles(p.nimg);
            if (ntiles <= 0) return cudaSuccess;
            const int grid = ntiles < num_sms ? ntiles : num_sms; // persistent: one CTA per SM
            return launch_pdl(conv_umma_kernel<C>, dim3(grid), dim3(C::NTHREADS), C::SMEM_BYTES, s, p);
        });
    }
};

} // namespace mlt
