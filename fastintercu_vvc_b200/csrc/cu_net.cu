// cu_net.cu -- size dispatch for the smaller-CU conv kernels (instantiated in cu_net_64.cu / cu_net_32.cu / cu_net_16.cu).
#include "cu_net.cuh"

namespace mlt {

extern template struct CuNetOps<64>;
extern template struct CuNetOps<32>;
extern template struct CuNetOps<16>;

#define MLT_CU_SIZE(size, CALL)                       \
    switch (size) {                                   \
    case 64: return CuNetOps<64>::CALL;               \
    case 32: return CuNetOps<32>::CALL;               \
    case 16: return CuNetOps<16>::CALL;               \
    default: return cudaErrorInvalidValue;            \
    }

cudaError_t cu_conv_init(int size) { MLT_CU_SIZE(size, init()) }
cudaError_t cu_conv_info(int size, int layer, CuLayerInfo *info) { MLT_CU_SIZE(size, info(layer, info)) }
cudaError_t cu_conv_prepare(int size, int layer, ConvParams *p, const __half *in, const ActLayout &in_l, const __half *x, const ActLayout *x_l)
{
    MLT_CU_SIZE(size, prepare(layer, p, in, in_l, x, x_l))
}
cudaError_t launch_cu_conv(int size, int layer, const ConvParams &p, int num_sms, cudaStream_t s) { MLT_CU_SIZE(size, launch(layer, p, num_sms, s)) }

} // namespace mlt
