// cu_net_64.cu -- tcgen05 conv kernels of the 64x64-CU network (see cu_net.cuh); one translation unit per CU size.
#include "cu_net.cuh"

namespace mlt {
template struct CuNetOps<64>;
} // namespace mlt
