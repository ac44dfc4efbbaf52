// cu_net_32.cu -- tcgen05 conv kernels of the 32x32-CU network (see cu_net.cuh); one translation unit per CU size.
#include "cu_net.cuh"

namespace mlt {
template struct CuNetOps<32>;
} // namespace mlt
