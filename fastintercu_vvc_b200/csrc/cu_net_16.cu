// cu_net_16.cu -- tcgen05 conv kernels of the 16x16-CU network (see cu_net.cuh); one translation unit per CU size.
#include "cu_net.cuh"

namespace mlt {
template struct CuNetOps<16>;
} // namespace mlt
