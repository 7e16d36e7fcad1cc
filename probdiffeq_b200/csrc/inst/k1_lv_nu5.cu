#include "../pdeq_dispatch.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K1(LotkaVolterra, 5, 2)
}  // namespace pdeq
