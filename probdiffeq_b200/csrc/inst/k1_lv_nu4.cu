// Explicit instantiations: Lotka-Volterra (d = 2), nu = 4 -- the headline configuration.
#include "../pdeq_dispatch.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K1_WITH_SPEC(LotkaVolterra, 4, 2)
}  // namespace pdeq
