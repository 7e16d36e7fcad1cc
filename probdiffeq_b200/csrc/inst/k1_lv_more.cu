// Explicit instantiations: Lotka-Volterra with other prior orders (A0 benchmark uses nu = 5).
#include "../pdeq_dispatch.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K1(LotkaVolterra, 3, 2)
}  // namespace pdeq
