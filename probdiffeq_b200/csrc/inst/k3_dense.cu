// Dense-factorisation kernels: HIRES (BASELINE config 4a: d = 8, nu = 5, ts1) and Lotka-Volterra (tests).
#include "../pdeq_dispatch_dense.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K3(Hires, 5)
PDEQ_INSTANTIATE_K3(LotkaVolterra, 4)
PDEQ_INSTANTIATE_K3(LotkaVolterra, 3)
}  // namespace pdeq
