// Lane-per-dimension kernels for Lotka-Volterra (used to test the fixed-point smoother against the oracle).
#include "../pdeq_dispatch_group.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K2(LotkaVolterra, 4, PDEQ_FACT_BLOCKDIAG, bd)
PDEQ_INSTANTIATE_K2(LotkaVolterra, 4, PDEQ_FACT_ISOTROPIC, iso)
}  // namespace pdeq
