// Van der Pol as a second-order scalar ODE (BASELINE config 4b): d = 1, so dense == isotropic == blockdiag.
#include "../pdeq_dispatch.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K1(VanDerPol, 4, 1)
}  // namespace pdeq
