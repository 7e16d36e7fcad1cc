// BASELINE config 3: Pleiades (d = 28), nu = 5, block-diagonal, fixed-point smoother (and the filter).
#include "../pdeq_dispatch_group.cuh"
namespace pdeq {
PDEQ_INSTANTIATE_K2_WITH_SPEC(Pleiades, 5, PDEQ_FACT_BLOCKDIAG, bd)
}  // namespace pdeq
