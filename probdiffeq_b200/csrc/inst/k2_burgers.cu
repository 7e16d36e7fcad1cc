// BASELINE config 5: Burgers (d up to 1024), nu = 3, block-diagonal filter.
#include "../pdeq_dispatch_group.cuh"
namespace pdeq {
static K2Registrar<Burgers, 3, PDEQ_FACT_BLOCKDIAG, true, false> _k2_burgers_bd_ts0;
static K2Registrar<Burgers, 3, PDEQ_FACT_BLOCKDIAG, false, false> _k2_burgers_bd_ts1;
static K2Registrar<Linear, 3, PDEQ_FACT_BLOCKDIAG, true, false> _k2_linear_bd_ts0;
static K2Registrar<Linear, 3, PDEQ_FACT_ISOTROPIC, true, false> _k2_linear_iso_ts0;
}  // namespace pdeq
