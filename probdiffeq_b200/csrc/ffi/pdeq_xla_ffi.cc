// XLA FFI shim: one handler per C-ABI loop entry point, so that the reference can call the kernels from inside
// jax.jit with jax.ffi.ffi_call (see INTEGRATION.md section 4 for the reference-side Python).
//
// NOT compiled by probdiffeq_b200/build.py and never RUN in this repository: JAX (hence xla/ffi/api/ffi.h) is not
// available in the build or GPU images. What is checked here (tests/test_ffi_shim_compiles.py) is that it compiles
// against a local stand-in of the xla::ffi names it uses (tests/stubs/xla/ffi/api/ffi.h): every pdeq_* call matches
// include/probdiffeq_b200.h and every handler's signature equals its binding. Build it where JAX is installed:
//
//   g++ -O2 -std=c++17 -fPIC -shared -DPDEQ_WITH_XLA_FFI -I"$(python -c 'import jax; print(jax.ffi.include_dir())')" \
//       -I include -I /usr/local/cuda/include probdiffeq_b200/csrc/ffi/pdeq_xla_ffi.cc \
//       -L probdiffeq_b200/lib -lprobdiffeq_b200 -o libpdeq_xla_ffi.so
//
// The POD pdeq_config travels as a byte-string attribute (the Python side builds it exactly like
// probdiffeq_b200.ivpsolve._lower); buffers are the batched arrays a vmap of the reference's solve would carry.
#ifdef PDEQ_WITH_XLA_FFI

#include <cstring>
#include <string_view>

#include <cuda_runtime_api.h>

#include "probdiffeq_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status_of(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, pdeq_last_error());
}

pdeq_config config_of(std::string_view bytes) {
  pdeq_config cfg;
  std::memset(&cfg, 0, sizeof cfg);
  std::memcpy(&cfg, bytes.data(), bytes.size() < sizeof cfg ? bytes.size() : sizeof cfg);
  return cfg;
}

// ivpsolve.solve_adaptive_save_at / solve_adaptive_terminal_values (pdeq_solve_adaptive_save_at)
ffi::Error SolveSaveAt(cudaStream_t stream, std::string_view cfg_bytes, ffi::Buffer<ffi::F64> tcoeffs,
                       ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> save_at, ffi::Buffer<ffi::F64> dt0,
                       double atol, double rtol, double eps, double damp, ffi::Buffer<ffi::U8> workspace,
                       ffi::ResultBuffer<ffi::F64> t, ffi::ResultBuffer<ffi::F64> mean,
                       ffi::ResultBuffer<ffi::F64> chol, ffi::ResultBuffer<ffi::F64> scale,
                       ffi::ResultBuffer<ffi::S32> num_steps, ffi::ResultBuffer<ffi::S32> num_attempts,
                       ffi::ResultBuffer<ffi::S32> status) {
  const pdeq_config cfg = config_of(cfg_bytes);
  pdeq_problem pr{};
  pr.num_instances = tcoeffs.dimensions()[0];
  pr.tcoeffs = tcoeffs.typed_data();
  pr.params = params.element_count() > 0 ? params.typed_data() : nullptr;
  pr.params_stride = params.dimensions().size() > 1 ? params.dimensions().back() : 0;
  pdeq_solution so{};
  so.t = t->typed_data();
  so.mean = mean->typed_data();
  so.chol = chol->typed_data();
  so.output_scale = scale->typed_data();
  so.num_steps = num_steps->typed_data();
  so.num_attempts = num_attempts->typed_data();
  so.status = status->typed_data();
  return status_of(pdeq_solve_adaptive_save_at(
      &cfg, &pr, save_at.typed_data(), (int32_t)save_at.element_count(), atol, rtol, dt0.typed_data(),
      dt0.element_count() > 1 ? 1 : 0, eps, damp, &so, workspace.untyped_data(), workspace.size_bytes(), stream));
}

// ivpsolve.solve_fixed_grid (pdeq_solve_fixed_grid)
ffi::Error SolveFixedGrid(cudaStream_t stream, std::string_view cfg_bytes, ffi::Buffer<ffi::F64> tcoeffs,
                          ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> grid, double damp,
                          ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F64> t,
                          ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> chol,
                          ffi::ResultBuffer<ffi::F64> scale, ffi::ResultBuffer<ffi::S32> num_steps,
                          ffi::ResultBuffer<ffi::S32> status) {
  const pdeq_config cfg = config_of(cfg_bytes);
  pdeq_problem pr{};
  pr.num_instances = tcoeffs.dimensions()[0];
  pr.tcoeffs = tcoeffs.typed_data();
  pr.params = params.element_count() > 0 ? params.typed_data() : nullptr;
  pr.params_stride = params.dimensions().size() > 1 ? params.dimensions().back() : 0;
  pdeq_solution so{};
  so.t = t->typed_data();
  so.mean = mean->typed_data();
  so.chol = chol->typed_data();
  so.output_scale = scale->typed_data();
  so.num_steps = num_steps->typed_data();
  so.status = status->typed_data();
  return status_of(pdeq_solve_fixed_grid(&cfg, &pr, grid.typed_data(), (int32_t)grid.element_count(), damp, &so,
                                         workspace.untyped_data(), workspace.size_bytes(), stream));
}

// jetexpand_ode_padded_scan (pdeq_taylor_init)
ffi::Error TaylorInit(cudaStream_t stream, std::string_view cfg_bytes, ffi::Buffer<ffi::F64> u0,
                      ffi::Buffer<ffi::F64> params, double t0, ffi::ResultBuffer<ffi::F64> tcoeffs) {
  const pdeq_config cfg = config_of(cfg_bytes);
  return status_of(pdeq_taylor_init(&cfg, u0.dimensions()[0], u0.typed_data(),
                                    params.element_count() > 0 ? params.typed_data() : nullptr,
                                    params.dimensions().size() > 1 ? params.dimensions().back() : 0, t0,
                                    tcoeffs->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    PdeqSolveSaveAt, SolveSaveAt,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<std::string_view>("cfg")
        .Arg<ffi::Buffer<ffi::F64>>()  // tcoeffs [B][n][d]
        .Arg<ffi::Buffer<ffi::F64>>()  // params  [B][P] or [P]
        .Arg<ffi::Buffer<ffi::F64>>()  // save_at [T]
        .Arg<ffi::Buffer<ffi::F64>>()  // dt0     [B] or [1]
        .Attr<double>("atol")
        .Attr<double>("rtol")
        .Attr<double>("eps")
        .Attr<double>("damp")
        .Arg<ffi::Buffer<ffi::U8>>()  // workspace (pdeq_workspace_bytes)
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    PdeqSolveFixedGrid, SolveFixedGrid,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<std::string_view>("cfg")
        .Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::F64>>()
        .Attr<double>("damp")
        .Arg<ffi::Buffer<ffi::U8>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PdeqTaylorInit, TaylorInit,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<std::string_view>("cfg")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<double>("t0")
                                  .Ret<ffi::Buffer<ffi::F64>>());

#endif  // PDEQ_WITH_XLA_FFI
