/*
 * probdiffeq_b200 -- C ABI of the B200-native adaptive probabilistic IVP step loop.
 *
 * The reference (pnkraemer/probdiffeq) has no FFI: its seam is the Python `Solver` protocol
 * (probdiffeq/_ivpsolve/solver_protocols.py:33-57) plus the injected `error` and `control`
 * collaborators of `solve_adaptive_save_at` (probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:46-54).
 * This header is the boundary a `jax.ffi` / ctypes binding for that path would target: the object
 * graph (prior x strategy x constraint x solver x error x control) is lowered to the POD
 * `pdeq_config`, and the three loop entry points of the reference become three calls.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless stated otherwise; all floating point is float64.
 *  - The caller owns every buffer. The library allocates nothing that outlives a call.
 *  - Calls enqueue work on `stream` (a cudaStream_t passed as void*) and return immediately.
 *  - Return value: 0 ok; <0 invalid argument / unsupported configuration; >0 CUDA error code.
 *    `pdeq_last_error()` returns a thread-local message for the last non-zero return.
 *  - Batched layout ("ensemble of B independent IVPs", the reference's `jax.vmap` axis leads):
 *      tcoeffs      [B][n][d]        Taylor coefficients u, u', ..., u^(nu) at t0 (unnormalised)
 *      mean out     [B][T][n][d]
 *      chol out     isotropic [B][T][n][n]; blockdiag [B][T][d][n][n]; dense [B][T][nd][nd]
 *                   (left square roots, cov = L L^T, exactly the reference's `cholesky_flat`)
 *      output_scale isotropic/dense [B][T]; blockdiag [B][T][d]
 *    n = num_derivatives + 1.
 */
#ifndef PROBDIFFEQ_B200_H
#define PROBDIFFEQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEQ_VERSION 100

/* state_space_model_{isotropic,blockdiag,dense}: probdiffeq/_probdiffeq/ssm_impl_*.py */
enum { PDEQ_FACT_ISOTROPIC = 0, PDEQ_FACT_BLOCKDIAG = 1, PDEQ_FACT_DENSE = 2 };
/* constraint_ode_ts0 / constraint_ode_ts1: probdiffeq/_probdiffeq/ssm_impl_api.py:520-558 */
enum { PDEQ_CONSTRAINT_TS0 = 0, PDEQ_CONSTRAINT_TS1 = 1 };
/* solver / solver_mle / solver_dynamic: probdiffeq/_probdiffeq/solvers.py:636,318,483 */
enum { PDEQ_SOLVER_PLAIN = 0, PDEQ_SOLVER_MLE = 1, PDEQ_SOLVER_DYNAMIC = 2 };
/* strategy_filter / strategy_smoother_fixedpoint / strategy_smoother_fixedinterval:
   probdiffeq/_probdiffeq/estimators_and_losses.py:347,473,594. The fixed-interval smoother is accepted by
   pdeq_solve_fixed_grid only (the reference: "use this strategy for fixed steps").
   PDEQ_STRATEGY_FIXEDINTERVAL reproduces the reference literally: Smoother.finalize (:437-470) marginalises the
   last grid state through its own backward conditional before the backward recursion, because solve_fixed_grid
   (_ivpsolve/solvers_via_fixed_steps.py:30-32) hands it the last grid state as the "overstepped" state; the
   returned marginals are therefore one interval late at the terminal grid point.
   PDEQ_STRATEGY_FIXEDINTERVAL_ALIGNED starts the recursion from the filtering marginal at the last grid point
   (the textbook Rauch-Tung-Striebel pass; what the reference's save-every-step flow yields when it oversteps). */
enum {
  PDEQ_STRATEGY_FILTER = 0,
  PDEQ_STRATEGY_FIXEDPOINT = 1,
  PDEQ_STRATEGY_FIXEDINTERVAL = 2,
  PDEQ_STRATEGY_FIXEDINTERVAL_ALIGNED = 3
};
/* error_residual_std / error_state_std: probdiffeq/_probdiffeq/solvers.py:850,999 */
enum { PDEQ_ERROR_RESIDUAL_STD = 0, PDEQ_ERROR_STATE_STD = 1 };
/* error_norm_scale_then_rms / error_norm_rms_then_scale: probdiffeq/_probdiffeq/solvers.py:770,794 */
enum { PDEQ_NORM_SCALE_THEN_RMS = 0, PDEQ_NORM_RMS_THEN_SCALE = 1 };
/* control_integral / control_proportional_integral: probdiffeq/_ivpsolve/controllers.py:66,24 */
enum { PDEQ_CONTROL_INTEGRAL = 0, PDEQ_CONTROL_PI = 1 };

/* per-instance status written by the loops (the reference has no runtime failure channel) */
enum { PDEQ_STATUS_OK = 0, PDEQ_STATUS_NONFINITE = 1, PDEQ_STATUS_MAX_ATTEMPTS = 2 };

#define PDEQ_MAX_COEFFS 8 /* n = num_derivatives + 1 <= 8 */

typedef struct pdeq_config {
  int32_t factorisation;   /* PDEQ_FACT_*        */
  int32_t num_derivatives; /* nu; n = nu + 1     */
  int32_t ode_dim;         /* d                  */
  int32_t vf_id;           /* pdeq_vf_id(name)   */
  int32_t constraint;      /* PDEQ_CONSTRAINT_*  */
  int32_t solver;          /* PDEQ_SOLVER_*      */
  int32_t strategy;        /* PDEQ_STRATEGY_*    */
  int32_t error;           /* PDEQ_ERROR_*       */
  int32_t error_norm;      /* PDEQ_NORM_*        */
  int32_t control;         /* PDEQ_CONTROL_*     */
  int32_t clip_dt;         /* solvers_via_adaptive_steps.py:20,51 */
  int32_t derivative_idx;  /* error_state_std(derivative_idx=)  solvers.py:1022 */
  int32_t error_per_unit_step;               /* solvers.py:912,1023 */
  int32_t re_linearize_after_calibration;    /* solver_dynamic    solvers.py:496 */
  int32_t correct_asymptotic_underconfidence; /* solver_mle        solvers.py:332 */
  int32_t max_attempts;    /* guard the reference lacks; <=0 means 2^31-1 */
  /* solver(..., constraint_init=constraint) (solvers.py:361-372, 526-537, 670-680): one Bayes update with the
     constraint at t0 before the first step, its gain through a minimum-norm least-squares solve (a zero pivot of the
     observed factor gives a zero gain component, which is what linalg.lstsq_svd returns for the diagonal / scalar
     factors of the isotropic and block-diagonal models and for a dense factor whose zero pivots are decoupled). */
  int32_t constraint_init;
  int32_t reserved0;       /* keeps the doubles 8-byte aligned; must be 0 */
  double safety, factor_min, factor_max;            /* controllers.py:30-33 */
  double exponent_integral, exponent_proportional;  /* controllers.py:34-35 */
  /* IWP system matrices, computed by the host exactly like the reference does
     (probdiffeq/_probdiffeq/utilities.py:57-84: flipped Pascal A, Cholesky of flipped Hilbert Q,
     factorials via exp(lgamma)); row-major, leading dimension PDEQ_MAX_COEFFS. */
  double sys_a[PDEQ_MAX_COEFFS][PDEQ_MAX_COEFFS];
  double sys_q[PDEQ_MAX_COEFFS][PDEQ_MAX_COEFFS];
  double factorials[PDEQ_MAX_COEFFS + 1]; /* factorials[k] = k! as the reference evaluates it */
  double inv_factorials[PDEQ_MAX_COEFFS + 1]; /* 1 / factorials[k], so the kernels multiply instead of divide */
  /* Optional (all zeros = off). error_state_std applies Bayes' rule to the zero-error extrapolation
     (solvers.py:1070-1086), whose factor is diag(|p|) sqrt(dt) lambda q: a constant matrix with scaled columns. For
     the ts0 constraint (h = e_order) and damp == 0 the triangularisation R = qr_r([[0, 0], [(h q)^T, q^T]])
     therefore commutes with the scaling, and the kernels need only err_const[0] = R[0][0] (observed factor) and
     err_const[1 + i] = ||R[1:2+i, 1+i]|| (std of coefficient i), times |p_order| resp. |p_i| times
     sqrt(dt) lambda. Computed on the host from sys_q (probdiffeq_b200/_iwp.py). */
  double err_const[PDEQ_MAX_COEFFS + 1];
} pdeq_config;

/* Inputs shared by the loop entry points. A `*_stride` of 0 broadcasts one row to all instances. */
typedef struct pdeq_problem {
  int64_t num_instances;        /* B */
  const double* tcoeffs;        /* [B][n][d] */
  const double* init_std;       /* NULL = exact initial condition; isotropic [.][n], else [.][n][d] */
  int64_t init_std_stride;      /* in doubles */
  const double* prior_scale;    /* NULL = ones; isotropic [.][1], blockdiag/dense [.][d] */
  int64_t prior_scale_stride;
  const double* params;         /* [.][pdeq_vf_num_params(vf_id)] */
  int64_t params_stride;
  /* Optional (NULL = index order): a permutation of 0..B-1 in which the persistent kernels hand instances to their
     lanes / warps / CTAs. Listing expensive instances first (longest processing time first) shortens the end of a
     solve that has only a few instances per lane, and lanes of a warp that run equally long instances finish (and
     refill) together instead of one by one. Results do not depend on it. (The reference has no counterpart: under
     jax.vmap every lane runs as long as the slowest instance, probdiffeq/backend/func.py:9-10.) */
  const int32_t* order;
} pdeq_problem;

/* Outputs of the loop entry points (T checkpoints per instance). Optional pointers may be NULL. */
typedef struct pdeq_solution {
  double* t;             /* [B][T] */
  double* mean;          /* [B][T][n][d] */
  double* chol;          /* see layout note above; may be NULL */
  double* output_scale;  /* may be NULL */
  int32_t* num_steps;    /* [B][T] accepted steps up to each checkpoint (entry 0 is 0) */
  int32_t* num_attempts; /* [B] attempted steps (accepted + rejected); may be NULL */
  int32_t* status;       /* [B] PDEQ_STATUS_* */
  /* optional (smoothers): the backward conditionals of the posterior (MarkovSequence.conditional), calibrated and
     in natural coordinates (preconditioner applied, cf. LatentCond.preconditioner_apply). Entry k maps grid
     point k to k-1: x_{k-1} | x_k ~ N(gain x_k + mean, chol chol^T); entry 0 is left untouched. */
  double* bw_gain;       /* blockdiag [B][T][d][n][n], isotropic [B][T][n][n] */
  double* bw_mean;       /* [B][T][n][d] */
  double* bw_chol;       /* like chol */
  /* optional (fixed-interval smoother on a fixed grid): the calibrated filtering marginals
     (SmoothingSolution.filtering, estimators_and_losses.py:464-467), laid out like mean / chol. Dense output of a
     smoothing solution starts from them (pdeq_offgrid_marginals). */
  double* filt_mean;
  double* filt_chol;
  /* optional attempt log (NULL = off): for attempt a < trace_capacity of instance b,
     trace[(b * trace_capacity + a) * 4 + {0,1,2,3}] = {t_from, dt_used, error_power, accepted ? 1 : 0}.
     Lets a caller check the accepted-step sequence against the reference attempt by attempt. */
  double* trace;
  int64_t trace_capacity;
} pdeq_solution;

int pdeq_version(void);
const char* pdeq_last_error(void);

/* Registry of vector-field device functors (the benchmark problems of the reference:
   benchmarks/A0..A5). Returns -1 for unknown names. */
int pdeq_vf_id(const char* name);
int pdeq_vf_num_params(int vf_id);
int pdeq_vf_ode_order(int vf_id);
/* fixed dimension of the problem, or 0 if the functor accepts any d (linear, burgers) */
int pdeq_vf_dim(int vf_id);
/* Add a vector field to the registry at run time and return its id. The kernels for it come from a plug-in: a
   shared object built from csrc/ with the functor (same interface as the built-in ones in csrc/pdeq_vf.cuh) and the
   instantiation macros, which registers its launchers under this id when loaded (probdiffeq_b200/plugins.py).
   The reference takes any Python callable (`probdiffeq.ode`, problems.py:283-312); this is the device-side equivalent. */
int pdeq_register_vf(const char* name, int32_t ode_order, int32_t num_params, int32_t fixed_dim);
/* Drop every kernel registered under a run-time vector-field id (returns how many loop kernels went): called before a
   NAME is given a different right-hand side, so that no kernel of the old code can be selected for it any more. */
int pdeq_vf_clear_kernels(int vf_id);

/* 0 if (cfg) is a combination the library has a kernel for, negative otherwise (see pdeq_last_error). */
int pdeq_config_supported(const pdeq_config* cfg);

/* Bytes of device scratch `workspace` the loop entry points need for (cfg, B, T). */
size_t pdeq_workspace_bytes(const pdeq_config* cfg, int64_t num_instances, int32_t num_checkpoints);

/* jetexpand_ode_padded_scan / jetexpand_ode_unroll (probdiffeq/_probdiffeq/jet_expansion_algorithms.py:49,110):
   u0 [B][ode_order][d] -> tcoeffs [B][n][d], n = num_derivatives + 1 (unnormalised derivatives at t0). */
int pdeq_taylor_init(const pdeq_config* cfg, int64_t num_instances, const double* u0,
                     const double* params, int64_t params_stride, double t0, double* tcoeffs,
                     void* stream);

/* ivpsolve.dt0 (probdiffeq/_ivpsolve/stepsize_initialisers.py:7-21): out[B]. */
int pdeq_dt0(const pdeq_config* cfg, int64_t num_instances, const double* u0, const double* params,
             int64_t params_stride, double t0, double scale, double nugget, double* out,
             void* stream);

/* ivpsolve.dt0_adaptive (probdiffeq/_ivpsolve/stepsize_initialisers.py:24-64, Hairer et al. II.4), first-order
   ODEs: u0 [B][d] -> out[B]. */
int pdeq_dt0_adaptive(const pdeq_config* cfg, int64_t num_instances, const double* u0, const double* params,
                      int64_t params_stride, double t0, double error_contraction_rate, double rtol, double atol,
                      double* out, void* stream);

/* ivpsolve.solve_adaptive_save_at (probdiffeq/_ivpsolve/solvers_via_adaptive_steps.py:46-148);
   solve_adaptive_terminal_values (:16-43) is the T == 2 case with save_at = {t0, t1}.
   save_at [T] is shared by all instances. dt0 [.] with stride 0 or 1. */
int pdeq_solve_adaptive_save_at(const pdeq_config* cfg, const pdeq_problem* problem,
                                const double* save_at, int32_t num_checkpoints, double atol,
                                double rtol, const double* dt0, int64_t dt0_stride, double eps,
                                double damp, const pdeq_solution* solution, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ivpsolve.solve_fixed_grid (probdiffeq/_ivpsolve/solvers_via_fixed_steps.py:11-34): one step per
   grid interval, every grid point is stored. grid [T] shared by all instances. */
int pdeq_solve_fixed_grid(const pdeq_config* cfg, const pdeq_problem* problem, const double* grid,
                          int32_t num_gridpoints, double damp, const pdeq_solution* solution,
                          void* workspace, size_t workspace_bytes, void* stream);

/* loss_lml_terminal_values (probdiffeq/_probdiffeq/estimators_and_losses.py:20-50): observe Taylor
   coefficient `tcoeff_index` of N(mean, chol chol^T) through noise `std` and evaluate the log-pdf of
   `data`. mean [B][n][d], chol as above with T == 1, data [.][d], std [.][1|d], out [B]. */
int pdeq_lml_terminal_values(const pdeq_config* cfg, int64_t num_instances, int32_t tcoeff_index,
                             const double* mean, const double* chol, const double* data,
                             int64_t data_stride, const double* std, int64_t std_stride,
                             double* out, void* stream);

/* loss_lml_timeseries (probdiffeq/_probdiffeq/estimators_and_losses.py:53-105 with MarkovSequence.evaluate_lml
   :180-218): log-density of `data` [.][T][d] observed at all T grid points through noise `std`
   ([.][T] isotropic, [.][T][d] block-diagonal) under the smoothing posterior. mean / chol are the solution
   arrays of a smoother run ([B][T]...; entry T-1 is the posterior's terminal marginal), bw_* the backward
   conditionals it emitted (pdeq_solution.bw_*). average_pdfs != 0 returns the running mean of the per-point
   log-densities like the reference's default. workspace: num_instances * ode_dim doubles. out [B]. */
int pdeq_lml_timeseries(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints,
                        int32_t tcoeff_index, int32_t average_pdfs, const double* mean, const double* chol,
                        const double* bw_gain, const double* bw_mean, const double* bw_chol,
                        const double* data, int64_t data_stride, const double* std, int64_t std_stride,
                        double* out, void* workspace, size_t workspace_bytes, void* stream);

/* MarkovSequence.sample (probdiffeq/_probdiffeq/estimators_and_losses.py:233-271): joint samples of the posterior
   at all T grid points, drawn backwards through the conditionals. The standard-normal numbers are the caller's
   (`base`: isotropic [B][S][T][n] -- one draw per coefficient shared by all dimensions, as
   IsotropicNormal.sample_flat, ssm_impl_isotropic.py:255-258 -- block-diagonal [B][S][T][d][n]), so the
   arithmetic is reproducible whatever generator made them. mean / chol / bw_* as in pdeq_lml_timeseries.
   out [B][S][T][n][d]. */
int pdeq_sample_posterior(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints, int32_t num_samples,
                          const double* mean, const double* chol, const double* bw_gain, const double* bw_mean,
                          const double* bw_chol, const double* base, double* out, void* stream);

/* ProbabilisticSolver.offgrid_marginals (probdiffeq/_probdiffeq/solvers.py:149-203) with
   strategy_filter.interpolate_offgrid_marginals (estimators_and_losses.py:403-414) or
   strategy_smoother_fixedinterval.interpolate_offgrid_marginals (:677-709): dense output. For each of the Q query
   times (strictly inside the grid and not on it, as the reference requires; shared by the ensemble) the marginal is
   extrapolated from the grid point to its left with the output scale of the grid point to its right; a smoothing
   solution (cfg->strategy fixed-interval) additionally pulls the marginal of the right grid point back through the
   new conditional and needs filt_mean / filt_chol. grid [T]; mean / chol / output_scale as in pdeq_solution;
   prior_scale as in pdeq_problem (NULL = ones); out_mean [B][Q][n][d], out_chol [B][Q]([d])[n][n]. */
int pdeq_offgrid_marginals(const pdeq_config* cfg, int64_t num_instances, int32_t num_gridpoints, const double* grid,
                           int32_t num_queries, const double* queries, const double* mean, const double* chol,
                           const double* filt_mean, const double* filt_chol, const double* output_scale,
                           const double* prior_scale, int64_t prior_scale_stride, double* out_mean,
                           double* out_chol, void* stream);

/* Sum `n` doubles in place across the ranks of an NCCL communicator (ncclComm_t passed as void*):
   the ensemble log-marginal-likelihood reduction, i.e. the sum over instances of the per-instance term of
   loss_lml_terminal_values (probdiffeq/_probdiffeq/estimators_and_losses.py:20-50) when the instances live on several
   GPUs. The only collective on this path; the reference has none (its only batching is jax.vmap,
   probdiffeq/backend/func.py:9-10). */
int pdeq_allreduce_sum_f64(void* nccl_comm, double* buf, int64_t n, void* stream);

/* Communicator plumbing for hosts that do not own an ncclComm_t (a ctypes / jax.ffi host): NCCL is resolved from the
   process (dlopen of libnccl.so.2, the copy already loaded if there is one). Rank 0 calls pdeq_nccl_unique_id, the
   128-byte id travels to the other ranks by whatever side channel the host has (torch.distributed, MPI, a file), and
   every rank calls pdeq_nccl_comm_init_rank with its CUDA device current. pdeq_nccl_comm_count returns the number
   of ranks of a communicator (-1 on error). */
int pdeq_nccl_unique_id(void* id128);
int pdeq_nccl_comm_init_rank(void** nccl_comm, int32_t nranks, const void* id128, int32_t rank);
int pdeq_nccl_comm_count(void* nccl_comm);
int pdeq_nccl_comm_destroy(void* nccl_comm);

/* Device-side FP64 FMA throughput probe used for the roofline denominator: runs `iters` dependent
   FMA chains on every SM and returns the elapsed milliseconds in *ms and the FLOP count in *flops. */
int pdeq_fp64_peak_probe(int32_t iters, double* ms, double* flops, void* stream);

/* Which build of the thread-per-instance loop kernel the launcher uses when a configuration matches the
   compile-time specialised combination (adaptive + clip_dt, `solver`, error_state_std on coefficient 0, ts0,
   damp = 0, unit prior scale): 0 = the general kernel, 1..5 = the specialised builds (csrc/pdeq_loop_thread.cuh).
   The library's default can be overridden per launch through the environment variable PDEQ_K1_SPEC; every build
   returns bitwise the same results. */
int pdeq_k1_spec_choice(void);

#ifdef __cplusplus
}
#endif
#endif /* PROBDIFFEQ_B200_H */
