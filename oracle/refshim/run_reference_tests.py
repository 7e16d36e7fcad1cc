"""Run the REFERENCE's own tests (/root/reference/tests) that concern this path on the NumPy backend of oracle/refshim
(`pytest_cases`, which the suite is built on, is not installed: backend/testing.py carries a small stand-in; modules
that need `diffeqzoo` or exercise parts outside the path -- MAP Taylor points, jet-lifted residuals, the other priors,
matrix-free models -- are not in the list): the reference's test-suite is what validates the backend the parity fixtures were produced on.
TEST INFRASTRUCTURE; needs the reference's sources, so it runs in the build container only.

    python -m oracle.refshim.run_reference_tests [extra pytest arguments]

The result is committed as profiles/r3k_reference_own_tests_on_numpy_backend.txt.
"""
import sys

import pytest

from oracle import refshim

# modules of the reference's suite that collect without pytest_cases and concern the path this repository serves
MODULES = (
    "test_ivpsolve/test_controllers.py",
    "test_ivpsolve/test_solve_adaptive_terminal_values.py",
    "test_ivpsolve/test_solve_fixed_grid.py",
    "test_probdiffeq/test_calibration/test_dynamic_across_factorisations.py",
    "test_probdiffeq/test_calibration/test_dynamic_vs_mle.py",
    "test_probdiffeq/test_constraints/test_ode_second_order.py",
    "test_probdiffeq/test_logpdf.py",
    "test_probdiffeq/test_priors/test_wiener_integrated.py",
    "test_probdiffeq/test_strategies/test_warnings_for_wrong_strategies.py",
    "test_util/test_cholesky_util.py",
    # ... and, through the small stand-in for pytest_cases in backend/testing.py (fixture / case /
    # parametrize_with_cases), the modules built on case functions:
    "test_ivpsolve/test_solve_adaptive_save_at.py",
    "test_ivpsolve/test_pytree_output_structure.py",
    "test_probdiffeq/test_calibration/test_mle.py",
    "test_probdiffeq/test_losses/test_lml_terminal_values.py",
    "test_probdiffeq/test_losses/test_lml_timeseries.py",
    "test_probdiffeq/test_strategies/test_smoother_fixedinterval_vs_fixedpoint.py",
    "test_probdiffeq/test_strategies/test_filter_vs_smoother.py",
    "test_probdiffeq/test_dense_output/test_interpolate_filter.py",
    "test_probdiffeq/test_dense_output/test_interpolate_smoother.py",
    "test_probdiffeq/test_dense_output/test_offgrid_marginals_vs_solve_and_save_at.py",
    "test_probdiffeq/test_dense_output/test_behaviour_close_to_t1.py",
    "test_probdiffeq/test_priors/test_is_exact.py",
    "test_probdiffeq/test_priors/test_output_scales.py",
    "test_probdiffeq/test_priors/test_diffuse_derivatives.py",
    "test_probdiffeq/test_sample.py",
    "test_probdiffeq/test_jacobian_handling.py",
    "test_probdiffeq/test_constraints/test_ts1_vs_residual.py",
    "test_probdiffeq/test_jetexpand/test_ode.py",  # (its three-body cases need diffeqzoo and skip)
)


# Six tests of those modules exercise JAX's AUTODIFF RULES (gradient of a log-pdf through a QR decomposition, jacrev at
# a zero matrix, the derivative of hypot at the origin) rather than the algorithm; the complex-step derivatives of this
# backend do not reproduce those rules, so they are deselected by name.
NOT_APPLICABLE = "not gradient_is_finite and not jacrev_zero_matrix and not hypot_derivative"


def main(argv):
    if not refshim.available():
        raise SystemExit("needs the reference sources under /root/reference")
    refshim.load()
    sys.dont_write_bytecode = True
    files = [str(refshim.REFERENCE / "tests" / m) for m in MODULES]
    return pytest.main([*files, "-q", "-p", "no:cacheprovider", "--import-mode=importlib", "--rootdir=/tmp",
                        "-c", "/dev/null", "-W", "ignore", "-k", NOT_APPLICABLE, *argv])  # fmt: skip


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
