"""ctypes binding of libprobdiffeq_b200.so (the C ABI declared in include/probdiffeq_b200.h).

There is deliberately no fallback: if the shared object is missing or a call fails, an exception is
raised.  The product path never touches ``oracle/``.
"""

from __future__ import annotations

import ctypes as C
import pathlib

MAX_COEFFS = 8

FACT = {"isotropic": 0, "blockdiag": 1, "dense": 2}
CONSTRAINT = {"ts0": 0, "ts1": 1}
SOLVER = {"solver": 0, "solver_mle": 1, "solver_dynamic": 2}
STRATEGY = {"filter": 0, "fixedpoint": 1, "fixedinterval": 2, "fixedinterval_aligned": 3}
ERROR = {"residual_std": 0, "state_std": 1}
NORM = {"scale_then_rms": 0, "rms_then_scale": 1}
CONTROL = {"integral": 0, "proportional_integral": 1}


class Config(C.Structure):
    _fields_ = [
        ("factorisation", C.c_int32),
        ("num_derivatives", C.c_int32),
        ("ode_dim", C.c_int32),
        ("vf_id", C.c_int32),
        ("constraint", C.c_int32),
        ("solver", C.c_int32),
        ("strategy", C.c_int32),
        ("error", C.c_int32),
        ("error_norm", C.c_int32),
        ("control", C.c_int32),
        ("clip_dt", C.c_int32),
        ("derivative_idx", C.c_int32),
        ("error_per_unit_step", C.c_int32),
        ("re_linearize_after_calibration", C.c_int32),
        ("correct_asymptotic_underconfidence", C.c_int32),
        ("max_attempts", C.c_int32),
        ("constraint_init", C.c_int32),
        ("reserved0", C.c_int32),
        ("safety", C.c_double),
        ("factor_min", C.c_double),
        ("factor_max", C.c_double),
        ("exponent_integral", C.c_double),
        ("exponent_proportional", C.c_double),
        ("sys_a", (C.c_double * MAX_COEFFS) * MAX_COEFFS),
        ("sys_q", (C.c_double * MAX_COEFFS) * MAX_COEFFS),
        ("factorials", C.c_double * (MAX_COEFFS + 1)),
        ("inv_factorials", C.c_double * (MAX_COEFFS + 1)),
        ("err_const", C.c_double * (MAX_COEFFS + 1)),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("num_instances", C.c_int64),
        ("tcoeffs", C.c_void_p),
        ("init_std", C.c_void_p),
        ("init_std_stride", C.c_int64),
        ("prior_scale", C.c_void_p),
        ("prior_scale_stride", C.c_int64),
        ("params", C.c_void_p),
        ("params_stride", C.c_int64),
        ("order", C.c_void_p),
    ]


class Solution(C.Structure):
    _fields_ = [
        ("t", C.c_void_p),
        ("mean", C.c_void_p),
        ("chol", C.c_void_p),
        ("output_scale", C.c_void_p),
        ("num_steps", C.c_void_p),
        ("num_attempts", C.c_void_p),
        ("status", C.c_void_p),
        ("bw_gain", C.c_void_p),
        ("bw_mean", C.c_void_p),
        ("bw_chol", C.c_void_p),
        ("filt_mean", C.c_void_p),
        ("filt_chol", C.c_void_p),
        ("trace", C.c_void_p),
        ("trace_capacity", C.c_int64),
    ]


# every symbol include/probdiffeq_b200.h declares: (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "pdeq_version": (C.c_int, []),
    "pdeq_last_error": (C.c_char_p, []),
    "pdeq_vf_id": (C.c_int, [C.c_char_p]),
    "pdeq_vf_num_params": (C.c_int, [C.c_int]),
    "pdeq_vf_ode_order": (C.c_int, [C.c_int]),
    "pdeq_vf_dim": (C.c_int, [C.c_int]),
    "pdeq_register_vf": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_int32]),
    "pdeq_vf_clear_kernels": (C.c_int, [C.c_int]),
    "pdeq_config_supported": (C.c_int, [_P(Config)]),
    "pdeq_workspace_bytes": (C.c_size_t, [_P(Config), C.c_int64, C.c_int32]),
    "pdeq_taylor_init": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p],
    ),
    "pdeq_dt0": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
         C.c_void_p, C.c_void_p],
    ),  # fmt: skip
    "pdeq_dt0_adaptive": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
         C.c_void_p, C.c_void_p],
    ),  # fmt: skip
    "pdeq_solve_adaptive_save_at": (
        C.c_int,
        [_P(Config), _P(Problem), C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_int64,
         C.c_double, C.c_double, _P(Solution), C.c_void_p, C.c_size_t, C.c_void_p],
    ),  # fmt: skip
    "pdeq_solve_fixed_grid": (
        C.c_int,
        [_P(Config), _P(Problem), C.c_void_p, C.c_int32, C.c_double, _P(Solution), C.c_void_p, C.c_size_t,
         C.c_void_p],
    ),  # fmt: skip
    "pdeq_lml_terminal_values": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
         C.c_int64, C.c_void_p, C.c_void_p],
    ),  # fmt: skip
    "pdeq_lml_timeseries": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),  # fmt: skip
    "pdeq_sample_posterior": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),  # fmt: skip
    "pdeq_offgrid_marginals": (
        C.c_int,
        [_P(Config), C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    ),  # fmt: skip
    "pdeq_allreduce_sum_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pdeq_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "pdeq_nccl_comm_init_rank": (C.c_int, [_P(C.c_void_p), C.c_int32, C.c_void_p, C.c_int32]),
    "pdeq_nccl_comm_count": (C.c_int, [C.c_void_p]),
    "pdeq_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "pdeq_fp64_peak_probe": (C.c_int, [C.c_int32, _P(C.c_double), _P(C.c_double), C.c_void_p]),
    "pdeq_k1_spec_choice": (C.c_int, []),
}

import os

# PDEQ_B200_LIB lets experiments point at an alternative build of the same library (never a fallback).
LIB_PATH = pathlib.Path(os.environ.get("PDEQ_B200_LIB") or pathlib.Path(__file__).resolve().parent / "lib" / "libprobdiffeq_b200.so")

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the shared object (once). Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryError(
            f"{LIB_PATH} is missing. Build it with `python -m probdiffeq_b200.build` "
            "(or `__graft_entry__.build()`); probdiffeq_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # raises AttributeError if the export is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pdeq_last_error().decode()
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise NativeLibraryError(f"{what}: {msg} (CUDA/NCCL code {rc})")
