"""ctypes binding of the plain-C restatement (oracle/c/pdeq_oracle.c). Test infrastructure / CPU baseline."""

from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

from oracle import linalg

HERE = pathlib.Path(__file__).resolve().parent / "c"
LIB = HERE / "libpdeq_oracle.so"
NMAX = 8


class Cfg(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("d", C.c_int), ("control_pi", C.c_int), ("clip_dt", C.c_int),
        ("safety", C.c_double), ("fmin_", C.c_double), ("fmax_", C.c_double),
        ("exp_i", C.c_double), ("exp_p", C.c_double),
        ("A", (C.c_double * NMAX) * NMAX), ("Q", (C.c_double * NMAX) * NMAX), ("fact", C.c_double * (NMAX + 1)),
    ]  # fmt: skip


def build(force: bool = False) -> pathlib.Path:
    src = HERE / "pdeq_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "libpdeq_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.pdeq_oracle_lv_terminal.restype = C.c_int64
        _lib.pdeq_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(load().pdeq_oracle_max_threads())


def make_cfg(num_derivatives: int, control: str = "pi", clip_dt: bool = True) -> Cfg:
    cfg = Cfg()
    n = num_derivatives + 1
    cfg.n, cfg.d = n, 2
    cfg.control_pi = 1 if control == "pi" else 0
    cfg.clip_dt = int(clip_dt)
    cfg.safety, cfg.fmin_, cfg.fmax_, cfg.exp_i, cfg.exp_p = 0.95, 0.2, 10.0, 0.3, 0.4
    A, Q = linalg.system_matrices_1d_iwp(num_derivatives)
    for i in range(n):
        for j in range(n):
            cfg.A[i][j] = A[i, j]
            cfg.Q[i][j] = Q[i, j]
    f = linalg.factorial(np.arange(n + 1))
    for k in range(n + 1):
        cfg.fact[k] = f[k]
    return cfg


def solve_lv_terminal(tcoeffs, params, *, t0, t1, atol, rtol, dt0=0.1, eps=1e-8, damp=0.0, control="pi",
                      clip_dt=True, num_threads=0):  # fmt: skip
    """Isotropic ts0 filter, `solver` + `error_state_std`, terminal values, Lotka-Volterra ensemble."""
    tcoeffs = np.ascontiguousarray(tcoeffs, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64)
    B, n, d = tcoeffs.shape
    assert d == 2 and params.shape == (B, 4)
    cfg = make_cfg(n - 1, control, clip_dt)
    mean = np.empty((B, n, d))
    chol = np.empty((B, n, n))
    t = np.empty(B)
    steps = np.empty(B, dtype=np.int32)
    attempts = np.empty(B, dtype=np.int32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    total = load().pdeq_oracle_lv_terminal(
        C.byref(cfg), C.c_int64(B), ptr(tcoeffs), ptr(params), C.c_double(t0), C.c_double(t1), C.c_double(atol),
        C.c_double(rtol), C.c_double(dt0), C.c_double(eps), C.c_double(damp), ptr(mean), ptr(chol), ptr(t),
        ptr(steps), ptr(attempts), C.c_int32(num_threads),
    )  # fmt: skip
    return dict(mean=mean, chol=chol, t=t, num_steps=steps, num_attempts=attempts, total_steps=int(total))
