#!/usr/bin/env python
"""Benchmark of the adaptive probabilistic IVP step loop (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repository)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

Headline workload: BASELINE.json configs[1] -- ONE Lotka-Volterra ensemble of 2^20 instances with randomised
parameters / initial values (seed 0), nu = 4, isotropic ts0 filter, `solver` +
`error_state_std` + PI control, terminal values on t in [0, 50], rtol 1e-6, atol 1e-8, float64.
One "step" is one solve of the whole ensemble.  For N > 1 the SAME ensemble is permuted
(`sharding.permutation`) and cut into N contiguous shards (`sharding.shard_bounds`), one rank per GPU, no
data-path collective: strong scaling (`"scaling": "strong"`).  The weak-scaling number of round 1 (every rank
solves its own 2^20-instance ensemble, seed = rank) is kept under the key `weak`.

`value`    accepted solver steps / second of the whole job with the inputs resident in HBM (CUDA events around
           the solve on the launching stream, max over ranks).
`e2e`      the same metric through the public API with pinned HOST inputs: H2D copy of parameters and initial
           values, on-device Taylor initialisation, the solve, D2H copy of terminal means and step counts.
`roofline` FP64: FLOPs of the algorithm the kernel executes (SURVEY.md section 8(d) model without the terms the
           kernel provably skips) / measured duration against the FP64 FMA peak measured in the same run
           (MEASURED_PEAKS.json carries no FP64 number); instruction-level fractions from the committed ncu
           profile are reported beside it and labelled as such.
`configs`  BASELINE configs 3, 4a, 4b and the config-5 variant at their full ensemble sizes (sharded the same way
           for N > 1): steps/s, attempts/s, roofline, e2e; the config-5 variant sums the ensemble
           log-marginal-likelihood over the ranks with `pdeq_allreduce_sum_f64` (ncclAllReduce through the C ABI).
`cpu_baseline` the plain-C restatement of the reference's algorithm (oracle/c) on the host cores, bounded sample:
           a PORT, not JAX (jax is not installable here).
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "accepted solver steps/sec (batched IVP ensemble, fp64)"
UNIT = "steps/s"
B_DEFAULT = 1 << 20
NU, D = 4, 2
T0, T1, RTOL, ATOL = 0.0, 50.0, 1e-6, 1e-8
BASE_LV = np.asarray([0.5, 0.05, 0.5, 0.05])
CPU_SAMPLE = 1 << 19  # cpu_baseline: ~10-30 s of host work
REF_SAMPLE = 1 << 17  # --impl reference: instances per step
PERM_SEED = 0
PILOT = 4096  # instances of the pilot solve the scheduling cost model is fitted on (set-up)

# Instruction mix of the headline kernel per attempted step, from the committed ncu capture named below
# (profile-derived constants, NOT measured in this run): executed FP64 FLOPs = 2 DFMA + DMUL + DADD.
K1_PROFILE = {
    "file": "profiles/r1e_k1_loop_lv_nu4_iso_ts0.ncu.txt",
    "dfma": 368, "dmul": 163, "dadd": 30, "fp64_instructions": 579, "instructions": 943,
    "fp64_pipe_active": 0.713, "dram_bytes_per_launch_2p20": 905.1e6,
}  # fmt: skip


def lv_ensemble(B: int, seed: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    draw = rng.uniform(0.8, 1.2, size=(B, 6))
    return BASE_LV[None, :] * draw[:, :4], 20.0 * draw[:, 4:]


def flops_per_attempt(n: int, d: int, error_qr: bool = True) -> float:
    """SURVEY.md section 8(d), isotropic ts0 `solver` + `error_state_std` (dense Householder model):
    prediction 2n^3 + 10n^3/3, two (n+1)^2 triangularisations (correction, error estimate), 4n^2 d, vector field.
    `error_qr=False` drops the error estimate's triangularisation: the benchmarked kernel replaces it by host
    constants (pdeq_config.err_const; ts0, damp = 0), so it is not work the kernel does."""
    c_vf = 10.0  # Lotka-Volterra right-hand side
    qr = 4 * (n + 1) ** 3 / 3
    return 2 * n**3 + 10 * n**3 / 3 + (2 if error_qr else 1) * qr + 4 * n**2 * d + c_vf


def config_dict(B_total: int, n_gpus: int) -> dict:
    return {
        "workload": "BASELINE configs[1]: ONE Lotka-Volterra ensemble of 2^20 instances, nu=4, isotropic ts0 filter, "
        "solver + error_state_std + PI control, terminal values t in [0,50], rtol=1e-6, atol=1e-8",
        "instances_total": B_total,
        "instances_per_gpu": B_total // n_gpus,
        "sharding": "permuted instance index (sharding.permutation, seed 0), contiguous shards (sharding.shard_bounds), "
        "no data-path collective",
        "l2": "256 MiB L2 flush between timed steps",
        "spinup": "untimed passes for >= 0.75 s and until three consecutive passes agree within 3 % before the W "
        "warm-up steps (clock ramp of a fresh process)",
        "launch_queue": "the K timed device-resident steps are enqueued behind a spin kernel (torch.cuda._sleep, sized from "
        "the host's measured enqueue time), so each step's CUDA-event pair brackets device work only and no step waits "
        "for its own launch (a 1/8 shard's step is 2.3 ms, about what Python needs to enqueue it); the e2e arms are not primed",
        "e2e_returns": "terminal means (B, 2) and accepted-step counts only; Cholesky factors are not computed into "
        "an output buffer nor copied back (want_cholesky=False) -- less than the reference's solution object holds",
    }


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled every 100 ms while the timed regions run: NVML in a thread of
    this process (two light queries per sample); an `nvidia-smi -lms 100` child only where NVML cannot be loaded. (The
    child used to be the only way; a fresh client attaching to the driver, and now and then one of its periodic
    nine-field queries, coincided with 10-25 ms stalls of whichever kernel was running.)"""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = None
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self.source = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip() != ""]
        if ids and all(v.strip().isdigit() for v in ids) and self.gpu_index < len(ids):
            return int(ids[self.gpu_index])
        return self.gpu_index

    def _start_nvml(self) -> bool:
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))  # fail here rather than in the thread
            int(get_reasons(handle))
            self.stop_flag = threading.Event()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)),
                                             int(get_reasons(handle))))  # fmt: skip
                    except Exception:
                        pass
                    self.stop_flag.wait(0.1)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return True
        except Exception:
            self.thread = None
            return False

    def start(self):
        if self._start_nvml():
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )  # fmt: skip
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm, mx, reasons = [], [], set()
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            for clock, bits in list(self.samples):
                sm.append(clock)
                mx.append(self.sm_max)
                for name, bit in self.REASON_BITS.items():
                    if bits & bit:
                        reasons.add(name)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            try:
                for line in open(self.path):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        sm.append(float(f[1]))
                        mx.append(float(f[2]))
                    except ValueError:
                        continue
                    for name, val in zip(names, f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
                os.unlink(self.path)
            except Exception:
                pass
        else:
            return out
        if sm:
            busy = sorted(sm)[len(sm) // 2 :]  # upper half = samples under load
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["source"] = self.source
        return out


# ---------------------------------------------------------------------------------------------------
# CPU baseline (the reference's algorithm restated in C, OpenMP over instances)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_pass(sample: int, seed: int = 0, threads: int = 0):
    from oracle import c_port
    from oracle import problems as o_problems

    params, u0 = lv_ensemble(B_DEFAULT, seed)
    params, u0 = params[:sample], u0[:sample]
    tcoeffs = o_problems.taylor_coefficients_batched("lotka_volterra", params, (u0,), T0, NU)
    c_port.load()
    # all host threads this process may use -- not OMP_NUM_THREADS, which torchrun pins to 1 for N > 1
    threads = threads or len(os.sched_getaffinity(0))

    def run():
        t = time.perf_counter()
        res = c_port.solve_lv_terminal(tcoeffs, params, t0=T0, t1=T1, atol=ATOL, rtol=RTOL, num_threads=threads)
        return time.perf_counter() - t, res

    return run, threads


def cpu_baseline_block(sample: int) -> dict:
    run, threads = cpu_reference_pass(sample)
    dt, res = run()
    return {
        "value": res["total_steps"] / dt,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"first {sample} instances of the same ensemble, one pass, {dt:.1f} s; plain-C restatement of the "
        "reference algorithm (oracle/c: dense unblocked Householder, libm pow, -O3 -march=x86-64-v3), OpenMP over "
        "instances; NOT JAX -- jax is not installable here. The port reproduces the reference's own output for this "
        "configuration (338 accepted steps, terminal value to 1e-8: tests/test_oracle_c_port.py against the "
        "reference-run fixture tests/golden/reference_numpy_backend.npz)",
        "attempts_per_s": float(res["num_attempts"].sum()) / dt,
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, threads = cpu_reference_pass(args.cpu_sample)
    for _ in range(args.warmup):
        run()
    times, steps = [], 0
    for _ in range(args.steps):
        dt, res = run()
        times.append(dt)
        steps = res["total_steps"]
    ms = 1e3 * sum(times) / len(times)
    value = steps / (ms / 1e3)
    cfg = config_dict(B_DEFAULT, args.gpus)
    cfg["reference_sample"] = f"each step solves the first {args.cpu_sample} instances of the ensemble on the host cores"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.cpu_sample} instances per step; plain-C port of the reference algorithm, not JAX"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# The other BASELINE configs (3, 4a, 4b, 5-variant): problem builders for the CUDA arm
# ---------------------------------------------------------------------------------------------------
# the largest power of two for which the ORACLE completes the ensemble (every 8th instance checked:
# 512 of 512 finite at d = 128, 2-8 k steps each; at d = 256 one instance in eight goes non-finite, at d = 1024 all do)
CONFIG5_D = 128
CONFIG5_CONSTRAINT = "ts0"


def other_configs():
    """name -> dict(host=fn(B) -> dict of host arrays with the ensemble axis leading, B, make=fn(dev inputs) -> run,
    flops=..., note=...). `run()` returns (solution, extras)."""
    import torch

    from probdiffeq_b200 import ivpsolve, probdiffeq
    from probdiffeq_b200 import problems as pb

    def c3_host(B):
        return dict(u0=pb.pleiades_ensemble(B, seed=1))

    def c3_make(dev_in):
        vf = probdiffeq.ode("pleiades")
        ssm = probdiffeq.state_space_model_blockdiag()
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_smoother_fixedpoint(), constraint=ts0)
        error = probdiffeq.error_residual_std(constraint=ts0)
        solve = ivpsolve.solve_adaptive_save_at(solver=solver, error=error, control=ivpsolve.control_integral())
        save_at = torch.linspace(0.0, 3.0, 33, dtype=torch.float64, device=dev_in["u0"].device)

        def run():
            u0 = dev_in["u0"]
            tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
            dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
            sol = solve(ssm.prior_wiener_integrated(tcoeffs), save_at=save_at, atol=1e-9, rtol=1e-6, dt0=dt0,
                        want_posterior=False)
            return sol, {}

        return run

    def c4a_host(B):
        rng = np.random.Generator(np.random.PCG64(2))
        u0 = np.repeat(pb.HIRES_U0[None, :], B, axis=0)
        sc = rng.uniform(0.9, 1.1, size=(B, 2))
        u0[:, 0] *= sc[:, 0]
        u0[:, 7] *= sc[:, 1]
        return dict(u0=u0)

    def c4a_make(dev_in):
        vf = probdiffeq.ode("hires")
        ssm = probdiffeq.state_space_model_dense()
        ts1 = ssm.constraint_ode_ts1(vf)
        solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
        error = probdiffeq.error_residual_std(constraint=ts1)
        solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error,
                                                        control=ivpsolve.control_proportional_integral())

        def run():
            u0 = dev_in["u0"]
            tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=5)(vf, (u0,), t=0.0)
            dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
            sol = solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=321.8122, atol=1e-11, rtol=1e-8, dt0=dt0,
                        want_cholesky=False)
            return sol, {}

        return run

    def c4b_host(B):
        rng = np.random.Generator(np.random.PCG64(2))
        rng.uniform(0.9, 1.1, size=(16384, 2))  # continue the stream of config 4a
        return dict(u0=2.0 * rng.uniform(0.9, 1.1, size=(B, 1)), du0=np.zeros((B, 1)), params=np.full((B, 1), 1e3))

    def c4b_make(dev_in):
        ssm = probdiffeq.state_space_model_dense()

        def run():
            vf = probdiffeq.ode("vanderpol", params=dev_in["params"])
            ts1 = ssm.constraint_ode_ts1(vf)
            solver = probdiffeq.solver_dynamic(strategy=probdiffeq.strategy_filter(), constraint=ts1)
            error = probdiffeq.error_state_std(constraint=ts1)
            solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=ivpsolve.control_integral())
            tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (dev_in["u0"], dev_in["du0"]), t=0.0)
            sol = solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=6.3, atol=1e-11, rtol=1e-8, want_cholesky=False)
            return sol, {}

        return run

    d5 = CONFIG5_D

    def c5_host(B):
        rng = np.random.Generator(np.random.PCG64(3))
        nu = 0.01 * rng.uniform(0.5, 2.0, size=(B, 1))
        return dict(params=nu, u0=np.repeat(pb.burgers_u0(d5)[None, :], B, axis=0))

    def c5_make(dev_in):
        ssm = probdiffeq.state_space_model_blockdiag()
        lml = probdiffeq.loss_lml_terminal_values()
        dev = dev_in["u0"].device

        def solve_for(params, u0):
            vf = probdiffeq.ode("burgers", params=params)
            cons = ssm.constraint_ode_ts0(vf) if CONFIG5_CONSTRAINT == "ts0" else ssm.constraint_ode_ts1(vf)
            solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=cons)
            error = probdiffeq.error_state_std(constraint=cons)
            solve = ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error,
                                                            control=ivpsolve.control_proportional_integral())
            tcoeffs, _ = probdiffeq.jetexpand_ode_padded_scan(num=3)(vf, (u0,), t=0.0)
            dt0 = ivpsolve.dt0(vf, (u0,), t=0.0)
            # the step size of the explicit (ts0) linearisation is stability-limited at dx^2 / viscosity, so the
            # viscosity itself ranks the instances by cost (1.9 k ... 7.6 k steps): longest first
            hint = params[:, 0] if params.shape[0] > 1 else None
            return solve(ssm.prior_wiener_integrated(tcoeffs), t0=0.0, t1=1.0, atol=1e-7, rtol=1e-4, dt0=dt0,
                         cost_hint=hint)

        # data: terminal mean of the viscosity-0.01 instance (solved by this same path) + 1e-2 N(0, 1), std = 1e-2
        ref = solve_for(torch.full((1, 1), 0.01, dtype=torch.float64, device=dev), dev_in["u0"][:1])
        noise = np.random.Generator(np.random.PCG64(33)).normal(size=(1, d5))
        data = ref.u.mean[0][:1] + 1e-2 * torch.from_numpy(noise).to(dev)
        std = torch.full((1, d5), 1e-2, dtype=torch.float64, device=dev)

        def run():
            sol = solve_for(dev_in["params"], dev_in["u0"])
            ll = lml(data, marginals=sol.u, std=std)
            return sol, {"lml_local": ll.sum().reshape(1)}

        return run

    n3, n4, n5 = 6, 6, 4
    N4 = n4 * 8
    f3 = 28 * (4 * (2 * n3) ** 3 / 3 + n3**3 + 4 * n3**3 + 10 * n3**3 / 3 + 4 * (n3 + 1) ** 3 / 3) + 1.1e3
    f4a = 2 * N4**3 + 10 * N4**3 / 3 + 2 * 8 * N4**2 + 4 * (N4 + 8) ** 3 / 3 + 8**2 * N4
    f4b = flops_per_attempt(5, 1)
    err_qr5 = CONFIG5_CONSTRAINT != "ts0"  # ts0, damp = 0: the error estimate's triangularisation is host constants
    f5 = d5 * (2 * n5**3 + 10 * n5**3 / 3 + (2 if err_qr5 else 1) * 4 * (n5 + 1) ** 3 / 3 + 4 * n5**2) + 6.0 * d5
    io5 = 8.0 * (n5 * d5 + 1) + 2 * 8.0 * (1 + n5 * d5 + d5 * n5 * n5 + d5) + 4 * 4.0  # per instance: in + 2 checkpoints out
    return {
        "3": dict(B=65536, host=c3_host, make=c3_make, flops_per_attempt=f3, kernel="k2_loop_kernel<Pleiades,5,blockdiag,ts0,fixedpoint>",
                  launches=5 + 1 + 1,
                  workload="BASELINE configs[2]: Pleiades d=28, nu=5, blockdiag ts0, fixed-point smoother, solver_dynamic + "
                           "error_residual_std + I control, save_at = linspace(0,3,33), rtol=1e-6, atol=1e-9, dt0 from dt0()",
                  flops_model="SURVEY 8(d): d * (4(2n)^3/3 + n^3 + 4n^3 + 10n^3/3 + 4(n+1)^3/3) + c_vf, n=6, d=28"),
        "4a": dict(B=16384, host=c4a_host, make=c4a_make, flops_per_attempt=f4a, kernel="k3_loop_kernel<Hires,5,ts1>",
                   launches=5 + 1 + 1,
                   workload="BASELINE configs[3] (a): HIRES d=8, nu=5, dense ts1 (analytic Jacobian), filter, solver_dynamic + "
                            "error_residual_std + PI control, terminal values t1=321.8122, rtol=1e-8, atol=1e-11",
                   flops_model="SURVEY 8(d): 2N^3 + 10N^3/3 + 2dN^2 + 4(N+d)^3/3 + d^2 N, N=48, d=8"),
        "4b": dict(B=16384, host=c4b_host, make=c4b_make, flops_per_attempt=f4b, kernel="k1_loop_kernel<VanDerPol,4,isotropic,1,ts1>",
                   launches=3 + 1,
                   workload="BASELINE configs[3] (b): Van der Pol, second-order form, stiffness 1e3, nu=4, dense ts1 (d=1), "
                            "solver_dynamic + error_state_std + I control, terminal values t1=6.3, rtol=1e-8, atol=1e-11",
                   flops_model="isotropic model of SURVEY 8(d) with n=5, d=1 (16384 instances fill 29 % of the resident "
                               "lanes: this config measures the latency of the slowest instance, not throughput)"),
        "5": dict(B=4096, host=c5_host, make=c5_make, flops_per_attempt=f5, kernel="k2_loop_kernel<Burgers,3,blockdiag,%s,filter>" % CONFIG5_CONSTRAINT,
                  launches=3 + 1 + 1 + 1, io_bytes_per_instance=io5, lml=True,
                  workload="BASELINE configs[4] VARIANT: Burgers semi-discretisation with d=%d instead of 1024 (the specified "
                           "d=1024 blockdiag ts0 solve does not complete in the reference algorithm: DESIGN.md section 7), nu=3, "
                           "blockdiag %s filter, solver + error_state_std + PI control, terminal values t1=1, rtol=1e-4, atol=1e-7, "
                           "viscosity 0.01 U(0.5,2); ensemble log-marginal-likelihood (loss_lml_terminal_values, std 1e-2) "
                           "summed over GPUs with ncclAllReduce; instances served stiffest-first (cost_hint = viscosity)"
                           % (d5, CONFIG5_CONSTRAINT),
                  flops_model="SURVEY 8(d): d * (2n^3 + 10n^3/3 + %s 4(n+1)^3/3 + 4n^2) + c_vf, n=4, d=%d"
                              % ("2 *" if err_qr5 else "", d5)),
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------------
# The CUDA arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import ctypes as C

    import torch
    import torch.distributed as dist

    from probdiffeq_b200 import _lib, ivpsolve, probdiffeq, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B_total = args.instances
    launches = 0  # kernels of this library launched inside timed regions (counted where they are issued)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    per_step = {}

    def timed(fn, steps, warmup, tag=None):
        # Warm-up holds the previous step's result while the next one is produced, exactly as the timed loop below does
        # (`last`): both sets of output buffers then sit in torch's caching allocator, and no timed step pays for a
        # cudaMalloc (which would stall the launch behind it for tens of milliseconds).
        held = cur = None
        host_ms = 0.0
        for _ in range(max(warmup, 2)):
            h0 = time.perf_counter()
            cur = fn()
            host_ms = (time.perf_counter() - h0) * 1e3  # how long the HOST takes to enqueue one step
            flush.fill_(1)
            held = cur
        # BOTH names go: a surviving `cur` kept a third set of output buffers alive, and the second timed step then
        # paid for a cudaMalloc with the device idle inside its event pair (r2o, r2y: 16.9, 41.7, 16.9, ... ms)
        del held, cur
        barrier()
        if tag != "cfg" and os.environ.get("PDEQ_BENCH_NO_PRIME") != "1":
            # A shard's step is a few milliseconds at N = 8 -- the order of what Python needs to enqueue it (with eight
            # ranks on sixteen host cores: r2x measured 2.9, 2.8, 2.3, 2.3, 2.3 ms, the first steps waiting for their own
            # launch). Keep the device busy with a spin kernel while the host enqueues the K steps, so that every event
            # pair brackets device work only. (The e2e arms below are NOT primed: there the host is part of the step.)
            prime_ms = min(250.0, 1.5 * steps * (host_ms + 0.3) + 2.0)
            torch.cuda._sleep(int(prime_ms * torch.cuda.get_device_properties(dev).clock_rate))
        evs = []
        last = None
        for _ in range(steps):
            flush.fill_(0)  # evict L2 between timed steps (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            last = fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        if tag and tag != "cfg":
            per_step[tag] = [round(x, 3) for x in ms]
        per_step["_last"] = ms
        return sum(ms) / len(ms), last

    def timed_pipelined(fn, steps, warmup):
        """The same K end-to-end steps issued on two alternating CUDA streams (the public API runs on torch's
        current stream): step i+1's host->device copy and Taylor pass overlap step i's loop kernel, and step i's
        device->host read overlaps step i+1's kernel, as in any double-buffered serving loop. Every step still copies
        its own inputs from pinned host memory and reads its own result back; the time is the device time from before
        the first step to after the last, divided by K."""
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        cur = torch.cuda.current_stream(dev)
        for i in range(max(warmup, 2)):
            with torch.cuda.stream(streams[i % 2]):
                fn(i % 2)
        barrier()
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(streams[i % 2]):
                fn(i % 2)
        for st in streams:
            cur.wait_stream(st)
        e1.record(cur)
        barrier()
        return e0.elapsed_time(e1) / steps

    def reduce_max(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(*vals):
        if world == 1:
            return [int(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.tolist()

    # =============================================================================================
    # headline: BASELINE configs[1], ONE ensemble split over the ranks
    # =============================================================================================
    ssm = probdiffeq.state_space_model_isotropic()
    jetexpand = probdiffeq.jetexpand_ode_padded_scan(num=NU)

    def build_solver(vf):
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        control = ivpsolve.control_proportional_integral()
        return ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=control)

    def lv_arms(params_np, u0_np):
        """Resident and end-to-end step functions for one shard."""
        Bs = params_np.shape[0]
        params_h = torch.from_numpy(np.ascontiguousarray(params_np)).pin_memory()
        u0_h = torch.from_numpy(np.ascontiguousarray(u0_np)).pin_memory()
        mean_hs = [torch.empty((Bs, D), dtype=torch.float64).pin_memory() for _ in range(2)]
        steps_hs = [torch.empty((Bs,), dtype=torch.int32).pin_memory() for _ in range(2)]
        params_d, u0_d = params_h.to(dev), u0_h.to(dev)
        vf_d = probdiffeq.ode("lotka_volterra", params=params_d)
        tcoeffs_d, _ = jetexpand(vf_d, (u0_d,), t=T0)
        prior_d = ssm.prior_wiener_integrated(tcoeffs_d)
        solve_d = build_solver(vf_d)

        # Scheduling hint (set-up, untimed): a quadratic model of the attempt count in (params, u0), fitted on a pilot
        # solve of the shard's first PILOT instances. Inside every timed step the model is EVALUATED for the whole
        # shard and the instances are sorted by it (longest first) -- that work is part of the step.
        model = None
        if not args.no_cost_hint and ivpsolve.cost_hint_is_used(Bs, dev):  # large shards: the hint would be ignored
            M = min(PILOT, Bs)
            vf_p = probdiffeq.ode("lotka_volterra", params=params_d[:M])
            tc_p, _ = jetexpand(vf_p, (u0_d[:M],), t=T0)
            pilot = build_solver(vf_p)(ssm.prior_wiener_integrated(tc_p), t0=T0, t1=T1, atol=ATOL, rtol=RTOL,
                                       want_cholesky=False)
            model = sharding.QuadraticCostModel.fit(
                np.concatenate([params_np[:M], u0_np[:M]], axis=1), pilot.num_attempts.cpu().numpy().astype(np.float64))
            del pilot

        def hint_for(p, u):
            return None if model is None else model.predict(torch.cat([p, u], dim=1))

        def step_resident():
            return solve_d(prior_d, t0=T0, t1=T1, atol=ATOL, rtol=RTOL, cost_hint=hint_for(params_d, u0_d))

        def step_plain():
            return solve_d(prior_d, t0=T0, t1=T1, atol=ATOL, rtol=RTOL)

        def step_e2e(slot: int = 0):
            p = params_h.to(dev, non_blocking=True)
            u = u0_h.to(dev, non_blocking=True)
            vf = probdiffeq.ode("lotka_volterra", params=p)
            tc, _ = jetexpand(vf, (u,), t=T0)
            prior = ssm.prior_wiener_integrated(tc)
            sol = build_solver(vf)(prior, t0=T0, t1=T1, atol=ATOL, rtol=RTOL, want_cholesky=False,
                                   cost_hint=hint_for(p, u))
            mean_hs[slot].copy_(sol.u.mean[0], non_blocking=True)
            steps_hs[slot].copy_(sol.num_steps, non_blocking=True)
            return sol

        h2d = int(params_h.numel() * 8 + u0_h.numel() * 8)
        d2h = int(mean_hs[0].numel() * 8 + steps_hs[0].numel() * 4)
        return step_resident, step_e2e, steps_hs, h2d, d2h, step_plain, model is not None

    params_all, u0_all = lv_ensemble(B_total, seed=0)
    if world > 1:
        (params_np, u0_np), _idx = sharding.shard((params_all, u0_all), rank, world, seed=PERM_SEED)
    else:
        params_np, u0_np = params_all, u0_all
    B = params_np.shape[0]
    step_resident, step_e2e, steps_hs, h2d_bytes, d2h_bytes, step_plain, hint_active = lv_arms(params_np, u0_np)

    # Untimed spin-up on top of the W warm-up steps: a fresh process finds the GPU at idle clocks, and W = 3 passes
    # can end before the clocks have ramped -- a whole run then reads ~30 % slow. Keep the device busy for at least
    # 0.75 s and until three consecutive passes agree within 3 % (at most 400 passes) before anything is timed.
    # The clock sampler (an `nvidia-smi -lms 100` child) starts BEFORE the spin-up: its start-up (a new client attaching
    # to the GPU) stalled whichever kernel was running for 10-15 ms -- once, ~100 ms after the fork, which used to be
    # the second timed step (r2o: 16.95, 30.90, 16.87, 16.87, 16.86 ms). Its periodic queries do not show in the times.
    sampler = ClockSampler(local_rank)
    sampler.start()
    spin_t0, spinup, spin_ms = time.perf_counter(), 0, []
    while spinup < 400:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        torch.cuda.synchronize()
        spin_ms.append(e0.elapsed_time(e1))
        spinup += 1
        settled = len(spin_ms) >= 3 and max(spin_ms[-3:]) <= 1.03 * min(spin_ms[-3:])
        if time.perf_counter() - spin_t0 >= 0.75 and settled:
            break

    ms_res, sol = timed(step_resident, args.steps, args.warmup, "resident")
    k1_launches = 1  # one persistent loop kernel per solve (the cost model and the sort are torch kernels)
    launches += args.steps * k1_launches
    ms_plain = None
    if hint_active:
        ms_plain, _ = timed(step_plain, args.steps, args.warmup, "resident_no_hint")
        launches += args.steps * k1_launches
    steps_pass = int(sol.num_steps.sum().item())
    attempts_pass = int(sol.num_attempts.sum().item())
    bad = int((sol.status != 0).sum().item())
    ms_e2e_serial, _ = timed(step_e2e, args.steps, max(args.warmup, 1), "e2e_serial")
    e2e_steps_pass = int(steps_hs[0].to(torch.int64).sum().item())
    ms_e2e = timed_pipelined(step_e2e, args.steps, max(args.warmup, 1))
    launches += 2 * args.steps * (k1_launches + NU)
    clocks = sampler.stop()
    assert int(steps_hs[1].to(torch.int64).sum().item()) == e2e_steps_pass == int(steps_hs[0].to(torch.int64).sum().item())
    del sol

    # FP64 FMA peak, measured now on this GPU
    pms, pfl = C.c_double(0), C.c_double(0)
    _lib.check(lib.pdeq_fp64_peak_probe(1 << 17, C.byref(pms), C.byref(pfl), None), "fp64 probe")
    fp64_peak_tflops = pfl.value / (pms.value * 1e-3) / 1e12

    ms_rank = ms_res
    ms_res, ms_e2e, ms_e2e_serial = reduce_max(ms_res, ms_e2e, ms_e2e_serial)
    if ms_plain is not None:
        (ms_plain,) = reduce_max(ms_plain)
    steps_all, attempts_all, e2e_steps_all, bad = reduce_sum(steps_pass, attempts_pass, e2e_steps_pass, bad)

    # ---- the weak-scaling number of round 1 (N > 1 only): every rank its own 2^20-instance ensemble, seed = rank
    weak = None
    if world > 1 and not args.no_weak:
        wp, wu = lv_ensemble(B_total, seed=rank)
        w_resident, _w_e2e, _sh, _a, _b, _w_plain, _w_hint = lv_arms(wp, wu)
        ms_w, wsol = timed(w_resident, max(2, args.steps // 2), 2, "weak")
        launches += max(2, args.steps // 2) * k1_launches
        (ms_w,) = reduce_max(ms_w)
        (w_steps,) = reduce_sum(int(wsol.num_steps.sum().item()))
        weak = {"value": w_steps / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w,
                "instances_per_gpu": B_total, "what": "every rank solves its own 2^20-instance ensemble (seed = rank)"}  # fmt: skip
        del wsol, w_resident

    # =============================================================================================
    # the other configs
    # =============================================================================================
    comm = None
    configs_out = {}
    wanted = [c for c in args.configs.split(",") if c]
    if wanted:
        table = other_configs()
        for name in wanted:
            spec = table[name]
            Bc = spec["B"] if args.config_instances <= 0 else min(spec["B"], args.config_instances)
            try:
                host_all = spec["host"](Bc)
                keys = list(host_all)
                if world > 1:
                    shards, _ = sharding.shard([host_all[k] for k in keys], rank, world, seed=PERM_SEED)
                    host = dict(zip(keys, shards))
                else:
                    host = host_all
                pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in host.items()}
                dev_in = {k: v.to(dev) for k, v in pinned.items()}
                run = spec["make"](dev_in)
                ms_c, (csol, extra) = timed(run, args.config_steps, 1, "cfg")
                n_passes = args.config_steps
                launches += n_passes * spec["launches"]
                if reduce_max(ms_c)[0] < 200.0:  # decided collectively: every rank must time the same passes
                    # a short pass (configs 4b, 5): one 10-20 ms hiccup of the box (seen about once per 40 launches, also
                    # in the headline's per-step list) would be most of it -- time more passes and report their median
                    n_passes = max(args.config_steps, min(15, int(600.0 / max(reduce_max(ms_c)[0], 1.0))))
                    _, (csol, extra) = timed(run, n_passes, 1, "cfg")
                    launches += n_passes * spec["launches"]
                    ms_c = statistics.median(per_step["_last"])
                Bl = csol.num_steps.shape[0]
                c_steps = int(csol.num_steps.reshape(Bl, -1)[:, -1].sum().item())
                c_att = int(csol.num_attempts.sum().item())
                c_bad = int((csol.status != 0).sum().item())

                # end to end: pinned host inputs -> device -> Taylor init, dt0, solve (, lml) -> host
                mean_h = torch.empty(csol.u.mean[0].reshape(Bl, -1, csol.u.mean[0].shape[-1])[:, -1].shape, dtype=torch.float64).pin_memory()
                nsteps_h = torch.empty((Bl,), dtype=torch.int32).pin_memory()
                lml_h = torch.empty((1,), dtype=torch.float64).pin_memory()

                def run_e2e(spec=spec, pinned=pinned, mean_h=mean_h, nsteps_h=nsteps_h, lml_h=lml_h, Bl=Bl):
                    d_in = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
                    s2, ex = spec["make"](d_in)()
                    m0 = s2.u.mean[0]
                    mean_h.copy_(m0.reshape(Bl, -1, m0.shape[-1])[:, -1], non_blocking=True)
                    nsteps_h.copy_(s2.num_steps.reshape(Bl, -1)[:, -1], non_blocking=True)
                    if "lml_local" in ex:
                        lml_h.copy_(ex["lml_local"], non_blocking=True)
                    return s2

                ms_ce, _ = timed(run_e2e, max(1, args.config_steps // 2), 1, "cfg")
                launches += max(1, args.config_steps // 2) * spec["launches"]
                ms_c_rank = ms_c
                ms_c, ms_ce = reduce_max(ms_c, ms_ce)
                c_steps, c_att, c_bad = reduce_sum(c_steps, c_att, c_bad)
                fl = spec["flops_per_attempt"]
                tfl = fl * c_att / (ms_c * 1e-3) / 1e12 / world  # per GPU
                block = {
                    "workload": spec["workload"], "instances_total": Bc, "kernel": spec["kernel"],
                    "ms_per_pass": ms_c, "value": c_steps / (ms_c * 1e-3), "unit": UNIT,
                    "attempts_per_s": c_att / (ms_c * 1e-3), "accepted_steps_per_pass": c_steps,
                    "attempts_per_pass": c_att, "rejection_ratio": 1.0 - c_steps / max(c_att, 1),
                    "failed_instances": c_bad, "timed_passes": n_passes,
                    "statistic": "mean over the timed passes" if n_passes == args.config_steps else "median over the timed passes",
                    "timed_region": "H2D-resident u0/params -> Taylor init, dt0, solve" + (", lml" if spec.get("lml") else ""),
                    "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                                 "frac": tfl / fp64_peak_tflops, "flops_per_attempt": fl, "flops_model": spec["flops_model"],
                                 "per": "GPU", "traffic": None,
                                 "peak_source": "pdeq_fp64_peak_probe (FP64 FMA, measured in this run)"},
                    "e2e": {"value": c_steps / (ms_ce * 1e-3), "unit": UNIT, "ms_per_step": ms_ce,
                            "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in pinned.values())),
                            "d2h_bytes_per_step": int(mean_h.numel() * 8 + nsteps_h.numel() * 4 + (8 if spec.get("lml") else 0))},
                }  # fmt: skip
                if "io_bytes_per_instance" in spec:
                    peaks_path = ROOT / "MEASURED_PEAKS.json"
                    hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if peaks_path.exists() else 6650.0
                    gbs = spec["io_bytes_per_instance"] * Bl / (ms_c_rank * 1e-3) / 1e9
                    block["roofline"]["hbm"] = {
                        "algorithmic_bytes_per_launch": spec["io_bytes_per_instance"] * Bl, "achieved_gbs": gbs,
                        "peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                        "note": "the carried state stays in shared memory for the whole solve: only inputs and the two "
                                "checkpoints cross HBM, so the FP64 pipe binds, not HBM",
                        "peak_source": "MEASURED_PEAKS.json" if peaks_path.exists() else "fallback 6.65 TB/s"}  # fmt: skip
                if spec.get("lml"):
                    # the one collective on the path: sum the per-rank log-marginal-likelihoods with ncclAllReduce
                    # issued through the C ABI, on a communicator created through the C ABI
                    if comm is None:
                        comm = sharding.Communicator.from_torch_distributed(dev)
                    local = extra["lml_local"].clone()
                    via_abi = comm.allreduce_sum(local.clone())
                    via_torch = sharding.allreduce_sum(local.clone())
                    torch.cuda.synchronize()
                    block["ensemble_lml"] = float(via_abi.item())
                    block["lml_allreduce"] = {
                        "via": "pdeq_allreduce_sum_f64 -> ncclAllReduce(sum, f64) on a pdeq_nccl_comm_init_rank communicator",
                        "comm_nranks": comm.count(), "torch_all_reduce": float(via_torch.item()),
                        "abs_diff_vs_torch": abs(float(via_abi.item()) - float(via_torch.item())),
                        "local_lml_rank0": float(local.item())}  # fmt: skip
                configs_out[name] = block
                del csol, run, dev_in
            except Exception as exc:  # one config failing must not hide the headline
                configs_out[name] = {"error": repr(exc)[:400]}
            torch.cuda.empty_cache()

    if rank == 0:
        n = NU + 1
        spec_id = lib.pdeq_k1_spec_choice()
        shortcut = spec_id != 0  # the specialised builds never triangularise for the error estimate
        fl_model = flops_per_attempt(n, D, error_qr=True)
        fl = flops_per_attempt(n, D, error_qr=not shortcut)
        kernel_s = ms_res * 1e-3
        achieved_tflops = fl * attempts_all / kernel_s / 1e12 / world  # per GPU
        fl_exec = 2 * K1_PROFILE["dfma"] + K1_PROFILE["dmul"] + K1_PROFILE["dadd"]
        exec_tflops = fl_exec * attempts_all / kernel_s / 1e12 / world
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if peaks_path.exists() else 6650.0
        bytes_algo = B * 8.0 * (n * D + 4 + (2 * (1 + n * D + n * n + 1))) + B * 4.0 * (2 + 2)
        line = {
            "metric": METRIC, "value": steps_all / kernel_s, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(B_total, world),
            "accepted_steps_per_pass": steps_all, "attempts_per_pass": attempts_all,
            "rejection_ratio": 1.0 - steps_all / max(attempts_all, 1), "failed_instances": bad,
            "attempts_per_s": attempts_all / kernel_s,
            "schedule": {
                "cost_hint": hint_active,
                "what": ("every timed step evaluates a quadratic cost model of (params, u0) for its shard and serves the "
                         "instances longest-predicted-first (solve(..., cost_hint=) -> pdeq_problem.order; a lane's first "
                         "ticket is its thread index, so a warp starts on 32 neighbours of that order). The model was "
                         "fitted in set-up (untimed) on a pilot solve of the shard's first %d instances" % PILOT)
                        if hint_active else
                        "index order: with more than %d instances per resident lane the solver ignores a cost hint "
                        "(ivpsolve.cost_hint_is_used), so none is computed" % ivpsolve.COST_HINT_MAX_ROUNDS,
                "no_hint": None if ms_plain is None else {
                    "value": steps_all / (ms_plain * 1e-3), "ms_per_step": ms_plain,
                    "what": "the same solve in index order on the full grid (round 1's schedule)"},
            },
            "e2e": {
                "value": e2e_steps_all / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes * world if world > 1 else h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes * world if world > 1 else d2h_bytes,
                "returns": "terminal means and accepted-step counts (no Cholesky factors: want_cholesky=False)",
                "how": "K public-API steps on two alternating CUDA streams (double-buffered pinned result buffers): "
                       "each step copies its inputs host->device, runs the Taylor pass and the loop kernel, and reads "
                       "its terminal values and step counts back; device time over the K steps / K",
                "serial": {"value": e2e_steps_all / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial,
                           "how": "the same steps one after the other on one stream, L2 flushed in between"},
            },
            "gpu_launches": launches,
            "gpu_launches_note": "kernels of this library issued inside timed regions on rank 0: 1 loop kernel per "
                                 "resident step (both schedules); 1 loop + 4 Taylor-pass kernels per e2e step (two e2e "
                                 "arms); per config pass: Taylor passes + dt0 + loop (+ lml)",
            "roofline": {
                "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved_tflops / fp64_peak_tflops, "per": "GPU",
                "flops_per_attempt": fl, "attempts_per_launch": attempts_all // world,
                "flops_model": "SURVEY 8(d) dense-Householder model of the executed algorithm: prediction 2n^3 + 10n^3/3, "
                               "ONE (n+1)^2 triangularisation (the correction), 4n^2 d, vector field"
                               + ("; the error estimate's triangularisation is NOT counted -- the kernel replaces it by host constants"
                                  if shortcut else "; plus the error estimate's triangularisation (general kernel)"),
                "frac_survey_model_incl_skipped_qr": fl_model * attempts_all / kernel_s / 1e12 / world / fp64_peak_tflops,
                "frac_executed": exec_tflops / fp64_peak_tflops,
                "frac_executed_note": "executed FP64 FLOPs per attempt (2 DFMA + DMUL + DADD = %d) from the ncu capture %s "
                                      "(profile-derived, not measured in this run); FP64 pipe active there: %.3f"
                                      % (fl_exec, K1_PROFILE["file"], K1_PROFILE["fp64_pipe_active"]),
                "traffic": None,
                "traffic_note": "not measurable inside this run; the ncu capture %s reports %.1f MB per 2^20-instance launch "
                                "(dram__bytes_read.sum + dram__bytes_write.sum)" % (K1_PROFILE["file"], K1_PROFILE["dram_bytes_per_launch_2p20"] / 1e6),
                "kernel": f"k1_loop_kernel<LotkaVolterra,4,isotropic,2,ts0,SPEC={spec_id}>",
                "peak_source": "pdeq_fp64_peak_probe (FP64 FMA, measured in this run; MEASURED_PEAKS.json has no FP64 figure)",
                "hbm": {"algorithmic_bytes_per_launch": bytes_algo, "achieved_gbs": bytes_algo / (ms_rank * 1e-3) / 1e9,
                        "peak_gbs": hbm_peak, "frac": bytes_algo / (ms_rank * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks_path.exists() else "fallback 6.65 TB/s"},
            },
            "clocks": clocks,
            "ms_steps": {k: v for k, v in per_step.items() if not k.startswith("_")},
        }  # fmt: skip
        if weak is not None:
            line["weak"] = weak
        if configs_out:
            line["configs"] = configs_out
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_block(args.cpu_sample)
            except Exception as exc:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}
        print(json.dumps(line))
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    # The driver reads ONE JSON line from stdout. Libraries write there too (NCCL prints its version banner on the first
    # communicator): file descriptor 1 is pointed at stderr for the whole run and `print` keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=B_DEFAULT, help="instances of the WHOLE ensemble (split over the GPUs)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="instances in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cost-hint", action="store_true", help="index order on the full grid (no scheduling hint)")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling pass for N > 1")
    ap.add_argument("--configs", default="3,4a,4b,5", help="other BASELINE configs to time ('' = none)")
    ap.add_argument("--config-steps", type=int, default=2, help="timed passes per other config")
    ap.add_argument("--config-instances", type=int, default=0, help="cap the other configs' ensembles (0 = full size)")
    args = ap.parse_args()
    if args.cpu_sample is None:
        args.cpu_sample = REF_SAMPLE if args.impl == "reference" else CPU_SAMPLE
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
