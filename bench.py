#!/usr/bin/env python
"""Benchmark of the adaptive probabilistic IVP step loop (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repository)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

Workload (N = 1): BASELINE.json configs[1] -- a Lotka-Volterra ensemble of 2^20 instances with randomised
parameters / initial values (seed 0), nu = 4, isotropic ts0 filter, `solver` +
`error_state_std` + PI control, terminal values on t in [0, 50], rtol 1e-6, atol 1e-8, float64.
One "step" is one solve of the whole ensemble.  For N > 1 every rank solves its own 2^20-instance shard
(seed = rank; instances are independent, there is no data-path collective): weak scaling.

`value`    accepted solver steps / second with the inputs resident in HBM (CUDA events around the solve).
`e2e`      the same metric through the public API with pinned HOST inputs: H2D copy of parameters and initial
           values, on-device Taylor initialisation, the solve, D2H copy of terminal values and step counts.
`roofline` FP64: algorithmic FLOPs of the loop kernel / its measured duration against the FP64 FMA peak measured
           in the same run (MEASURED_PEAKS.json carries no FP64 number); HBM traffic is reported beside it.
`cpu_baseline` the plain-C restatement of the reference's algorithm (oracle/c) on the host cores, bounded sample.
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


def lv_ensemble(B: int, seed: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    draw = rng.uniform(0.8, 1.2, size=(B, 6))
    return BASE_LV[None, :] * draw[:, :4], 20.0 * draw[:, 4:]


def flops_per_attempt(n: int, d: int) -> float:
    """SURVEY.md section 8(d), isotropic ts0 `solver` + `error_state_std` (dense Householder model)."""
    c_vf = 10.0  # Lotka-Volterra right-hand side
    return 2 * n**3 + 10 * n**3 / 3 + 2 * 4 * (n + 1) ** 3 / 3 + 4 * n**2 * d + c_vf


def config_dict(B: int, n_gpus: int) -> dict:
    return {
        "workload": "BASELINE configs[1]: Lotka-Volterra ensemble, nu=4, isotropic ts0 filter, solver + "
        "error_state_std + PI control, terminal values t in [0,50], rtol=1e-6, atol=1e-8",
        "instances_per_gpu": B,
        "instances_total": B * n_gpus,
        "sharding": "by instance index, no collective",
        "l2": "256 MiB L2 flush between timed steps; per-step inputs+outputs (0.4 GB) exceed the 126 MB L2",
        "spinup": "untimed passes for >= 0.75 s and until three consecutive passes agree within 3 % before the W warm-up steps (clock ramp of a fresh process)",
    }


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
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
        if sm:
            busy = sorted(sm)[len(sm) // 2 :]  # upper half = samples under load
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
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
        "reference algorithm (oracle/c), OpenMP over instances; JAX is not installable here so the reference itself cannot run",
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
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.cpu_sample} instances per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# The CUDA arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from probdiffeq_b200 import _lib, ivpsolve, probdiffeq

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
    B = args.instances

    params_np, u0_np = lv_ensemble(B, seed=rank)
    params_h = torch.from_numpy(params_np).pin_memory()
    u0_h = torch.from_numpy(u0_np).pin_memory()
    mean_h = torch.empty((B, D), dtype=torch.float64).pin_memory()
    steps_h = torch.empty((B,), dtype=torch.int32).pin_memory()

    ssm = probdiffeq.state_space_model_isotropic()
    jetexpand = probdiffeq.jetexpand_ode_padded_scan(num=NU)

    def build_solver(vf):
        ts0 = ssm.constraint_ode_ts0(vf)
        solver = probdiffeq.solver(strategy=probdiffeq.strategy_filter(), constraint=ts0)
        error = probdiffeq.error_state_std(constraint=ts0)
        control = ivpsolve.control_proportional_integral()
        return ivpsolve.solve_adaptive_terminal_values(solver=solver, error=error, control=control)

    # ---- device-resident arm: inputs already in HBM ----
    params_d = params_h.to(dev)
    u0_d = u0_h.to(dev)
    vf_d = probdiffeq.ode("lotka_volterra", params=params_d)
    tcoeffs_d, _ = jetexpand(vf_d, (u0_d,), t=T0)
    prior_d = ssm.prior_wiener_integrated(tcoeffs_d)
    solve_d = build_solver(vf_d)

    def step_resident():
        return solve_d(prior_d, t0=T0, t1=T1, atol=ATOL, rtol=RTOL)

    # ---- end-to-end arm: pinned host inputs -> public API -> host results ----
    # Two sets of pinned result buffers: the pipelined arm below has two steps in flight.
    mean_hs = [mean_h, torch.empty((B, D), dtype=torch.float64).pin_memory()]
    steps_hs = [steps_h, torch.empty((B,), dtype=torch.int32).pin_memory()]

    def step_e2e(slot: int = 0):
        p = params_h.to(dev, non_blocking=True)
        u = u0_h.to(dev, non_blocking=True)
        vf = probdiffeq.ode("lotka_volterra", params=p)
        tc, _ = jetexpand(vf, (u,), t=T0)
        prior = ssm.prior_wiener_integrated(tc)
        sol = build_solver(vf)(prior, t0=T0, t1=T1, atol=ATOL, rtol=RTOL, want_cholesky=False)
        mean_hs[slot].copy_(sol.u.mean[0], non_blocking=True)
        steps_hs[slot].copy_(sol.num_steps, non_blocking=True)
        return sol

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        # Warm-up holds the previous step's result while the next one is produced, exactly as the timed loop below does
        # (`last`): both sets of output buffers then sit in torch's caching allocator, and no timed step pays for a
        # cudaMalloc (which would stall the launch behind it for tens of milliseconds).
        held = None
        for _ in range(max(warmup, 2)):
            cur = fn()
            flush.fill_(1)
            held = cur
        del held, cur
        barrier()
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
        per_step.append([round(x, 3) for x in ms])
        return sum(ms) / len(ms), last

    per_step = []

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

    # Untimed spin-up on top of the W warm-up steps: a fresh process finds the GPU at idle clocks, and W = 3 passes
    # (~80 ms) can end before the clocks have ramped -- a whole run then reads ~30 % slow. Keep the device busy for
    # at least 0.75 s and until three consecutive passes agree within 3 % (at most 80 passes) before anything is timed.
    import time as _time

    spin_t0, spinup, spin_ms = _time.perf_counter(), 0, []
    while spinup < 80:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        torch.cuda.synchronize()
        spin_ms.append(e0.elapsed_time(e1))
        spinup += 1
        # stop once the device has been busy for 0.75 s AND the last three passes agree within 3 % (clocks settled)
        settled = len(spin_ms) >= 3 and max(spin_ms[-3:]) <= 1.03 * min(spin_ms[-3:])
        if _time.perf_counter() - spin_t0 >= 0.75 and settled:
            break

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_res, sol = timed(step_resident, args.steps, args.warmup)
    steps_pass = int(sol.num_steps.sum().item())
    attempts_pass = int(sol.num_attempts.sum().item())
    bad = int((sol.status != 0).sum().item())
    ms_e2e_serial, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    e2e_steps_pass = int(steps_h.to(torch.int64).sum().item())
    ms_e2e = timed_pipelined(step_e2e, args.steps, max(args.warmup, 1))
    clocks = sampler.stop()
    assert int(steps_hs[1].to(torch.int64).sum().item()) == e2e_steps_pass == int(steps_hs[0].to(torch.int64).sum().item())

    # FP64 FMA peak, measured now on this GPU
    import ctypes as C

    pms, pfl = C.c_double(0), C.c_double(0)
    _lib.check(lib.pdeq_fp64_peak_probe(1 << 17, C.byref(pms), C.byref(pfl), None), "fp64 probe")
    fp64_peak_tflops = pfl.value / (pms.value * 1e-3) / 1e12

    # max over ranks / sums over ranks
    if world > 1:
        t = torch.tensor([ms_res, ms_e2e, ms_e2e_serial], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e, ms_e2e_serial = t.tolist()
        c = torch.tensor([steps_pass, attempts_pass, e2e_steps_pass, bad], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        steps_all, attempts_all, e2e_steps_all, bad = c.tolist()
    else:
        steps_all, attempts_all, e2e_steps_all = steps_pass, attempts_pass, e2e_steps_pass

    if rank == 0:
        n = NU + 1
        fl = flops_per_attempt(n, D)
        kernel_s = ms_res * 1e-3
        achieved_tflops = fl * attempts_pass / kernel_s / 1e12
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if peaks_path.exists() else 6650.0
        bytes_algo = B * 8.0 * (n * D + 4 + (2 * (1 + n * D + n * n + 1)) ) + B * 4.0 * (2 + 2)
        line = {
            "metric": METRIC, "value": steps_all / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(B, world),
            "accepted_steps_per_pass": steps_all, "attempts_per_pass": attempts_all,
            "rejection_ratio": 1.0 - steps_all / max(attempts_all, 1), "failed_instances": bad,
            "attempts_per_s": attempts_all / (ms_res * 1e-3),
            "e2e": {
                "value": e2e_steps_all / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(params_h.numel() * 8 + u0_h.numel() * 8),
                "d2h_bytes_per_step": int(mean_h.numel() * 8 + steps_h.numel() * 4),
                "how": "K public-API steps on two alternating CUDA streams (double-buffered pinned result buffers): "
                       "each step copies its inputs host->device, runs the Taylor pass and the loop kernel, and reads "
                       "its terminal values and step counts back; device time over the K steps / K",
                "serial": {"value": e2e_steps_all / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial,
                           "how": "the same steps one after the other on one stream, L2 flushed in between"},
            },
            "gpu_launches": args.steps * 1 + 2 * args.steps * 5,
            "gpu_launches_note": "1 loop kernel per step in the resident arm; each of the two e2e arms (serial, pipelined) launches 1 loop kernel + 4 Taylor-pass kernels per step",
            "roofline": {
                "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                "frac": achieved_tflops / fp64_peak_tflops,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, from the ncu --set full
                # capture summarised in profiles/r1e_k1_loop_lv_nu4_iso_ts0.ncu.txt (221.0 MB + 684.1 MB)
                "traffic": 905.1e6 if B == B_DEFAULT else None, "traffic_unit": "bytes per launch (ncu)",
                "kernel": f"k1_loop_kernel<LotkaVolterra,4,isotropic,2,ts0,SPEC={lib.pdeq_k1_spec_choice()}>",
                "flops_per_attempt": fl, "attempts_per_launch": attempts_pass,
                "peak_source": "pdeq_fp64_peak_probe (FP64 FMA, measured in this run; MEASURED_PEAKS.json has no FP64 figure)",
                "hbm": {"algorithmic_bytes_per_launch": bytes_algo, "achieved_gbs": bytes_algo / kernel_s / 1e9,
                        "peak_gbs": hbm_peak, "frac": bytes_algo / kernel_s / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks_path.exists() else "fallback 6.65 TB/s"},
            },
            "clocks": clocks,
            "ms_steps": {"resident": per_step[0], "e2e_serial": per_step[1]},
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_block(args.cpu_sample)
            except Exception as exc:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc!r}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=B_DEFAULT, help="instances per GPU")
    ap.add_argument("--cpu-sample", type=int, default=None, help="instances in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.cpu_sample is None:
        args.cpu_sample = REF_SAMPLE if args.impl == "reference" else CPU_SAMPLE
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
