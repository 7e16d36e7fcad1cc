"""CPU-only checks of bench.py's output contract: the reference arm (which times the oracle's C port on the host
cores and therefore runs without a GPU), the algorithmic FLOP model behind `roofline.achieved`, and the refusal of
the product arm to run without a CUDA device (there is no CPU fallback)."""

import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=300)


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "1024")
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("accepted solver steps/sec")
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_flop_model_is_the_survey_formula():
    sys.path.insert(0, str(ROOT))
    import bench

    n, d, c_vf = 5, 2, 10
    expect = 2 * n**3 + 10 * n**3 / 3 + 2 * 4 * (n + 1) ** 3 / 3 + 4 * n * n * d + c_vf  # SURVEY 8(d), isotropic ts0
    assert abs(bench.flops_per_attempt(n, d) - expect) < 1e-9
    assert abs(bench.flops_per_attempt(n, d) - 1452.6666666666667) < 1e-9


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return  # on the GPU box the product arm is exercised by the driver itself
    res = _run("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
