"""bench.py's output contract, checked on CPU: the reference arm prints one JSON line with the
required keys; the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                          capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_line():
    out = _run("--impl", "reference", "--config", "cfg2", "--steps", "5", "--warmup", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "simplex_pivots_per_sec"
    assert d["unit"] == "pivots/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 5 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "pivots/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = _run("--impl", "reference", "--config", "cfg2", "--steps", "3", env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = _run("--config", "cfg2", "--steps", "3")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.gpu
def test_product_arm_line_with_parity_block_on_config_2():
    """The product arm on a real GPU (config 2, seconds): the contract keys, a roofline record, and
    the parity block -- oracle prefix + the committed full-solve fixture -- all true."""
    out = _run("--config", "cfg2", "--steps", "50", "--warmup", "3", "--no-cfg4", "--no-cpu-baseline")
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "simplex_pivots_per_sec" and d["value"] > 0 and d["n_gpus"] == 1
    assert set(d["config"]) == {"workload", "m", "n", "R", "C"}
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["achieved"] > 0 and d["gpu_launches"] >= 1
    assert d["e2e"]["value"] > 0 and d["e2e"]["pivots"] == 696 and d["e2e"]["h2d_bytes_per_step"] > 0
    par = d["parity"]
    assert par["final_fixture"] == "tests/golden/cfg2_final.npz" and par["prefix_pivots"] == 10
    for key in ("prefix_trace_equal", "prefix_rhs_equal", "prefix_obj_row_equal", "prefix_basis_equal",
                "final_pivots_equal", "final_basis_equal", "final_rhs_equal", "final_obj_row_equal",
                "final_trace_equal"):
        assert par[key] is True, (key, par)
    assert par["objective"] == par["fixture_objective"]


@pytest.mark.gpu
def test_product_arm_reads_an_mps_file():
    """f4: bench.py --mps routes read_mps -> build_tableau -> b200lp_solve."""
    mps = os.path.join(ROOT, "tests", "golden", "mps", "simple-problem.mps")
    out = _run("--mps", mps, "--steps", "2", "--warmup", "3", "--no-cfg4", "--no-parity")
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][0])
    assert d["data"] == "file" and d["mps"]["file"] == "simple-problem.mps" and d["mps"]["rows"] >= 2
    assert d["e2e"]["status"] == 0
