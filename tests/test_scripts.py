"""The drivers end to end on the GPU (``-m gpu``): each script of scripts/ runs with its reference CLI on a small case
and the JSON / NPZ it emits is checked (the reference's four drivers + the parameter sweep)."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(REPO, "scripts")


def run_script(name, *args, timeout=600):
    proc = subprocess.run([sys.executable, os.path.join(SCRIPTS, name), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=SCRIPTS)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    return proc.stdout


def test_run_mpc(tmp_path, golden):
    out = str(tmp_path)
    run_script("run_mpc.py", "--config", os.path.join(REPO, "configs", "default.json"), "--no-strict", "--output-dir", out,
               "--no-visualization")
    summary = json.load(open(os.path.join(out, "run_mpc.json")))
    data = np.load(os.path.join(out, "run_mpc.npz"))
    g = golden("ref_nonstrict_walk")                     # the reference's own output for this command line
    assert summary["n_steps"] == 420 and summary["qp_solves"] == 838
    assert data["com"].shape == (420, 2) and data["y_hist"].shape == (420, 3, 1)
    assert np.max(np.abs(data["com"] - g["com"])) <= 1e-9 * np.max(np.abs(g["com"]))
    assert 0.03 < summary["max_zmp_y_violation_m"] < 0.045
    # --create-config writes the JSON schema the loader reads back
    cfg_path = os.path.join(out, "my_config.json")
    run_script("run_mpc.py", "--create-config", cfg_path)
    assert "mpc" in json.load(open(cfg_path)) and json.load(open(cfg_path))["mpc"]["horizon"] == 150


def test_run_compare_resistance(tmp_path):
    out = str(tmp_path)
    run_script("run_compare_resistance.py", "--sweep", "64", "--output-dir", out)
    summary = json.load(open(os.path.join(out, "compare_resistance.json")))
    data = np.load(os.path.join(out, "compare_resistance.npz"))
    assert summary["scenarios"] == 64 and summary["qp_solves_per_mode"] == 64 * 838
    assert summary["strict"]["max_zmp_violation_m"] < 1e-9 and summary["strict"]["not_converged"] == 0
    assert summary["non_strict"]["scenarios_with_violation"] > 0          # the unconstrained controller leaves the sole
    assert data["zmp_y_strict"].shape == (64, 420)
    inside = (data["zmp_y_strict"] <= data["z_max"][None, :, 1] + 1e-9) & (data["zmp_y_strict"] >= data["z_min"][None, :, 1] - 1e-9)
    assert inside.all()
    # --sweep 1 --F-ext: the reference's experiment (one push at n // 2)
    run_script("run_compare_resistance.py", "--sweep", "1", "--F-ext", "400", "--output-dir", out)
    one = np.load(os.path.join(out, "compare_resistance.npz"))
    assert one["force_tick"].tolist() == [210] and one["F_ext"].tolist() == [400.0]


def test_run_compare_runtime(tmp_path):
    out = str(tmp_path)
    stdout = run_script("run_compare_runtime.py", "--horizons", "20", "150", "300", "--batch", "512", "--output-dir", out)
    assert "Runtime Comparison" in stdout
    report = json.load(open(os.path.join(out, "runtime_comparison.json")))
    rows = {r["horizon"]: r for r in report["rows"]}
    assert sorted(rows) == [20, 150, 300]
    for N in (20, 150, 300):
        for leg in ("strict", "non_strict"):
            assert rows[N][leg]["batch_ms"] > 0 and rows[N][leg]["api_ms"] > 0 and rows[N][leg]["not_converged"] == 0
        assert rows[N]["non_strict"]["cpu_ms"] > 0 and rows[N]["non_strict"]["cpu_as_reference_ms"] > rows[N]["non_strict"]["cpu_ms"]
        assert "cpu_note" in rows[N]["strict"]                            # cvxpy / OSQP are not installed
    assert rows[150]["herdt"]["batch_ms"] > 0 and rows[150]["herdt"]["api_ms"] > 0 and rows[150]["herdt"]["not_converged"] == 0
    assert "skipped" in rows[300]["herdt"]
    data = np.load(os.path.join(out, "runtime_comparison.npz"))
    assert data["horizon"].tolist() == [20, 150, 300] and np.isnan(data["herdt_batch_ms"][2])


def test_run_herdt_method(tmp_path):
    out = str(tmp_path)
    run_script("run_herdt_method.py", "--output-dir", out)                  # the reference's single walk (class defaults)
    one = json.load(open(os.path.join(out, "herdt.json")))
    data = np.load(os.path.join(out, "herdt.npz"))
    assert one["n_steps"] == 302 and data["com"].shape == (302, 2) and data["foot_hist"].shape == (302, 2)
    assert 0.3 < one["final_com"][0] < 1.0                               # 0.3 m/s reference speed over ~2 s of walking
    run_script("run_herdt_method.py", "--plans", "96", "--output-dir", out)  # randomized plans, grouped by timeline
    plans = json.load(open(os.path.join(out, "herdt_plans.json")))
    assert plans["plans"] == 96 and plans["groups"] >= 3 and plans["joint_qp_solves"] > 96 * 250
    assert np.load(os.path.join(out, "herdt_plans.npz"))["final_com"].shape == (96, 2)


def test_run_param_sweep(tmp_path):
    out = str(tmp_path)
    stdout = run_script("run_param_sweep.py", "--scenarios", "2048", "--precision", "fp64", "--output-dir", out)
    summary = json.loads(stdout.strip().splitlines()[-1])
    assert summary["scenarios"] == 2048 and summary["gpus"] == 1 and summary["not_converged"] == 0
    assert summary["max_zmp_violation_m"] < 1e-9 and 0 < summary["fallen"] < 2048
    data = np.load(os.path.join(out, "param_sweep.npz"))
    assert sorted(data["scenario_id"].tolist()) == list(range(2048)) and data["fell"].sum() == summary["fallen"]
