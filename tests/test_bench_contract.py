"""bench.py's JSON contract, checked on the arm that runs without a GPU (`--impl reference`, tiny sizes)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "1500", "--landmarks", "60",
         "--dims", "6", "--cpu-sample", "1500", "--steps", "1", "--warmup", "0"],
        check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True
    assert d["metric"] == "cells/sec fit_predict" and d["unit"] == "cells/s" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0


import pytest


@pytest.mark.parametrize("config", ["headline", "2", "3", "4", "5"])
def test_cuda_arm_host_logic_on_the_test_double(config, monkeypatch, capsys):
    """bench.py's own arm, every --config, at toy sizes on the NumPy test double of the C ABI: the host logic of the
    bench (inputs per configuration, timing brackets, parity legs, JSON line) without a GPU."""
    import mellon_b200 as mb
    from fake_lib import FakeBackend

    sys.path.insert(0, ROOT)
    import bench

    argv = ["bench.py", "--config", config, "--cells", "1200", "--landmarks", "50", "--dims", "5", "--cpu-sample", "600",
            "--steps", "1", "--warmup", "1", "--no-clocks"]
    if config == "3":
        argv += ["--rank", "20"]
    if config in ("headline", "5"):
        argv += ["--predict-queries", "700"]
    monkeypatch.setattr(sys, "argv", argv)
    mb.set_backend(FakeBackend())
    try:
        bench.main()
    finally:
        mb.set_backend(None)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "ms_per_step", "config", "e2e", "gpu_launches", "roofline", "parity",
                "log_density_sha256"):
        assert key in d, key
    assert d["config"]["baseline_config"] == config and d["value"] > 0
    assert d["parity"] is not None and d["parity"]["ok"] in (True, False)
    if config == "5":
        assert d["metric"] == "queries/sec predict" and d["parity"]["rel_std_err"] < 1e-9
    else:
        assert d["cpu_baseline"]["kind"] == "port" and d["parity"]["rel_std_err_log_density"] < 1e-5
