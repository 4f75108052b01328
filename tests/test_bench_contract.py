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
