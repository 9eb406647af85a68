"""CPU: the reference arm of bench.py (`--impl reference`) keeps the driver's JSON contract: exactly one line on
stdout, the base keys, `impl`, a cpu_baseline describing the run and an e2e object with zero copy bytes.
(The GPU arm needs a B200; its line is checked by the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one JSON line, got {len(lines)}"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "agent updates/sec" and d["unit"] == "updates/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
