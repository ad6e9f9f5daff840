"""CPU check of bench.py's reference arm (`--impl reference`): the unmodified reference (oracle/_ref) on the host cores
through persistent workers, one JSON line with the contract's keys, and agreement between two runs' throughput (the
round-1 arm timed 40-sweep pool.map calls from outside and swung by 2 x between runs)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(steps, warmup):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_line_and_stability():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_N256_M3.so")):
        pytest.skip("oracle/_ref not built")
    a, b = _run(3, 1), _run(3, 1)
    for d in (a, b):
        assert d["impl"] == "reference" and d["metric"] == "pair_interactions_per_s" and d["unit"] == "pair-interactions/s"
        assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["gpu_launches"] == 0
        cb = d["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sweeps" in cb["sample"]
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["config"]["N"] == 256 and d["config"]["sweeps_per_step"] >= 40
        assert d["ms_per_step"] >= 200.0                    # a step is at least a quarter of a second of work per core
        # 2 N (N-1) pair-interactions per sweep at ~7 ns per reference loop iteration: a few 1e8 per core
        assert 2e7 * cb["cores"] < d["value"] < 2e9 * cb["cores"]
    assert abs(a["value"] - b["value"]) <= 0.25 * max(a["value"], b["value"])
