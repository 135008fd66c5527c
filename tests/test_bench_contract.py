"""The bench.py contract: keys of the JSON line.  The reference arm runs here (CPU); the lines of our own arm are the ones
archived from the last GPU run (profiles/r01i_bench_*.json)."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["value"] > 0 and d["ms_per_step"] > 0


@pytest.mark.parametrize("name", ["c2", "c2ml", "c4g_1gpu", "c4g_2gpu", "c5_2gpu"])
def test_archived_lines_of_our_arm(name):
    f = ROOT / "profiles" / f"r01i_bench_{name}.json"
    d = json.loads(f.read_text())
    _check_common(d)
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])


def test_reference_arm_prints_one_contract_line():
    from oracle.oracle import reference_available
    if not reference_available():
        pytest.skip("oracle/_ref is not built on this machine")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
