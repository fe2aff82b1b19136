"""bench.py's reference arm runs on CPU and prints the contract's JSON line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "remode_640x480", "--frames", "3",
                          "--steps", "1", "--warmup", "0", "--force-port", "--cpu-rows", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    j = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["vs_baseline"] is None and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert j["config"]["workload"] == "remode_640x480"
    # the reference arm runs none of the product: only the oracle and the input renderer are mapped
    assert j["native_libs_loaded"] is not None and "libdmf.so" not in j["native_libs_loaded"]
    assert "liboracle.so" in j["native_libs_loaded"]


def test_reference_arm_uses_compiled_reference_when_available():
    import oracle
    if oracle.ref_lib() is None:
        import pytest
        pytest.skip("oracle/_ref not built here")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "remode_640x480", "--frames", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert j["cpu_baseline"]["kind"] == "reference"
