"""bench.py's reference arm runs on the CPU: check here that it honours --steps/--warmup and prints the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--streams", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["metric"] == "blob_detection_frames_per_sec" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "chains" in cb["sample"]
    assert line["config"]["workload"].startswith("c2:")
