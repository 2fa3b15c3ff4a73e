"""CPU: `bench.py --impl reference` (the reference arm the driver runs beside ours) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                          "--warmup", "1", "--no-encode"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "hamming_retrieval_query_x_gallery_pairs_per_sec"
    assert line["unit"] == "pairs/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["workload"] == "C1" and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
