"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm (the reference's own CPU
target through oracle/_ref, or the oracle port when that did not travel) prints exactly one JSON line with the keys
the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "16", "--steps", "3",
                        "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [ln for ln in p.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cg_iterations_per_sec" and d["unit"] == "it/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "poisson3d_16" in d["config"]["workload"]


def test_algorithmic_bytes_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    N, nnz = 1 << 24, 117047296
    assert bench.alg_bytes(N, nnz) == (2948071428, 1740111876)        # SURVEY.md 8(d), BASELINE.md 4
