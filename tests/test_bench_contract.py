"""bench.py contract pieces that need no GPU: the reference arm on a CPU-only box (it falls back to the CPU port
of the reference's algorithm, oracle/gvv_oracle.cpp) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line_without_a_gpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("views/sec fwd+bwd at 1024") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("config2")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and abs(e["value"] - d["value"]) < 1e-6 and e["unit"] == d["unit"]


def test_algorithmic_bytes_follow_the_survey_formulas():
    sys.path.insert(0, ROOT)
    import bench
    sc = {"width": 1024, "height": 1024, "num_vertices": 34970, "faces": [0] * 69936}
    fwd, bwd, per_kernel = bench.algorithmic_bytes(sc)
    P, N, F = 1024 * 1024, 34970, 69936
    assert fwd == P * 24 + N * 36 + F * 12 and bwd == P * 24 + N * 60 + F * 12 + 216
    assert per_kernel["raster_kernel"] * 8 == 214754688                                     # the bench line's algorithmic_bytes_per_launch
