"""bench.py contract pieces that run without a GPU: the reference arm prints one well-formed JSON line;
the report module formats A-vs-B lines like the reference's benchmark report."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "points/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("fv_tp2d C384x72") and d["gpu_launches"] == 0


def test_non_zero_rank_of_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_report_lines():
    from b200stencil.bench import report

    rep = report.compare("cpu", {"fv_tp2d": [2.0, 2.2, 1.8], "remap": [1.0]}, "b200", {"fv_tp2d": [0.5, 0.4, 0.6]})
    text = str(rep)
    assert "fv_tp2d: 1.00x (2.000000s) - 4.00x (0.500000s)" in text and "remap" not in text
    g = json.dumps({"config": {"workload": "fv_tp2d transport step (x)"}, "value": 100.0, "n_gpus": 1, "dtype": "f64",
                    "e2e": {"value": 2.0}})
    r = json.dumps({"value": 4.0, "cpu_baseline": {"kind": "port", "cores": 16}})
    text = str(report.from_bench_lines(g, r))
    assert "25.00x" in text and "0.50x" in text
