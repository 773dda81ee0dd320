"""bench.py output contract (driver-facing JSON line): checked on the committed B200 run of the default command and on
a live run of the CPU reference arm (`--impl reference`, the only bench leg that needs no GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly ONE JSON line"
    return json.loads(lines[0])


def test_committed_b200_line_has_the_contract_fields():
    d = _line(open(os.path.join(ROOT, "profiles", "r1_bench_1gpu.log")).read())
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["unit"] == "clips/s" and "workload" in d["config"]
    assert abs(d["value"] - 72 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 1e9 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 1000
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_runs_on_cpu_and_prints_the_same_schema():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = _line(p.stdout)
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # kind: "reference" where the unmodified reference is importable (this container), "port" (oracle) on the GPU box
    want = "reference" if os.path.exists("/root/reference/models/adamml.py") else "port"
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == want
    mine = _line(open(os.path.join(ROOT, "profiles", "r1_bench_1gpu.log")).read())
    assert d["metric"] == mine["metric"] and d["config"]["workload"] == mine["config"]["workload"]
