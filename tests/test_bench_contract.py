"""The bench.py contract that does not need a GPU: the reference arm prints one JSON line
with the agreed keys, and the committed GPU lines under profiles/ carry the roofline,
cpu_baseline, e2e, launch-count and clock fields the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].startswith("gates/sec") and baseline["metric"].startswith("gates/sec")   # BASELINE's headline
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["higher_is_better"] is True and "workload" in d["config"]


def test_committed_gpu_lines_have_the_contract_fields():
    d = json.loads(open(os.path.join(ROOT, "profiles", "r1_bench_default_n1.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(d) and d["n_gpus"] == 1 and d["dtype"] == "c128"
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["traffic"] is not None and abs(roof["traffic"] / roof["bytes_per_launch"] - 1.0) < 0.02
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 16 << 30 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
    for name in ("r1_sharded_31q_2gpu.json", "r1_sharded_32q_4gpu.json", "r1_sharded_33q_8gpu.json"):
        s = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert BASE_KEYS <= set(s) and s["scaling"] == "weak" and s["gpu_launches"] > 0
        assert abs(s["config"]["final_norm"] - 1.0) < 1e-9
