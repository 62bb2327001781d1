"""The bench.py JSON contract: keys of the reference arm (run here, CPU only, a few MiB) and of the committed B200 line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mib", "8", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]


def test_committed_b200_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json")))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "u8" and d["bit_exact"] is True and d["roundtrip_ok"] is True
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] and 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 2.0      # both kernels of the pair read the source: <= 2x
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 2 ** 30
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not d["clocks"]["reasons"]
    assert d["gpu_launches"] == 3 * d["steps"]
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
