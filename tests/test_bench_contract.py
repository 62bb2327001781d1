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


def _check_b200_line(d, launches_per_step):
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] in ("u8", "u16") and d["bit_exact"] is True and d["roundtrip_ok"] is True
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 2.0      # both kernels of the Huff0 encode pair read the source: <= 2x
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 2 ** 30
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "reference"
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not d["clocks"]["reasons"]
    assert d["gpu_launches"] == launches_per_step * d["steps"]
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]


def test_committed_b200_lines_have_the_contract_keys():
    """round 1's headline line, and round 2's lines for BASELINE configs[1], [2] and [4] (all blocks compared with the reference)"""
    _check_b200_line(json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1.json"))), 3)
    for name, launches in (("r02_bench_huf_p14_n1.json", 5), ("r02_bench_fse_p80_n1.json", 2), ("r02_bench_u16_p50_n1.json", 2)):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        _check_b200_line(d, launches)
        assert d["bit_exact_detail"]["blocks_compared"] == d["bit_exact_detail"]["of_blocks"] == d["config"]["blocks_per_gpu"]
        assert d["bit_exact_detail"]["gpu_decodes_checker_output"] is True


def test_committed_multi_gpu_lines_scale_and_carry_the_scatter_gather_object():
    """round 2's 2-, 4- and 8-GPU lines: weak scaling of the device-timed metric (no data-path collective), the configs[3] object"""
    one = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_huf_p14_n1.json")))
    for n in (2, 4, 8):
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_huf_p14_n%d.json" % n)))
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["roundtrip_ok"] is True and not d["clocks"]["reasons"]
        assert 0.95 * n * one["value"] < d["value"] < 1.05 * n * one["value"]
        assert d["e2e"]["roundtrip_ok"] is True and d["e2e"]["value"] > one["e2e"]["value"]
        sg = d["scatter_gather"]
        assert sg["roundtrip_ok"] is True and sg["pieces"] in (2, 4, 8) and 0 < sg["frac_of_link_bound"] <= 1.0
        assert sg["bytes_gathered"] == (n - 1) * 2 ** 30 and sg["bytes_scattered"] < 0.6 * sg["bytes_gathered"]    # only used bytes travel


def test_docs_quote_the_committed_lines():
    """README / DESIGN figures are the committed JSON lines' (a stale table is a wrong claim)"""
    readme = open(os.path.join(ROOT, "README.md")).read()
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    for name in ("r02_bench_huf_p14_n1.json", "r02_bench_fse_p80_n1.json", "r02_bench_u16_p50_n1.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert "%.1f" % d["value"] in readme, (name, d["value"])
        assert str(d["value"]) in design, (name, d["value"])
    for n in (2, 4, 8):
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_huf_p14_n%d.json" % n)))
        assert "%.1f" % d["value"] in readme and str(d["value"]) in design
