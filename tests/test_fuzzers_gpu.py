"""the reference's OWN fuzzers, unmodified, behind the product library.

oracle/Makefile (target `fuzzers`) compiles programs/fuzzer.c, fuzzerHuff0.c and fuzzerU16.c (+ programs/xxhash.c) where they
lie under /root/reference and links them against finitestateentropy_b200/libfse_b200.so; the binaries travel to the GPU box
in oracle/_ref/.  Each run = the fuzzer's unit tests (FSE_normalizeCount / NCount corner cases, raw tables, FSE_countU16 ...)
followed by N round-trip + robustness iterations (too-small destinations, truncated and garbage inputs, guard bytes) --
programs/fuzzer.c:142-464, fuzzerHuff0.c:137-261, fuzzerU16.c:145-284.  A failing check makes the program exit non-zero."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("name,iters", [("fuzzer_b200", 200), ("fuzzerHuff0_b200", 200), ("fuzzerU16_b200", 200)])
def test_reference_fuzzer_passes_against_the_gpu_library(name, iters):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("prebuilt %s not present (built by oracle/Makefile where the reference tree is mounted)" % name)
    r = subprocess.run([exe, "-s1", "-i%d" % iters], capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert "tests passed" in tail or "tests completed" in tail.lower(), tail


def test_reference_harness_runs_unmodified_on_the_gpu_library(tmp_path):
    """SURVEY.md section 8b taken literally: the reference's own `fse -b` (programs/commandline.c + bench.c, unmodified) linked
    against libfse_b200.so benchmarks a probagen file -- one block per synchronous call, so slow, but every call lands in the
    CUDA kernels and bench.c's own XXH32 self-check (programs/bench.c:444-454) must pass; then its file mode (fileio.c) round-trips."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import probagen
    exe = os.path.join(BIN, "fse_cli_b200")
    if not os.path.exists(exe):
        pytest.skip("prebuilt fse_cli_b200 not present")
    src = tmp_path / "proba.bin"
    probagen(262143, 0.20).tofile(src)
    for flag in ("-e", "-h"):
        r = subprocess.run([exe, "-b", flag, "-i1", str(src)], capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
        tail = (r.stdout + r.stderr)[-800:]
        assert r.returncode == 0 and "ERROR" not in tail.upper().replace("NO ERROR", ""), tail
        assert "MB/s" in tail, tail
        out = tmp_path / ("p" + flag + ".fse"); back = tmp_path / ("p" + flag + ".out")
        r = subprocess.run([exe, "-f", flag, str(src), str(out)], capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
        assert r.returncode == 0, (r.stdout + r.stderr)[-500:]
        r = subprocess.run([exe, "-f", "-d", str(out), str(back)], capture_output=True, text=True, timeout=600, stdin=subprocess.DEVNULL)
        assert r.returncode == 0, (r.stdout + r.stderr)[-500:]
        assert open(back, "rb").read() == open(src, "rb").read()
