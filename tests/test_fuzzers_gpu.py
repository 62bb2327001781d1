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
