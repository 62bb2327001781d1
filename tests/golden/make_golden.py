#!/usr/bin/env python
"""Regenerates tests/golden/*.json|npz from the UNMODIFIED reference library compiled by oracle/Makefile
(oracle/_ref/libfse_ref.so <- /root/reference/lib/*.c).  Run in the build container:

    python tests/golden/make_golden.py

The reference ships no golden compressed vectors of its own, so these files
are the committed ground truth used by the CPU suite (oracle port) and by the -m gpu suite."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import load_ref, load_port, ptr, probagen, gen_u16, zoo, is_error  # noqa: E402

BLOCK = 32768
SLOT = 512 + BLOCK + (BLOCK >> 7) + 4 + 8        # FSE_compressBound(32768) = 33548, programs/bench.c:355,514


def blocks_of(data, bs=BLOCK):
    return [data[i:i + bs] for i in range(0, len(data), bs)]


def kat(ref, port, name, data, codec, msv, tl):
    """known-answer record for one bench-style run: per-block return values + XXH64 of the concatenation"""
    sizes, cat = [], bytearray()
    for b in blocks_of(data):
        b = np.ascontiguousarray(b)
        dst = np.zeros(SLOT, np.uint8)
        if codec == "fse":
            r = ref.FSE_compress2(ptr(dst), SLOT, ptr(b), len(b), msv, tl)
        elif codec == "huf":
            r = ref.HUF_compress2(ptr(dst), SLOT, ptr(b), len(b), msv, tl)
        else:
            r = ref.FSE_compressU16(ptr(dst), 32768, ptr(b), len(b) // 2, msv, tl)
        assert not is_error(r)
        sizes.append(int(r))
        if r > 1:
            cat += bytes(dst[:r])
    cat = np.frombuffer(bytes(cat), np.uint8)
    return {"name": name, "codec": codec, "maxSymbolValue": msv, "tableLog": tl, "srcBytes": int(len(data)),
            "srcMd5": hashlib.md5(data.tobytes()).hexdigest(), "cSizes": sizes,
            "total": int(sum(s if s else len(b) for s, b in zip(sizes, blocks_of(data)))),
            "xxh64": "%016x" % port.orc_xxh64(ptr(cat), len(cat), 0)}


def main():
    ref, port = load_ref(), load_port()
    assert ref is not None, "needs /root/reference (oracle/Makefile ref target)"
    kats = []
    for pct in (20, 14, 80, 2):
        data = probagen(1048575, pct / 100.0)      # programs/probaGenerator.c:47 : (1 MB) - 1
        for codec in ("fse", "huf"):
            kats.append(kat(ref, port, "proba%02d" % pct, data, codec, 255, 12))
    u16 = gen_u16(524288, 240, 0.50, 1).view(np.uint8)
    kats.append(kat(ref, port, "u16_p50", u16, "u16", 0, 12))
    json.dump(kats, open(os.path.join(HERE, "kat_bench.json"), "w"), indent=0)

    # small full-byte vectors: (input, compressed bytes, return value) for ragged / edge sizes
    rng = np.random.default_rng(20260922)
    vec = {}
    k = 0
    sizes = [0, 1, 2, 3, 4, 7, 11, 12, 13, 15, 16, 17, 31, 64, 100, 255, 256, 1000, 1001, 1499, 1500, 4096, 8191, 32765, 32766, 32767, 32768, 65536, 131072]
    for n in sizes + [int(rng.integers(20, 40000)) for _ in range(40)]:
        d = np.ascontiguousarray(zoo(rng, n))
        for codec, fn in (("fse", ref.FSE_compress2), ("huf", ref.HUF_compress2)):
            cap = 512 + n + (n >> 7) + 12
            dst = np.zeros(cap + 8, np.uint8)
            r = fn(ptr(dst), cap, ptr(d), n, 255, 12)
            vec["in_%d" % k] = d
            vec["codec_%d" % k] = np.array([0 if codec == "fse" else 1])
            vec["ret_%d" % k] = np.array([r], dtype=np.uint64)
            vec["out_%d" % k] = dst[:r].copy() if (not is_error(r) and r > 1) else (dst[:1].copy() if r == 1 and codec == "huf" else np.zeros(0, np.uint8))
            k += 1
    for n in (2, 3, 16, 1000, 16384, 16383, 40000):
        d = np.ascontiguousarray(gen_u16(n + 50, 240, float(rng.uniform(0.05, 0.8)), int(rng.integers(1, 1 << 30)))[50:])
        dst = np.zeros(2 * n + 600, np.uint8)
        r = ref.FSE_compressU16(ptr(dst), 2 * n + 592, ptr(d), n, 0, 12)
        vec["in_%d" % k] = d.view(np.uint8)
        vec["codec_%d" % k] = np.array([2])
        vec["ret_%d" % k] = np.array([r], dtype=np.uint64)
        vec["out_%d" % k] = dst[:r].copy() if (not is_error(r) and r > 1) else np.zeros(0, np.uint8)
        k += 1
    vec["count"] = np.array([k])
    np.savez_compressed(os.path.join(HERE, "vectors_small.npz"), **vec)
    print("wrote", len(kats), "KAT records and", k, "small vectors")


if __name__ == "__main__":
    main()
