"""Parity of the CUDA Huff0 path against the CPU checker, through the C-ABI (-m gpu)."""
import ctypes as C
import numpy as np
import pytest
import torch

from helpers import load_port, ptr, zoo, rand_size, probagen, is_error
from gpu_common import cpu_compress, cpu_decompress, checker, BLOCK, SLOT
import finitestateentropy_b200 as fb

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8) if a.dtype != np.uint64 else a.view(np.int64)).cuda()


def _decode_and_check(data, block, cbuf, cs, slot, orig_for_raw=True):
    d_c = _dev(cbuf); d_s = _dev(cs)
    d_o = _dev(data) if orig_for_raw else None
    out, res = fb.huf_decompress_batch(d_c, d_s, len(data), block, slot, orig=d_o)
    torch.cuda.synchronize()
    res = res.cpu().numpy().view(np.uint64)
    out = out.cpu().numpy()
    nb = (len(data) + block - 1) // block
    # expected verdicts from the CPU checker (a 1-bit code at tableLog 12 is undecodable for the reference itself)
    want_out, want = cpu_decompress("huf", cbuf, cs, data, block=block, slot=slot)
    for b in range(nb):
        n = min(block, len(data) - b * block)
        assert res[b] == want[b], (b, int(res[b]), int(want[b]), n, int(cs[b]))
        if not is_error(int(want[b])):
            assert res[b] == n
            assert np.array_equal(out[b * block: b * block + n], data[b * block: b * block + n]), b
    assert sum(1 for b in range(nb) if not is_error(int(want[b]))) > nb // 2


@pytest.mark.parametrize("p", [0.14, 0.20, 0.80, 0.02])
def test_decode_probagen_1mib(p):
    data = probagen(1048575, p)                      # programs/probaGenerator.c:47 size, 32 blocks, last one 32767
    cbuf, cs, slot = cpu_compress("huf", data)
    assert (cs > 1).all()
    _decode_and_check(data, BLOCK, cbuf, cs, slot)


def test_decode_zoo_ragged_blocks():
    """every block shape the fuzzers use, incl. raw (0), RLE (1), tiny and non-multiple-of-4 sizes"""
    rng = np.random.default_rng(11)
    for block in (32768, 4099, 1000, 131072, 12, 77):
        parts = [zoo(rng, block) for _ in range(int(rng.integers(3, 70)))]
        parts.append(zoo(rng, int(rng.integers(1, block + 1))))
        data = np.concatenate(parts)
        cbuf, cs, slot = cpu_compress("huf", data, block=block)
        _decode_and_check(data, block, cbuf, cs, slot)


def test_decode_error_verdicts():
    """truncated / bit-flipped blocks: same verdict as the CPU single-symbol decoder, nothing written past the block"""
    lib, isref = checker()
    dec = (lambda o, n, c, cs: lib.HUF_decompress4X1(o, n, c, cs)) if isref else (lambda o, n, c, cs: lib.orc_huf_decompress(o, n, c, cs))
    rng = np.random.default_rng(12)
    block = 8192
    data = np.concatenate([zoo(rng, block) for _ in range(160)])
    cbuf, cs, slot = cpu_compress("huf", data, block=block)
    nb = len(cs)
    want = np.zeros(nb, np.uint64)
    for b in range(nb):
        if cs[b] < 2:
            want[b] = block
            continue
        c = cbuf[b * slot: b * slot + int(cs[b])]
        mode = int(rng.integers(0, 3))
        if mode == 0:
            cs[b] = int(rng.integers(2, int(cs[b])))
        elif mode == 1:
            for _ in range(int(rng.integers(1, 4))):
                c[int(rng.integers(0, len(c)))] ^= int(rng.integers(1, 256))
        tmp = np.concatenate([cbuf[b * slot: b * slot + int(cs[b])], np.zeros(32, np.uint8)])
        o = np.zeros(block + 8, np.uint8)
        if cs[b] == block:
            want[b] = block
        else:
            want[b] = dec(ptr(o), block, ptr(tmp), int(cs[b]))
    guard = torch.full((len(data) + 4096,), 0x5A, dtype=torch.uint8, device="cuda")
    out, res = fb.huf_decompress_batch(_dev(cbuf), _dev(cs), len(data), block, slot, out=guard, orig=_dev(data))
    torch.cuda.synchronize()
    res = res.cpu().numpy().view(np.uint64)
    bad = [(b, int(res[b]), int(want[b])) for b in range(nb) if is_error(int(res[b])) != is_error(int(want[b])) or (is_error(int(want[b])) and res[b] != want[b])]
    assert not bad, bad[:10]
    assert (guard[len(data):] == 0x5A).all()
    assert sum(is_error(int(x)) for x in want) > 10


def test_single_block_host_api():
    """the reference-named one-block entry point with host pointers (lib/huf.h:67)"""
    L = fb.lib()
    rng = np.random.default_rng(13)
    for n in (32768, 1000, 70000, 13):
        d = zoo(rng, n) if n != 13 else np.arange(13, dtype=np.uint8) % 3
        cbuf, cs, slot = cpu_compress("huf", d, block=n)
        if cs[0] < 2:
            continue
        out = np.zeros(n, np.uint8)
        r = L.HUF_decompress(ptr(out), n, ptr(cbuf), int(cs[0]))
        assert r == n and np.array_equal(out, d)
