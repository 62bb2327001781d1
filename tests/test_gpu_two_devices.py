"""One process, two GPUs (-m gpu; skipped on a 1-GPU box): per-device launcher state (the opt-in shared-memory
attribute, SM count, tier-2 workspace and host pipeline used to be process-wide singletons bound to the first device)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import probagen, ptr
from gpu_common import cpu_compress, BLOCK, SLOT
import finitestateentropy_b200 as fb

pytestmark = pytest.mark.gpu


def test_both_devices_from_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = fb.lib()
    L.HUF_compress2.restype = C.c_size_t; L.HUF_compress2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
    L.FSE_compress2.restype = C.c_size_t; L.FSE_compress2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
    data = probagen(64 * BLOCK + 777, 0.14)
    for codec, enc, dec, one in (("huf", fb.huf_compress_batch, fb.huf_decompress_batch, L.HUF_compress2),
                                 ("fse", fb.fse_compress_batch, fb.fse_decompress_batch, L.FSE_compress2)):
        want_c, want_cs, _ = cpu_compress(codec, data, slot=SLOT)
        for d in (1, 0, 1):                                         # the second device first: nothing may be bound to device 0
            with torch.cuda.device(d):
                src = torch.from_numpy(data).cuda(d)
                cbuf, cs = enc(src, BLOCK, SLOT, 255, 12)
                out, res = dec(cbuf, cs, len(data), BLOCK, SLOT, orig=src)
                torch.cuda.synchronize(d)
                assert np.array_equal(cs.cpu().numpy().view(np.uint64), want_cs), (codec, d)
                assert torch.equal(out, src), (codec, d)
                blk = np.ascontiguousarray(data[:BLOCK]); o = np.zeros(SLOT, np.uint8)
                r = one(ptr(o), SLOT, ptr(blk), BLOCK, 255, 12)      # tier 2 (host pointers) on the current device
                assert r == want_cs[0] and np.array_equal(o[:r], want_c[:r]), (codec, d)
