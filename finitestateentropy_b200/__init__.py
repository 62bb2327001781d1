"""finitestateentropy_b200 -- B200-native (sm_100a) block entropy coding: the FSE / Huff0 32 KB block
hot path of Cyan4973/FiniteStateEntropy behind the reference's own C API.

The product is `libfse_b200.so` (hand-written CUDA kernels + a C-ABI, see include/fse_b200.h).  This
package is only the thin Python binding used by tests/ and bench.py: ctypes over the C-ABI, torch for
device memory / streams / torch.distributed.  There is NO CPU implementation here: if the CUDA
library is missing, importing `lib()` raises."""
import ctypes as C
import os

from . import _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_sz = C.c_size_t
c_vp = C.c_void_p

BLOCK_SIZE = 32768                     # programs/bench.c:98
ERR_MAXCODE = 9                        # lib/error_public.h:55
ERROR_NAMES = {1: "GENERIC", 2: "dstSize_tooSmall", 3: "srcSize_wrong", 4: "corruption_detected",
               5: "tableLog_tooLarge", 6: "maxSymbolValue_tooLarge", 7: "maxSymbolValue_tooSmall",
               8: "workSpace_tooSmall"}


def compress_bound(n):
    """FSE_compressBound (lib/fse.h:290-292): per-block slot size used by programs/bench.c:355,514"""
    return 512 + n + (n >> 7) + 4 + 8


def is_error(code):
    return int(code) > (1 << 64) - ERR_MAXCODE


def error_code(code):
    return ((1 << 64) - int(code)) if is_error(code) else 0


def lib():
    """Loads (building in-tree first if needed) the CUDA library.  Fails loudly; never falls back."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if _build.lib_is_stale():
            try:
                path = _build.build_lib()
            except Exception as exc:  # nvcc missing on the box: use the prebuilt .so if there is one
                if not os.path.exists(path):
                    raise RuntimeError("libfse_b200.so is missing and cannot be built: %r -- there is no CPU fallback" % (exc,))
        L = C.CDLL(path)
        _declare(L)
        _LIB = L
    return _LIB


def _declare(L):
    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("FSEB200_batch_blocks", c_sz, c_sz, c_sz)
    sig("FSEB200_HUF_decompress_batch", c_sz, c_vp, c_sz, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp)
    sig("HUF_decompress", c_sz, c_vp, c_sz, c_vp, c_sz)
    for name in ("FSEB200_HUF_compress_batch", "FSEB200_FSE_compress_batch", "FSEB200_FSE_decompress_batch",
                 "FSEB200_FSEU16_compress_batch", "FSEB200_FSEU16_decompress_batch"):
        if hasattr(L, name):
            if "decompress" in name:
                sig(name, c_sz, c_vp, c_sz, c_sz, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp)
            else:
                sig(name, c_sz, c_vp, c_sz, c_vp, c_vp, c_sz, c_sz, C.c_uint, C.c_uint, c_vp)


from .batch import (huf_decompress_batch, huf_compress_batch, fse_compress_batch, fse_decompress_batch,  # noqa: E402,F401
                    fseu16_compress_batch, fseu16_decompress_batch, nblocks)
