"""Batched, device-resident entry points on torch CUDA tensors (thin wrappers over FSEB200_*_batch).

Geometry is the one programs/bench.c:530-548 builds: a flat uncompressed buffer split into
`block_size` blocks (last one shorter), compressed block b in the fixed slot `cbuf[b*slot : (b+1)*slot]`,
`csizes[b]` = the reference's return value for that block (0 = stored raw, 1 = RLE, error codes in-band)."""
import torch


def nblocks(total, block_size):
    return (total + block_size - 1) // block_size


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _check(t, dtype):
    assert t.is_cuda and t.is_contiguous() and t.dtype == dtype, (t.device, t.dtype, t.is_contiguous())


def _ret(code, what):
    from . import is_error, error_code, ERROR_NAMES
    if is_error(code):
        raise RuntimeError("%s failed: %s" % (what, ERROR_NAMES.get(error_code(code), code)))


def _compress(fn_name, src, block_size, slot, msv, tlog, cbuf=None, csizes=None):
    from . import lib
    _check(src, torch.uint8)
    nb = nblocks(src.numel(), block_size)
    if cbuf is None:
        cbuf = torch.empty(nb * slot + 64, dtype=torch.uint8, device=src.device)
    if csizes is None:
        csizes = torch.empty(nb, dtype=torch.int64, device=src.device)
    _check(cbuf, torch.uint8); _check(csizes, torch.int64)
    assert cbuf.numel() >= nb * slot and csizes.numel() >= nb
    r = getattr(lib(), fn_name)(cbuf.data_ptr(), slot, csizes.data_ptr(), src.data_ptr(), src.numel(), block_size,
                                msv, tlog, _stream_ptr())
    _ret(r, fn_name)
    return cbuf, csizes


def _decompress(fn_name, cbuf, csizes, total, block_size, slot, out=None, results=None, orig=None):
    from . import lib
    _check(cbuf, torch.uint8); _check(csizes, torch.int64)
    nb = nblocks(total, block_size)
    if out is None:
        out = torch.empty(total, dtype=torch.uint8, device=cbuf.device)
    if results is None:
        results = torch.empty(nb, dtype=torch.int64, device=cbuf.device)
    _check(out, torch.uint8); _check(results, torch.int64)
    assert out.numel() >= total and results.numel() >= nb and csizes.numel() >= nb
    r = getattr(lib(), fn_name)(out.data_ptr(), total, block_size, cbuf.data_ptr(), slot, csizes.data_ptr(),
                                results.data_ptr(), orig.data_ptr() if orig is not None else None, _stream_ptr())
    _ret(r, fn_name)
    return out, results


def huf_compress_batch(src, block_size=32768, slot=None, max_symbol_value=255, table_log=12, cbuf=None, csizes=None):
    from . import compress_bound
    return _compress("FSEB200_HUF_compress_batch", src, block_size, slot or compress_bound(block_size),
                     max_symbol_value, table_log, cbuf, csizes)


def huf_decompress_batch(cbuf, csizes, total, block_size=32768, slot=None, out=None, results=None, orig=None):
    from . import compress_bound
    return _decompress("FSEB200_HUF_decompress_batch", cbuf, csizes, total, block_size,
                       slot or compress_bound(block_size), out, results, orig)


def fse_compress_batch(src, block_size=32768, slot=None, max_symbol_value=255, table_log=12, cbuf=None, csizes=None):
    from . import compress_bound
    return _compress("FSEB200_FSE_compress_batch", src, block_size, slot or compress_bound(block_size),
                     max_symbol_value, table_log, cbuf, csizes)


def fse_decompress_batch(cbuf, csizes, total, block_size=32768, slot=None, out=None, results=None, orig=None):
    from . import compress_bound
    return _decompress("FSEB200_FSE_decompress_batch", cbuf, csizes, total, block_size,
                       slot or compress_bound(block_size), out, results, orig)


def fseu16_compress_batch(src, block_size=32768, slot=32768, max_symbol_value=0, table_log=12, cbuf=None, csizes=None):
    """src: uint8 view of little-endian uint16 symbols; block_size in BYTES (programs/bench.c:221 halves it)"""
    return _compress("FSEB200_FSEU16_compress_batch", src, block_size, slot, max_symbol_value, table_log, cbuf, csizes)


def fseu16_decompress_batch(cbuf, csizes, total, block_size=32768, slot=32768, out=None, results=None, orig=None):
    return _decompress("FSEB200_FSEU16_decompress_batch", cbuf, csizes, total, block_size, slot, out, results, orig)
