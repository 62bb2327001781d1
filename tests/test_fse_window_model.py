"""Host-side model of the FSE decoder's check-free bit window (finitestateentropy_b200/csrc/fse_codec.cu, fse_decode_cta_kernel):
a 64-bit shift register with `avail` valid bits, refilled with one 32-bit word every TWO symbols when avail <= 32.  Checked here
against a plain big-integer bit reader over random streams: the same bits come out, the register never underflows, and the chunk
accounting (bits retired = 32 * refills + avail_before - avail_after) is exact.  CPU only."""
import random

import pytest


def window_decode(words, k, nbits):
    """words: stream words, words[0] the highest; k: bit offset into words[0] (0..31); nbits: bits wanted per symbol (pairs)."""
    hi = ((words[0] << 32 | words[1]) << k >> 32) & 0xFFFFFFFF
    lo = (words[1] << k) & 0xFFFFFFFF
    avail, q, nxt = 64 - k, words[2], 3
    avail0, refills, out = avail, 0, []
    for i in range(0, len(nbits), 2):
        for nb in nbits[i:i + 2]:
            assert nb <= avail
            out.append(hi >> (32 - nb) if nb else 0)
            hi = ((hi << nb) | (lo >> (32 - nb) if nb else 0)) & 0xFFFFFFFF
            lo = (lo << nb) & 0xFFFFFFFF
            avail -= nb
        if avail <= 32:                                     # the predicated refill
            assert lo == 0
            hi |= q >> avail if avail < 32 else 0
            lo = (q << (32 - avail)) & 0xFFFFFFFF if avail else 0
            avail += 32
            q = words[nxt]; nxt += 1; refills += 1
        assert 32 < avail <= 64
    return out, 32 * refills + avail0 - avail


@pytest.mark.parametrize("max_nb", [12, 13])                # FSE_MAX_TABLELOG, the U16 variant's limit
@pytest.mark.parametrize("seed", range(8))
def test_shift_register_window_reads_the_same_bits(seed, max_nb):
    rng = random.Random(seed * 31 + max_nb)
    nsym = 4096
    for shape in ("random", "max", "zeros_and_max"):
        nbits = [{"random": rng.randint(0, max_nb), "max": max_nb, "zeros_and_max": rng.choice((0, 0, max_nb))}[shape] for _ in range(nsym)]
        total = sum(nbits)
        words = [rng.getrandbits(32) for _ in range(total // 32 + 8)]
        for k in (0, 5, 31):
            big = 0
            for w in words:
                big = big << 32 | w
            nb_total = 32 * len(words)
            pos, want = k, []
            for nb in nbits:
                want.append((big >> (nb_total - pos - nb)) & ((1 << nb) - 1) if nb else 0)
                pos += nb
            got, retired = window_decode(words, k, nbits)
            assert got == want
            assert retired == total
