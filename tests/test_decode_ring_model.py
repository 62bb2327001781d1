"""Host-side model of the Huff0 decoder's per-lane stream ring (finitestateentropy_b200/csrc/huf_decode.cu, "stream feeder"):
8 words per lane, topped up by 4-word chunks at two kinds of check, 8 symbols apart.  The kernel recovers the number of unread words
from slot numbers alone, which is only valid if that number stays in [1, 8] at every check and a fetch never finds the ring empty --
the induction in the kernel's comment, checked here over random and adversarial code lengths (CPU only, no GPU code involved)."""
import random

import pytest

MAX_CODE_BITS = 12          # HUF_TABLELOG_ABSOLUTEMAX (lib/huf.h:100)


def run_lane(lengths, r0, unread0):
    """lengths: code length of every symbol (multiple of 32 symbols = whole iterations of the hot loop).
    r0: bit offset into the first window word (0..31); unread0: words in the ring after set-up (the kernel tops up to >= 5)."""
    r, unread = r0, unread0
    it = iter(lengths)
    for _ in range(len(lengths) // 32):
        for h in range(8):                                  # 8 groups of 4 symbols
            if h & 3 == 0:                                  # group-start check: the common store
                assert 1 <= unread <= 8
                if unread <= 4:
                    unread += 4
                assert unread >= 5
            elif h & 3 == 2:                                # mid-group check: only when the lane ran low
                assert 1 <= unread <= 8
                if unread <= 3:
                    unread += 4
            for _pair in range(2):
                bits = next(it) + next(it)
                assert bits <= 2 * MAX_CODE_BITS < 32       # a pair uses up at most one word: one rotation per advance
                old = r
                r += bits
                if (r ^ old) & 32:                          # bit 5 flips exactly when a 32-bit word has been used up
                    assert unread >= 1, "fetch from an empty ring"
                    unread -= 1
                assert (r >> 5) - (old >> 5) in (0, 1)
    return unread


@pytest.mark.parametrize("seed", range(20))
def test_ring_never_runs_dry_and_unread_stays_recoverable(seed):
    rng = random.Random(seed)
    n = 32 * 256                                            # one 8,192-symbol stream
    shapes = [
        [MAX_CODE_BITS] * n,                                # longest codes only: a word every 2.7 symbols
        [1] * n,                                            # shortest
        [rng.randint(1, MAX_CODE_BITS) for _ in range(n)],
        [MAX_CODE_BITS if (i // 7) & 1 else 1 for i in range(n)],
        [rng.choice((1, 1, 1, 11, 12)) for _ in range(n)],
    ]
    for lengths in shapes:
        for r0 in (0, 1, 17, 31):
            for unread0 in (5, 6, 7, 8):
                assert 1 <= run_lane(lengths, r0, unread0) <= 8
