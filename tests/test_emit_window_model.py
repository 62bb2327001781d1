"""Host-side model of the Huff0 emit kernel's stream window (finitestateentropy_b200/csrc/huf_encode.cu): lanes OR their bit strings
(<= 88 bits, 4 consecutive words, never wrapping: three spill cells past the end take the overhang of the last words) into a circular
window of W words; finished words leave 128 at a time, and the spill cells are folded in when words 0..2 of the next lap are flushed.
Checked against a plain concatenation over random bit strings, including the 16-byte alignment offset of the first word.  CPU only."""
import random

import pytest

W = 512


class Window:
    def __init__(self, a0):
        self.win = [0] * (W + 4)
        self.out = {}
        self.bitpos = 8 * a0
        self.flushed = 0

    def put(self, at, value, nbits):
        assert nbits <= 88
        v = value << (at & 31)
        j0 = (at >> 5) & (W - 1)                            # only the first word's position wraps
        for t in range(4):
            self.win[j0 + t] |= (v >> (32 * t)) & 0xFFFFFFFF
        assert v >> 128 == 0

    def flush(self, everything):
        up_to = (self.bitpos + 31) // 32 if everything else self.bitpos >> 5
        while (self.flushed < up_to) if everything else (self.flushed + 128 <= up_to):
            for lane in range(32):
                j = self.flushed + 4 * lane
                if j >= up_to:
                    continue
                i = j & (W - 1)
                v = self.win[i:i + 4]
                self.win[i:i + 4] = [0, 0, 0, 0]
                if i == 0:
                    for t in range(3):
                        v[t] |= self.win[W + t]; self.win[W + t] = 0
                for t in range(4):
                    assert j + t not in self.out
                    self.out[j + t] = v[t]
            self.flushed += 128


@pytest.mark.parametrize("seed", range(10))
def test_window_equals_plain_concatenation(seed):
    rng = random.Random(seed)
    for a0 in (0, 3, 15):
        for shape in ("max", "min", "random"):
            wdw = Window(a0)
            ref, ref_bits = 0, 8 * a0
            for _step in range(40):                         # a step = two 256-symbol groups under one scan
                lens = [[{"max": 88, "min": 8, "random": rng.randint(8, 88)}[shape] for _ in range(32)] for _ in range(2)]
                vals = [[rng.getrandbits(n) for n in grp] for grp in lens]
                excl = [[sum(grp[:l]) for l in range(32)] for grp in lens]
                sum_a = sum(lens[0])
                for l in range(32):
                    wdw.put(wdw.bitpos + excl[0][l], vals[0][l], lens[0][l])
                    wdw.put(wdw.bitpos + sum_a + excl[1][l], vals[1][l], lens[1][l])
                for grp, vs in zip(lens, vals):
                    for n, v in zip(grp, vs):
                        ref |= v << ref_bits; ref_bits += n
                wdw.bitpos += sum_a + sum(lens[1])
                assert ((wdw.bitpos + 31) >> 5) + 3 - wdw.flushed <= W      # nothing was ORed into a cell that still held an unflushed word
                wdw.flush(False)
                assert (wdw.bitpos >> 5) - wdw.flushed < 128
            wdw.put(wdw.bitpos, 1, 1); ref |= 1 << ref_bits; ref_bits += 1; wdw.bitpos += 1     # end mark
            wdw.flush(True)
            for j in range((ref_bits + 31) // 32):
                assert wdw.out.get(j, 0) == (ref >> (32 * j)) & 0xFFFFFFFF, (a0, shape, j)
