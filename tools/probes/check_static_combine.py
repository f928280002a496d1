#!/usr/bin/env python
"""Host model of the stream-K schedule (StreamK in bndm_b200/csrc/common.cuh) that checks the assumptions behind
tools/probes/patches/static_combine.patch without a GPU:
  * every row tile that is split over several CTAs has exactly one combiner, first_cta(tile);
  * the combiner's LAST schedule unit lies in that tile (the combine is its final piece of work) and it is found by
    decoding that unit, as the patch does;
  * a CTA combines at most one tile;
  * the other contributors' partial-tile slots are the consecutive slots after the combiner's.
    python tools/probes/check_static_combine.py"""
import itertools
import sys

NPIX, STAGE_K = 4096, 32


class StreamK:
    def __init__(self, n_tiles, dense, n_colblk, num_sms, sub):
        self.n_tiles, self.dense, self.n_colblk, self.sub = n_tiles, dense, n_colblk, sub
        self.Stot = self.cum(n_tiles)
        self.W = n_colblk * self.Stot
        self.G = min(num_sms, self.W)

    def cum(self, i):
        return ((NPIX // STAGE_K) * i if self.dense else 2 * i * (i + 1)) // self.sub

    def cta_begin(self, c):
        return c * self.W // self.G

    def cta_of(self, g):
        return ((g + 1) * self.G - 1) // self.W

    def tile_begin(self, cb, t):
        return cb * self.Stot + self.Stot - self.cum(t + 1)

    def tile_end(self, cb, t):
        return cb * self.Stot + self.Stot - self.cum(t)

    def decode(self, g):
        cb, r = divmod(g, self.Stot)
        a = self.Stot - 1 - r
        i = 0
        while self.cum(i + 1) <= a:
            i += 1
        return cb, i, r - (self.Stot - self.cum(i + 1))

    def slot(self, cta, cb, t):
        return cta + cb * self.n_tiles + (self.n_tiles - 1 - t)

    def first_cta(self, cb, t):
        return self.cta_of(self.tile_begin(cb, t))

    def last_cta(self, cb, t):
        return self.cta_of(self.tile_end(cb, t) - 1)


def check(n_tiles, dense, n_colblk, num_sms, sub):
    k = StreamK(n_tiles, dense, n_colblk, num_sms, sub)
    combiner_of = {}
    for cta in range(k.G):                                   # what the patch computes per CTA
        cb, t, _ = k.decode(k.cta_begin(cta + 1) - 1)
        if k.first_cta(cb, t) == cta and k.last_cta(cb, t) != cta:
            assert cta not in combiner_of
            combiner_of[cta] = (cb, t)
    tiles = set(combiner_of.values())
    assert len(tiles) == len(combiner_of), "two CTAs combine the same tile"
    n_multi = 0
    for cb, t in itertools.product(range(n_colblk), range(n_tiles)):
        c0, c1 = k.first_cta(cb, t), k.last_cta(cb, t)
        if c0 == c1:
            assert (cb, t) not in tiles
            continue
        n_multi += 1
        assert combiner_of.get(c0) == (cb, t), f"tile {(cb, t)}: first_cta {c0} is not its combiner"
        # contributors c0+1 .. c1 own consecutive slots after the combiner's, in ascending-k order
        for s, c in enumerate(range(c0 + 1, c1 + 1)):
            assert k.slot(c, cb, t) == k.slot(c0 + 1, cb, t) + s
            b, e = max(k.cta_begin(c), k.tile_begin(cb, t)), min(k.cta_begin(c + 1), k.tile_end(cb, t))
            assert e > b, "contributor without a unit in the tile"
        # the combiner's piece is the lowest-k piece of the tile and the end of its range
        assert k.cta_begin(c0) <= k.tile_begin(cb, t) < k.cta_begin(c0 + 1) <= k.tile_end(cb, t)
    assert n_multi == len(combiner_of)
    return n_multi, k.G


if __name__ == "__main__":
    for n_tiles, dense, n_colblk, sms, sub in itertools.product((16, 32), (0, 1), (1, 2, 3), (148, 132, 64, 7), (1, 2)):
        n_multi, G = check(n_tiles, dense, n_colblk, sms, sub)
        if sms == 148 and n_colblk == 1:
            print(f"n_tiles={n_tiles} dense={dense} sub={sub}: {G} CTAs, {n_multi} combiners")
    print("static-combine schedule assumptions hold")
    sys.exit(0)
