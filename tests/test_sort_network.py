"""The tile list sort's register network (csrc/kernels_span.cuh: warpBitonicSort<E>): 32 * E keys, E consecutive keys per lane,
exchanges at distances below E inside a lane, the others by a lane-xor shuffle. This restates the kernel's index algebra
(kk / j loops, `up`, `lower`, partner lane j / E) lane by lane and checks that it sorts — a guard for whoever edits the network;
the kernel itself is checked on the GPU by every parity test (a list out of submission order breaks the depth dead band)."""
import random

import pytest


def network(keys, per_lane):
    n = 32 * per_lane
    v = [[keys[lane * per_lane + r] for r in range(per_lane)] for lane in range(32)]
    kk = 2
    while kk <= n:
        j = kk >> 1
        while j > 0:
            if j >= per_lane:
                partner = j // per_lane
                nxt = [[0] * per_lane for _ in range(32)]
                for lane in range(32):
                    for r in range(per_lane):
                        other = v[lane ^ partner][r]                  # __shfl_xor_sync(v[r], j / E)
                        i = lane * per_lane + r
                        up, lower = (i & kk) == 0, (i & j) == 0
                        nxt[lane][r] = min(v[lane][r], other) if lower == up else max(v[lane][r], other)
                v = nxt
            else:
                for lane in range(32):
                    for r in range(per_lane):
                        if (r & j) == 0:
                            up = ((lane * per_lane + r) & kk) == 0
                            x, y = v[lane][r], v[lane][r | j]
                            if (x > y) == up:
                                v[lane][r], v[lane][r | j] = y, x
            j >>= 1
        kk <<= 1
    return [v[lane][r] for lane in range(32) for r in range(per_lane)]


@pytest.mark.parametrize("per_lane", [1, 2, 4, 8, 16])
def test_register_network_sorts(per_lane):
    rng = random.Random(100 + per_lane)
    n = 32 * per_lane
    for trial in range(40):
        used = rng.randint(2, n)
        keys = [rng.randrange(1 << 24) for _ in range(used)] + [0xFFFFFFFF] * (n - used)     # the kernel pads with ~0
        if trial % 4 == 0:
            keys[:used] = sorted(keys[:used], reverse=True)
        if trial % 7 == 0:
            keys[:used] = [keys[0]] * used                                                  # ties
        assert network(keys, per_lane) == sorted(keys)
