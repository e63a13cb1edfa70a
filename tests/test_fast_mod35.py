"""fast_mod35 (csrc/ckm_common.cuh): key % d for keys < 2^35 and 64 <= d < 2^32 from 32-bit pieces.  The device function is
exercised by every GPU parity test; this restates its arithmetic with Python integers and checks the claims its comment makes
(no intermediate leaves 64 bits, q is floor(key/d) or one less, one correction suffices) over every bucket count the
reference's builder can choose (build_signature_kmers.cc:862-865) and the edges of the domain."""
import numpy as np

from close_kmers_b200 import synth

MAX_ENCODED = 20**8


def fast_mod35(key: int, d: int) -> int:
    m = (1 << 35) // d  # magic35()
    lo, hi = key & 0xFFFFFFFF, key >> 32
    t = hi * m
    assert t < 1 << 32  # (uint32_t)(key >> 32) * m35 does not wrap
    prod = lo * m + (t << 32)
    assert prod < 1 << 64
    q = prod >> 35
    assert q < 1 << 32
    r = key - q * d
    assert 0 <= r < 2 * d  # q is floor(key/d) or one less
    return r - d if r >= d else r


def test_fast_mod35_matches_integer_modulo():
    rng = np.random.default_rng(35)
    ds = [64, 65, 127, 4096, (1 << 32) - 2] + [p for p in synth.PRIMES if 64 <= p < (1 << 32) - 1]
    for d in ds:
        assert (1 << 35) // d <= 1 << 29
        keys = rng.integers(0, MAX_ENCODED, 20_000, dtype=np.uint64).tolist()
        keys += [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, MAX_ENCODED - 1, (MAX_ENCODED // d) * d, (MAX_ENCODED // d) * d - 1]
        for k in keys:
            if 0 <= k < MAX_ENCODED:
                assert fast_mod35(k, d) == k % d, (k, d)
