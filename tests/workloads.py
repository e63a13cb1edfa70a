"""Shared seeded workloads + comparison helpers for the parity tests."""
from __future__ import annotations

import numpy as np

from close_kmers_b200 import api, synth


def small_world(seed=1, n_protos=400, n_sigs=100_000, n_functions=None, otu_mode="mixed", mean_len=300, sd=0.0):
    protos = synth.make_prototypes(seed, n_protos, mean_len, sd)
    sig = synth.make_signatures(protos, n_sigs, n_functions=n_functions, otu_mode=otu_mode)
    nb = synth.bucket_count(len(sig.keys))
    img = api.build_image(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    return protos, sig, img


EDGE_SEQS = [
    b"",                                  # empty
    b"A",                                 # shorter than a k-mer
    b"ACDEFGHI",                          # exactly 8: nothing probed (kguts.cc:792)
    b"ACDEFGHIK",                         # 9: one window
    b"XXXXXXXXXXXXXXXXXXXX",              # all ambiguous
    b"acdefghiklmnpqrstvwy" * 3,          # lowercase is invalid on the query path
    b"ACDEFGHIKLMNPQRSTVWY" * 10,
    b"ACDEFGHIKLMNPQRSTVWYBJOUXZ*-" * 4,
    b"MKV\x00ACDEFGHIKLMNPQRSTVWYACDEFGHIKLMNPQRSTVWY",  # embedded NUL: strlen() ends the scan (kguts.cc:791)
]


def edge_batch(protos, seed=5):
    """Edge cases + prototype-derived sequences around the ambiguity / bound rules."""
    rng = np.random.default_rng(seed)
    seqs = list(EDGE_SEQS)
    aa = synth.AA
    for k in range(40):
        p = int(rng.integers(0, protos.n))
        lo, hi = int(protos.offsets[p]), int(protos.offsets[p + 1])
        s = aa[protos.codes[lo:hi]].copy()
        mode = k % 8
        if mode == 0:
            s = s[: int(rng.integers(1, 20))]            # very short prefixes
        elif mode == 1:
            s[rng.integers(0, len(s), 5)] = ord("X")      # scattered X
        elif mode == 2:
            s[7] = ord("X")                               # ambiguity exactly at the first window's end
        elif mode == 3:
            s[-1] = ord("*")                              # trailing stop
        elif mode == 4:
            s = np.concatenate([s[:50], np.frombuffer(b"X" * 9, np.uint8), s[50:]])
        elif mode == 5:
            q = int(rng.integers(0, protos.n))
            t = aa[protos.codes[int(protos.offsets[q]):int(protos.offsets[q + 1])]]
            s = np.concatenate([s[:120], t[:40], s[120:], t[200:]])  # F1 | F2 | F1 | F2 sandwiches
        elif mode == 6:
            s = np.concatenate([s, s])                    # repeated domain
        else:
            s = s[::-1].copy()                            # reversed: mostly misses
        seqs.append(s.tobytes())
    return synth.batch_from_strings(seqs)


def concat_batches(a, b):
    off = np.concatenate([a.offsets, b.offsets[1:] + a.offsets[-1]])
    return synth.Batch(np.concatenate([a.residues, b.residues]), off.astype(np.uint64))


def assert_results_equal(got: dict, want: dict, what: str, check_ambig_indices: bool = True):
    """Bit-exact comparison of two ckm_batch_out-style dicts (integer fields AND f32 scores)."""
    for key in ("call_offsets", "calls", "hit_offsets", "hits", "otu_offsets", "otus", "best"):
        if key not in want:
            continue
        assert key in got, f"{what}: {key} missing"
        a, b = got[key], want[key]
        assert len(a) == len(b), f"{what}: {key} length {len(a)} != {len(b)}"
        if key == "best" and not check_ambig_indices:
            a, b = a.copy(), b.copy()
            for f in ("ambig_a", "ambig_b"):
                a[f] = 0
                b[f] = 0
        if a.tobytes() != b.tobytes():
            bad = np.nonzero(a != b)[0]
            raise AssertionError(f"{what}: {key} differs at {bad[:5]}: got {a[bad[:3]]} want {b[bad[:3]]}")
    for key in ("n_probes", "n_hits"):
        if key in want and want[key] and key in got:
            assert got[key] == want[key], f"{what}: {key} {got[key]} != {want[key]}"
