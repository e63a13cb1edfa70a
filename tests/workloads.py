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


def _close(a, b, rel=1e-6):
    return abs(float(a) - float(b)) <= rel * max(1.0, abs(float(a)), abs(float(b)))


def assert_family_equal(got, ref, fam, function_names, what):
    """got: FAMILY_DT records (ids); ref: dict of strings + exact f32 scores from the reference.  Scores of the
    best family and the best call must be bit-equal; the rolled-up PGF score within 1e-6 relative; names equal
    except where the scores show an exact tie (the reference breaks ties by unordered_map iteration order)."""
    n = len(got)
    ties = 0
    for i in range(n):
        g = got[i]
        fn = function_names[g["function_index"]] if g["function_index"] >= 0 else "hypothetical protein"
        assert g["lfam_score"] == ref["lscore"][i], f"{what}[{i}] lfam_score {g['lfam_score']} != {ref['lscore'][i]}"
        assert g["score"] == ref["score"][i], f"{what}[{i}] score"
        assert fn == ref["function"][i], f"{what}[{i}] function {fn!r} != {ref['function'][i]!r}"
        assert _close(g["gfam_score"], ref["gscore"][i]), f"{what}[{i}] gfam_score {g['gfam_score']} vs {ref['gscore'][i]}"
        gname = fam.pgf_names[g["gfam"]] if g["gfam"] >= 0 else ""
        lname = fam.plf[g["lfam"]] if g["lfam"] >= 0 else ""
        assert (gname == "") == (ref["gfam"][i] == "") and (lname == "") == (ref["lfam"][i] == "")
        if (gname, lname) != (ref["gfam"][i], ref["lfam"][i]):
            ties += 1
    assert ties <= max(3, n // 50), f"{what}: {ties} id differences is more than exact ties explain"
    return ties


def assert_family_records_equal(got, want, what):
    """CUDA vs oracle: same tie rule on both sides, so ids match exactly; only the PGF roll-up order differs."""
    assert len(got) == len(want)
    for f in ("lfam", "lfam_score", "score", "function_index"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, f"{what}: {f} differs at {bad[:5]}: {got[f][bad[:5]]} vs {want[f][bad[:5]]}"
    for i in range(len(got)):
        assert _close(got["gfam_score"][i], want["gfam_score"][i]), f"{what}[{i}] gfam_score"
        if got["gfam"][i] != want["gfam"][i]:  # a near-tie decided by summation order
            assert _close(got["gfam_score"][i], want["gfam_score"][i])


def assert_fq_equal(got, ref, fam, function_names, what):
    """got: dict from fq_batch (ids); ref: list of (frame, score, [(len, gfam, gscore, lfam, lscore, function, score)])."""
    assert got["n"] == len(ref)
    for i, (fr, bs, ms) in enumerate(ref):
        a, b = int(got["match_offsets"][i]), int(got["match_offsets"][i + 1])
        assert got["best_frame"][i] == fr, f"{what}[{i}] frame {got['best_frame'][i]} != {fr}"
        assert got["best_score"][i] == bs, f"{what}[{i}] best_score"
        assert b - a == len(ms), f"{what}[{i}] n_matches {b - a} != {len(ms)}"
        for k, m in enumerate(ms):
            x = got["matches"][a + k]
            fn = function_names[x["function_index"]] if x["function_index"] >= 0 else "hypothetical protein"
            assert (x["length"], x["lfam_score"], x["score"], fn) == (m[0], np.float32(m[4]), np.float32(m[6]), m[5]), f"{what}[{i}][{k}]"
            assert _close(x["gfam_score"], m[2]), f"{what}[{i}][{k}] gfam_score"


def assert_fq_records_equal(got, want, what):
    for f in ("best_frame", "best_score", "match_offsets"):
        assert np.array_equal(got[f], want[f]), f"{what}: {f}"
    for f in ("length", "lfam", "lfam_score", "score", "function_index"):
        assert np.array_equal(got["matches"][f], want["matches"][f]), f"{what}: matches.{f}"
    assert np.allclose(got["matches"]["gfam_score"], want["matches"]["gfam_score"], rtol=1e-6, atol=0)
    assert got["n_fragments"] == want["n_fragments"]


def matrix_text_from_pairs(pairs, eids, batch, id_to_peg):
    """process_results (matrix_request.cc:163-189) over ordered, merged COO entries; floats like ostream << float."""
    from close_kmers_b200 import api
    pairs = api.merge_pairs(pairs)
    lens = {}
    for i, e in enumerate(eids):
        lens[int(e)] = int(batch.offsets[i + 1] - batch.offsets[i])
    out = []
    for p in pairs:
        score = np.float32(p["count"]) / np.float32(lens[int(p["eid_i"])] + lens[int(p["eid_j"])])
        out.append(f"{id_to_peg[int(p['eid_i'])]}\t{id_to_peg[int(p['eid_j'])]}\t{int(p['count'])}\t{float(score):.6g}\n")
    return "".join(out)


def assert_fq_text_equal(got: str, want: str):
    """/fq_lookup text: same reads, every numeric field and function byte-identical.  The PGF / PLF *names* may differ where
    two families tie exactly on the printed score (the reference breaks ties by unordered_map iteration order, here the
    smallest id wins); returns how many lines did."""
    a, b = got.splitlines(), want.splitlines()
    assert len(a) == len(b)
    named = 0
    for x, y in zip(a, b):
        if x == y:
            continue
        fx, fy = x.split("\t"), y.split("\t")
        assert len(fx) == len(fy), (x, y)
        for k, (u, v) in enumerate(zip(fx, fy)):
            if u != v:
                assert k >= 3 and (k - 3) % 7 in (1, 3), (k, x, y)  # a family name; its score (next field) is compared like the rest
        named += 1
    return named
