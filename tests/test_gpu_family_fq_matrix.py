"""GPU parity: family voting, fastq 6-frame path, /add + /matrix and the handler texts, through the C ABI,
against the plain-C oracle and (when oracle/_ref is present) the reference's own object code."""
import os

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth
from test_oracle_family_fq_matrix import DNA_EDGE, nr_chunks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(checkers, tmp_path_factory):
    protos, sig, img = wl.small_world(otu_mode="mixed")
    fam = synth.make_families(7, sig)
    d = str(tmp_path_factory.mktemp("kmerdir"))
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    guts = api.KmerGuts(kmer_dir=d)
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    ref = None
    if os.path.exists(checkers.REF_SO):
        ref = checkers.Ref().open(d)
        ref.set_params()
        ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    yield protos, sig, fam, orc, guts, ref, d
    guts.close()
    orc.close()


def test_family_voting(checkers, world):
    protos, sig, fam, orc, guts, ref, _ = world
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 3000))
    for prm in (dict(), dict(min_hits=3, max_gap=50)):
        orc.set_params(**prm)
        guts.set_parameters(prm)
        want = orc.family_batch(batch)
        got = guts.find_best_family_match_batch(batch.residues, batch.offsets)
        wl.assert_family_records_equal(got, want, f"cuda vs oracle {prm}")
    guts.set_default_parameters()
    orc.set_params()
    if ref is not None:
        wl.assert_family_equal(got if not prm else guts.find_best_family_match_batch(batch.residues, batch.offsets),
                               ref.family_batch(batch), fam, synth.function_names(sig.n_functions), "cuda vs reference")
    assert int((want["lfam"] >= 0).sum()) > 1000


def test_family_nr_load_on_gpu(checkers, world):
    """N2: the k-mer -> family table built on the device from families.nr proteins equals the one the reference's
    NRLoader / KmerInserter / add_fam_mapping build, and votes identically afterwards."""
    import dataclasses
    protos, sig, fam, _, _, _, d = world
    _, _, img = wl.small_world(otu_mode="mixed")
    chunks = nr_chunks(protos, fam.n_fams)
    orc = checkers.Oracle().open_image(img)
    want = orc.family_nr_build(chunks)
    guts = api.KmerGuts(kmer_dir=d)
    try:
        for rep in range(2):  # a second begin() starts from an empty table; it also dedupes after every chunk
            os.environ["CKM_FAMNR_COMPACT_AT"] = "1000" if rep else str(1 << 28)
            guts.family_nr_begin()
            for fam_ids, batch in chunks:
                guts.family_nr_add(fam_ids, batch.residues, batch.offsets)
            nk, ne = guts.family_nr_finish(fam.pgf, fam.plf, fam.function)
            assert (nk, ne) == (len(want[0]), len(want[2]))
            got = guts.family_export(nk, ne)
            for a, b, name in zip(got, want, ("kmers", "fam_off", "fam_ids")):
                np.testing.assert_array_equal(a, b, err_msg=name)
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(78, protos, 2500))
        orc.family_load(dataclasses.replace(fam, kmers=want[0], fam_off=want[1], fam_ids=want[2]))
        got_m = guts.find_best_family_match_batch(batch.residues, batch.offsets)
        wl.assert_family_records_equal(got_m, orc.family_batch(batch), "cuda vs oracle on the nr-loaded table")
        if os.path.exists(checkers.REF_SO):
            ref = checkers.Ref().open(d)
            ref.set_params()
            for fam_ids, b in chunks:
                ref.family_nr_add(fam_ids, b)
            for a, b, name in zip(got, ref.family_table(), ("kmers", "fam_off", "fam_ids")):
                np.testing.assert_array_equal(a, b, err_msg="reference " + name)
            ref.family_set_data(fam.pgf, fam.plf, fam.function)
            wl.assert_family_equal(got_m, ref.family_batch(batch), fam, synth.function_names(sig.n_functions), "cuda vs reference")
            ref.close()
        # an empty load installs an empty table: nobody matches
        guts.family_nr_begin()
        assert guts.family_nr_finish(fam.pgf, fam.plf, fam.function) == (0, 0)
        assert int((guts.find_best_family_match_batch(batch.residues, batch.offsets)["lfam"] >= 0).sum()) == 0
    finally:
        os.environ.pop("CKM_FAMNR_COMPACT_AT", None)
        guts.close()
        orc.close()


def test_family_large_fanout_uses_global_maps(checkers):
    """Proteins whose hits touch more family-list entries than fit the shared-memory map (E > 640)."""
    protos, sig, img = wl.small_world(seed=31, n_protos=120, n_sigs=30_000, n_functions=3, otu_mode="minus1")
    fam = synth.make_families(9, sig, fams_per_function=40, max_list=40, coverage=1.0)
    # long lists: every k-mer gets 10..40 families
    rng = np.random.default_rng(2)
    cnt = rng.integers(10, 41, len(fam.kmers))
    fam.fam_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    owner = np.repeat(np.arange(len(cnt)), cnt)
    rank = np.arange(int(fam.fam_off[-1])) - fam.fam_off[:-1].astype(np.int64)[owner]
    start = rng.integers(0, fam.n_fams, len(cnt))[owner]
    fam.fam_ids = ((start + rank) % fam.n_fams).astype(np.uint32)
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    guts = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    batch = synth.make_proteins(3, protos, 300)
    wl.assert_family_records_equal(guts.find_best_family_match_batch(batch.residues, batch.offsets), orc.family_batch(batch),
                                   "large fan-out")
    guts.close()
    orc.close()


def test_family_many_distinct_families_overflow_small_maps(checkers):
    """Proteins whose map could live in shared memory (E <= 512 entries) but that touch more distinct families than the
    optimistic 256-slot map takes: the SMALL vote kernel hands them to the LARGE one.  Mixed with proteins that stay small
    and proteins with global-scratch maps (E > 512)."""
    protos, sig, img = wl.small_world(seed=37, n_protos=300, n_sigs=24_000, n_functions=3, otu_mode="minus1", mean_len=90)
    fam = synth.make_families(9, sig, fams_per_function=1500, max_list=8, coverage=1.0)
    rng = np.random.default_rng(3)
    cnt = rng.integers(4, 9, len(fam.kmers))
    fam.fam_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    # every list: distinct ids spread over all 4500 families
    start = rng.integers(0, fam.n_fams, len(cnt))
    owner = np.repeat(np.arange(len(cnt)), cnt)
    rank = np.arange(int(fam.fam_off[-1])) - fam.fam_off[:-1].astype(np.int64)[owner]
    fam.fam_ids = ((start[owner] + rank * 517) % fam.n_fams).astype(np.uint32)
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    guts = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    try:
        batch = wl.concat_batches(synth.make_proteins(4, protos, 1500), wl.edge_batch(protos))
        want_sc, want_off = orc.family_scores(batch)
        per_protein = np.diff(want_off.astype(np.int64))
        assert int((per_protein > 128).sum()) > 200 and int((per_protein <= 128).sum()) > 20  # both classes present
        got = guts.family_scores(batch.residues, batch.offsets)
        np.testing.assert_array_equal(got["score_offsets"], want_off)
        for f in ("id", "hit_count", "weighted_total"):
            np.testing.assert_array_equal(got["scores"][f], want_sc[f], err_msg=f)
        wl.assert_family_records_equal(guts.find_best_family_match_batch(batch.residues, batch.offsets), orc.family_batch(batch),
                                       "overflow hand-off")
    finally:
        guts.close()
        orc.close()


def test_six_frame_translation(checkers, world):
    protos, _, _, orc, guts, _, _ = world
    rng = np.random.default_rng(5)
    seqs = list(DNA_EDGE) + [bytes(rng.choice(np.frombuffer(b"ACGTacgtNRYU", np.uint8), int(rng.integers(0, 200)))) for _ in range(300)]
    seqs.append(b"".join(bytes([a, b, c]) for a in b"ACGT" for b in b"ACGT" for c in b"ACGT"))
    batch = synth.batch_from_strings(seqs)
    for min_len in (0, 10):
        got = guts.get_possible_proteins_batch(batch.residues, batch.offsets, min_len)
        for i, s in enumerate(seqs):
            want = [(f, [t for t in toks if len(t) > min_len]) for f, toks in orc.six_frames(s)] if len(s) >= 0 else None
            assert got[i] == want, (s, min_len)


def test_fq_best_frame(checkers, world):
    protos, sig, fam, orc, guts, ref, _ = world
    batch = wl.concat_batches(synth.batch_from_strings(DNA_EDGE), synth.make_reads(3, protos, 4000))
    want = orc.fq_batch(batch)
    got = guts.fq_batch(batch.residues, batch.offsets)
    wl.assert_fq_records_equal(got, want, "cuda vs oracle")
    assert int((want["best_frame"] != 0).sum()) > 2000 and got["n_probes"] > 0
    if ref is not None:
        sub = synth.Batch(batch.residues[: int(batch.offsets[700])], batch.offsets[:701])
        wl.assert_fq_equal(guts.fq_batch(sub.residues, sub.offsets), ref.fq_batch(sub), fam,
                           synth.function_names(sig.n_functions), "cuda vs reference")


def _lines_by_id(text):
    return {ln.split("\t")[0]: ln for ln in text.splitlines()}


def test_handler_texts_match_reference(checkers, world):
    """Byte-identical response text for /query (all three modes) and /fq_lookup (modulo exact family ties)."""
    protos, sig, fam, orc, guts, ref, _ = world
    if ref is None:
        pytest.skip("oracle/_ref/libckm_ref.so absent")
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(5, protos, 400))
    ids = [f"fig|83333.1.peg.{i}" for i in range(batch.n)]
    guts.set_default_parameters()
    for details, fbc in ((0, 0), (1, 0), (0, 1)):
        assert guts.query_text(ids, batch.residues, batch.offsets, details, fbc) == ref.query_text(ids, batch, details, fbc)
    reads = wl.concat_batches(synth.batch_from_strings(DNA_EDGE), synth.make_reads(8, protos, 500))
    rids = [f"read{i}" if i != 3 else "" for i in range(reads.n)]
    got_text, want_text = guts.fq_text(rids, reads.residues, reads.offsets), ref.fq_text(rids, reads)
    mine, theirs = _lines_by_id(got_text), _lines_by_id(want_text)
    assert mine.keys() == theirs.keys() and len(mine) > 300
    # a line may differ from the reference's only in the NAME of a family, and only where the score printed next to it is the
    # same on both sides: an exact tie, which the reference breaks by unordered_map iteration order (family_mapper.cc:137-197)
    named = wl.assert_fq_text_equal(got_text, want_text)
    assert named == sum(1 for k in mine if mine[k] != theirs[k])


def test_add_and_matrix(checkers, world):
    protos, sig, fam, orc, guts, ref, _ = world
    sub = synth.Prototypes(protos.codes[: int(protos.offsets[60])], protos.offsets[:61])
    batch = synth.make_proteins(11, sub, 600, mix=(0.9, 0.1, 0.0, 0.0))
    ids = [f"fig|{i % 570}.peg.{i % 570}" for i in range(batch.n)]
    mapping = api.KmerPegMapping()
    guts.postings_clear()
    orc.postings_new()
    orc.set_params()
    guts.set_default_parameters()
    if ref is not None:
        ref.mapping_new()
    half = batch.n // 2
    for lo, hi, silent in ((0, half, 0), (half, batch.n, 1)):
        part = synth.Batch(batch.residues[int(batch.offsets[lo]):int(batch.offsets[hi])], batch.offsets[lo:hi + 1] - batch.offsets[lo])
        text = guts.add_text(mapping, ids[lo:hi], part.residues, part.offsets, silent)
        if ref is not None:
            assert text == ref.add_text(ids[lo:hi], part, silent)
        orc.postings_add([mapping.encode_id(x) for x in ids[lo:hi]], part)
    assert guts.postings_count == orc.L.orc_postings_count(orc.post) > 10_000
    order = np.random.default_rng(1).permutation(batch.n)[:450]
    req = synth.batch_from_strings([batch.seq(i) for i in order])
    req_ids = [ids[i] for i in order]
    eids = np.array([mapping.encode_id(x) for x in req_ids], np.uint32)
    want = api.merge_pairs(orc.matrix_rows(eids, req))
    got = api.merge_pairs(guts.matrix_rows(eids, req.residues, req.offsets))
    assert got.tobytes() == want.tobytes() and len(want) > 500
    # row-block sharding: any partition of the rows gives the same matrix
    parts = np.concatenate([guts.matrix_rows(eids, req.residues, req.offsets, a, b) for a, b in ((0, 100), (100, 101), (101, 450))])
    assert api.merge_pairs(parts).tobytes() == want.tobytes()
    if ref is not None:
        assert guts.matrix_text(mapping, req_ids, req.residues, req.offsets) == ref.matrix_text(req_ids, req)


def test_matrix_popular_kmers_use_global_maps(checkers):
    """Rows whose hits walk more postings than the shared-memory map holds (every protein shares its k-mers)."""
    protos, sig, img = wl.small_world(seed=41, n_protos=20, n_sigs=5_000, otu_mode="minus1")
    orc = checkers.Oracle().open_image(img)
    guts = api.KmerGuts(image=img)
    batch = synth.make_proteins(3, protos, 400, mix=(1.0, 0.0, 0.0, 0.0), sub_rate=0.01)
    eids = np.arange(batch.n, dtype=np.uint32)
    orc.postings_new()
    orc.postings_add(eids, batch)
    guts.postings_add(eids, batch.residues, batch.offsets)
    want = api.merge_pairs(orc.matrix_rows(eids, batch))
    got = api.merge_pairs(guts.matrix_rows(eids, batch.residues, batch.offsets))
    assert got.tobytes() == want.tobytes() and len(want) > 3000
    guts.close()
    orc.close()
