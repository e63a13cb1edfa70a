"""Pins the oracle's family voting (F1-F3), 6-frame translation / fq best frame (D1-D3) and matrix (M1)
restatements to the reference's own object code (oracle/_ref)."""
import os
import tempfile

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

DNA_EDGE = [b"", b"A", b"AC", b"ACG", b"ACGT", b"acgtnACGTN" * 12, b"ATGAAATAGTAATGA" * 10, b"T" * 41, b"ACGURYKMSWBDHVN" * 8,
            b"TAATAGTGA" * 6, b"ATGGCC" + b"TAA" + b"GCT" * 14 + b"TGA"]


@pytest.fixture(scope="module")
def world(checkers):
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so not built (needs /root/reference)")
    protos, sig, img = wl.small_world(otu_mode="minus1")
    fam = synth.make_families(7, sig)
    d = tempfile.mkdtemp(prefix="ckm_ref_")
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 0)
    ref = checkers.Ref().open(d)
    ref.set_params()
    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    yield protos, sig, fam, ref, orc
    ref.close()
    orc.close()


def test_translation_all_codons_and_frames(checkers, world):
    _, _, _, ref, orc = world
    rng = np.random.default_rng(5)
    seqs = list(DNA_EDGE[3:]) + [bytes(rng.choice(np.frombuffer(b"ACGTacgtNRYU", np.uint8), int(rng.integers(3, 200)))) for _ in range(200)]
    # every codon over ACGT
    seqs.append(b"".join(bytes([a, b, c]) for a in b"ACGT" for b in b"ACGT" for c in b"ACGT"))
    for s in seqs:
        mine = "".join(f"{f}\t{','.join(t.decode() for t in toks)}\n" for f, toks in orc.six_frames(s))
        assert mine == ref.six_frames(s), s


def test_family_voting_matches_reference(checkers, world):
    protos, sig, fam, ref, orc = world
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 1500))
    r, o = ref.family_batch(batch), orc.family_batch(batch)
    wl.assert_family_equal(o, r, fam, synth.function_names(sig.n_functions), "oracle vs reference")
    assert int((o["lfam"] >= 0).sum()) > 500


def test_fq_best_frame_matches_reference(checkers, world):
    protos, sig, fam, ref, orc = world
    batch = wl.concat_batches(synth.batch_from_strings(DNA_EDGE), synth.make_reads(3, protos, 600))
    r, o = ref.fq_batch(batch), orc.fq_batch(batch)
    wl.assert_fq_equal(o, r, fam, synth.function_names(sig.n_functions), "oracle vs reference")
    assert int((o["best_frame"] != 0).sum()) > 300


def test_matrix_matches_reference(checkers, world):
    protos, sig, fam, ref, orc = world
    # 40 prototypes x 6 mutated copies, /add-ed in two chunks, then one matrix request over all of them
    sub = synth.Prototypes(protos.codes[: int(protos.offsets[40])], protos.offsets[:41])
    batch = synth.make_proteins(11, sub, 240, mix=(0.9, 0.1, 0.0, 0.0))
    ids = [f"fig|{i % 230}.peg.{i % 230}" for i in range(batch.n)]  # ten ids appear twice
    ref.mapping_new()
    orc.postings_new()
    eid_of = {}
    half = batch.n // 2
    for lo, hi in ((0, half), (half, batch.n)):
        part = synth.Batch(batch.residues[int(batch.offsets[lo]):int(batch.offsets[hi])], batch.offsets[lo:hi + 1] - batch.offsets[lo])
        ref.add_text(ids[lo:hi], part, silent=1)
        eids = [eid_of.setdefault(x, len(eid_of)) for x in ids[lo:hi]]
        orc.postings_add(eids, part)
    order = np.random.default_rng(1).permutation(batch.n)[:200]
    req = synth.batch_from_strings([batch.seq(i) for i in order])
    req_ids = [ids[i] for i in order]
    want = ref.matrix_text(req_ids, req)
    eids = np.array([eid_of[x] for x in req_ids], np.uint32)
    pairs = orc.matrix_rows(eids, req)
    got = wl.matrix_text_from_pairs(pairs, eids, req, {v: k for k, v in eid_of.items()})
    assert got == want
    assert len(pairs) > 100
    # row blocks partition the result
    parts = np.concatenate([orc.matrix_rows(eids, req, a, b) for a, b in ((0, 70), (70, 71), (71, 200))])
    assert wl.matrix_text_from_pairs(parts, eids, req, {v: k for k, v in eid_of.items()}) == want


def nr_chunks(protos, n_fams, seed=21, n=900):
    """families.nr stand-in: three chunks of proteins with their family ids; the middle chunk holds one protein without a
    family, which ends that chunk (nr_loader.cc:154-160)."""
    rng = np.random.default_rng(seed)
    chunks = []
    for c in range(3):
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(seed + c, protos, n // 3)) if c == 0 else \
            synth.make_proteins(seed + c, protos, n // 3)
        fam_ids = rng.integers(0, n_fams, batch.n).astype(np.uint32)
        fam_ids[: batch.n // 2] = fam_ids[: batch.n // 2] % 5  # many proteins of few families: duplicate (k-mer, family) pairs
        if c == 1:
            fam_ids[batch.n * 2 // 3] = 0xFFFFFFFF
        chunks.append((fam_ids, batch))
    return chunks


def test_family_nr_load_matches_reference(checkers, world):
    """N2: NRLoader family-mode load (thread_load -> KmerInserter -> add_fam_mapping) restated by the oracle."""
    import dataclasses
    protos, sig, fam, _, _ = world
    _, _, img = wl.small_world(otu_mode="minus1")
    d = tempfile.mkdtemp(prefix="ckm_ref_nr_")
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 0)
    ref = checkers.Ref().open(d)
    ref.set_params()
    orc = checkers.Oracle().open_image(img)
    chunks = nr_chunks(protos, fam.n_fams)
    for fam_ids, batch in chunks:
        ref.family_nr_add(fam_ids, batch)
    want = ref.family_table()
    got = orc.family_nr_build(chunks)
    for a, b, name in zip(got, want, ("kmers", "fam_off", "fam_ids")):
        np.testing.assert_array_equal(a, b, err_msg=name)
    assert len(got[0]) > 1000 and len(got[2]) > len(got[0])
    # a protein after the family-less one contributes nothing: rebuilding without the tail of chunk 1 gives the same table
    f1, b1 = chunks[1]
    cut = int(np.flatnonzero(f1 == 0xFFFFFFFF)[0])
    head = synth.Batch(b1.residues[: int(b1.offsets[cut])], b1.offsets[: cut + 1])
    again = orc.family_nr_build([chunks[0], (f1[:cut], head), chunks[2]])
    for a, b in zip(again, got):
        np.testing.assert_array_equal(a, b)
    # and the voting that runs on the loaded table agrees
    ref.family_set_data(fam.pgf, fam.plf, fam.function)
    orc.family_load(dataclasses.replace(fam, kmers=got[0], fam_off=got[1], fam_ids=got[2]))
    batch = synth.make_proteins(77, protos, 800)
    wl.assert_family_equal(orc.family_batch(batch), ref.family_batch(batch), fam, synth.function_names(sig.n_functions),
                           "oracle vs reference on the nr-loaded table")
    ref.close()
    orc.close()


def family_rows(fam):
    """family_data_ rows for the /lookup checks: genus ids cycle over 0..2, sizes and counts are arbitrary but distinct."""
    return [(fam.pgf[f], fam.plf[f], fam.function[f], f % 3, 1000 + 37 * f, 1 + f % 50) for f in range(fam.n_fams)]


def parse_lookup_blocks(text):
    """/lookup listing -> [(id, [row fields...])]"""
    blocks, cur = [], None
    for line in text.split("\n")[:-1]:
        if cur is None:
            cur = (line, [])
        elif line == "//":
            blocks.append(cur)
            cur = None
        elif line == "///":  # end of a family's representative rows (find_reps)
            cur[1][-1] = cur[1][-1] + ("///",)
        else:
            cur[1].append(tuple(line.split("\t")))
    assert cur is None
    return blocks


def assert_lookup_listing_equal(got, want, weight_col=2):
    """Same blocks; rows may be permuted among equal weighted_total (the reference's std::sort over an unordered_map leaves
    that order open), and a listing that stops inside such a tie group may hold a different subset of it."""
    g, w = parse_lookup_blocks(got), parse_lookup_blocks(want)
    assert [b[0] for b in g] == [b[0] for b in w]
    lenient = 0
    for (sid, a), (_, b) in zip(g, w):
        if sorted(a) == sorted(b):
            assert [float(r[weight_col]) for r in a] == sorted((float(r[weight_col]) for r in a), reverse=True)
            continue
        lenient += 1
        floor = max(float(a[-1][weight_col]) if a else np.inf, float(b[-1][weight_col]) if b else np.inf)
        assert sorted(r for r in a if float(r[weight_col]) > floor) == sorted(r for r in b if float(r[weight_col]) > floor), sid
    assert lenient <= max(2, len(g) // 20), lenient


def test_lookup_family_scores_match_reference(checkers, world):
    """N3: LookupRequest::on_hit accumulation (family mode) restated by the oracle, against the reference's listing."""
    protos, sig, fam, ref, orc = world
    rows = family_rows(fam)
    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)  # the matrix test replaced the mapping
    ref.family_set_extra([r[3] for r in rows], [r[4] for r in rows], [r[5] for r in rows])
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(31, protos, 400))
    ids = [f"s{i}" for i in range(batch.n)]
    sc, so = orc.family_scores(batch)
    # threshold 0 lists every family the sequence touched
    blocks = parse_lookup_blocks(ref.lookup_text(ids, batch, family_mode=True, kmer_hit_threshold=0))
    assert len(blocks) == batch.n
    fmt = lambda x: "%g" % np.float32(x)
    for i, (sid, got_rows) in enumerate(blocks):
        assert sid == ids[i]
        mine = sorted((str(e["hit_count"]), str(e["hit_count"]), fmt(e["weighted_total"]), rows[e["id"]][0], rows[e["id"]][1],
                       str(rows[e["id"]][4]), str(rows[e["id"]][5]), fmt(np.float32(e["hit_count"]) / np.float32(rows[e["id"]][4])),
                       rows[e["id"]][2]) for e in sc[int(so[i]):int(so[i + 1])])
        assert mine == sorted(got_rows), sid
    assert int(so[-1]) > 2000
