"""Replays tests/golden/golden_v1.npz -- outputs of the reference's own object code (see make_golden.py) --
against the plain-C oracle (CPU, always) and against the CUDA path through the C ABI (GPU)."""
import os

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
N_PARAM_SETS = 4


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN, allow_pickle=False)


def _batch(G, which="proteins"):
    if which == "proteins":
        return synth.Batch(G["residues"], G["offsets"])
    return synth.Batch(G["read_bases"], G["read_offsets"])


def _fam(G):
    return synth.FamilyTables(G["fam_kmers"], G["fam_fam_off"], G["fam_fam_ids"], list(G["fam_pgf"]), list(G["fam_plf"]),
                              list(G["fam_function"]), G["fam_fam_func_sid"], G["fam_fam_pgf"], list(G["fam_pgf_names"]),
                              G["fam_func_sid"], 0)


def _want(G, k):
    return {key: G[f"p{k}_{key}"] for key in ("call_offsets", "calls", "hit_offsets", "hits", "otu_offsets", "otus", "best")}


def _fmt(x):
    return f"{float(x):.6g}"


def _fq_text(res, ids, fam, names):
    """fq_process_request.cc:349-362 over structured results (ids -> names)."""
    out = []
    for r in range(res["n"]):
        if not ids[r] or not res["best_score"][r] > 0:
            continue
        a, b = int(res["match_offsets"][r]), int(res["match_offsets"][r + 1])
        cols = [ids[r], str(int(res["best_frame"][r])), _fmt(res["best_score"][r])]
        for m in res["matches"][a:b]:
            cols += [str(int(m["length"])), fam.pgf_names[m["gfam"]] if m["gfam"] >= 0 else "", _fmt(m["gfam_score"]),
                     fam.plf[m["lfam"]] if m["lfam"] >= 0 else "", _fmt(m["lfam_score"]),
                     names[m["function_index"]] if m["function_index"] >= 0 else "hypothetical protein", _fmt(m["score"])]
        out.append("\t".join(cols) + "\n")
    return "".join(out)


# ---------------------------------------------------------------------------------------------- CPU: oracle
def test_oracle_calls_hits_otus_best(checkers, G):
    orc = checkers.Oracle().open_image(G["image"].copy())
    names = synth.function_names(int(G["n_functions"]))
    for k in range(N_PARAM_SETS):
        oc, mh, mwh, mg = (int(x) for x in G[f"p{k}_params"])
        orc.set_params(oc, mh, mwh, mg)
        got = orc.call_batch(_batch(G), 15)
        wl.assert_results_equal(got, _want(G, k), f"oracle vs golden p{k}", check_ambig_indices=False)
        fn = [checkers.best_function_string(r, lambda i: names[i] if 0 <= i < len(names) else "INVALID_OFFSET") for r in got["best"]]
        assert fn == list(G[f"p{k}_best_function"])
    orc.close()


def test_oracle_translation_family_fq_matrix(checkers, G):
    orc = checkers.Oracle().open_image(G["image"].copy())
    fam = _fam(G)
    orc.family_load(fam)
    names = synth.function_names(int(G["n_functions"]))
    reads = _batch(G, "reads")
    for i in range(reads.n):
        s = reads.seq(i)
        if len(s) >= 3:
            assert "".join(f"{f}\t{','.join(t.decode() for t in toks)}\n" for f, toks in orc.six_frames(s)) == str(G["six_frames"][i])
    ref = dict(gscore=G["family_gscore"], lscore=G["family_lscore"], score=G["family_score"], gfam=list(G["family_gfam"]),
               lfam=list(G["family_lfam"]), function=list(G["family_function"]))
    wl.assert_family_equal(orc.family_batch(_batch(G)), ref, fam, names, "oracle vs golden family")
    assert _fq_text(orc.fq_batch(reads), list(G["read_ids"]), fam, names) == str(G["fq_text"])
    # /add in two chunks then /matrix
    prot = _batch(G)
    ids = list(G["ids"])
    orc.postings_new()
    eids = np.arange(prot.n, dtype=np.uint32)
    orc.postings_add(eids, prot)
    assert wl.matrix_text_from_pairs(orc.matrix_rows(eids, prot), eids, prot, dict(enumerate(ids))) == str(G["matrix_text"])
    orc.close()


# ---------------------------------------------------------------------------------------------- GPU: CUDA path
@pytest.mark.gpu
def test_cuda_against_golden(checkers, G, tmp_path):
    d = str(tmp_path)
    G["image"].tofile(os.path.join(d, "kmer.table.mem_map"))
    synth.write_index_files(d, int(G["n_functions"]), int(G["n_otus"]))
    guts = api.KmerGuts(kmer_dir=d)
    prot, reads, ids = _batch(G), _batch(G, "reads"), list(G["ids"])
    for k in range(N_PARAM_SETS):
        oc, mh, mwh, mg = (int(x) for x in G[f"p{k}_params"])
        guts.set_parameters(dict(order_constraint=oc, min_hits=mh, min_weighted_hits=mwh, max_gap=mg))
        got = guts.process_aa_seq_batch(prot.residues, prot.offsets, 15)
        wl.assert_results_equal(got, _want(G, k), f"cuda vs golden p{k}", check_ambig_indices=False)
        assert [guts.best_function(r) for r in got["best"]] == list(G[f"p{k}_best_function"])
    guts.set_default_parameters()
    for details, fbc in ((0, 0), (1, 0), (0, 1)):
        assert guts.query_text(ids, prot.residues, prot.offsets, details, fbc) == str(G[f"query_text_{details}{fbc}"])
    fam = _fam(G)
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    ref = dict(gscore=G["family_gscore"], lscore=G["family_lscore"], score=G["family_score"], gfam=list(G["family_gfam"]),
               lfam=list(G["family_lfam"]), function=list(G["family_function"]))
    wl.assert_family_equal(guts.find_best_family_match_batch(prot.residues, prot.offsets), ref, fam,
                           synth.function_names(int(G["n_functions"])), "cuda vs golden family")
    assert guts.fq_text(list(G["read_ids"]), reads.residues, reads.offsets) == str(G["fq_text"])
    mapping = api.KmerPegMapping()
    half = int(G["add_split"])
    for lo, hi, silent, key in ((0, half, 0, "add_text_0"), (half, prot.n, 1, "add_text_1")):
        part = synth.Batch(prot.residues[int(prot.offsets[lo]):int(prot.offsets[hi])], prot.offsets[lo:hi + 1] - prot.offsets[lo])
        assert guts.add_text(mapping, ids[lo:hi], part.residues, part.offsets, silent) == str(G[key])
    assert guts.matrix_text(mapping, ids, prot.residues, prot.offsets) == str(G["matrix_text"])
    guts.close()
