"""The fused K1 (ckm_warp_scan.cuh: the ordered scoring scan and find_best_call run inside the probing warp, no hit record
leaves the SM) against the oracle and against the unfused pair K1 + scan_kernel.  Calls, their order, every f32
weighted_hits sum and every best-call record must be bit-identical for every parameter set, with and without the
neighbour copy of the table."""
import os

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth
from test_gpu_chain import indel_batch, tricky_world

pytestmark = pytest.mark.gpu

ALL = api.WANT_CALLS | api.WANT_HITS | api.WANT_OTU | api.WANT_BEST
PARAM_SETS = [
    dict(),
    dict(min_hits=3, max_gap=50),                        # gaps inside a 128-window step: every step takes the serial path
    dict(min_hits=2, min_weighted_hits=20, max_gap=10),
    dict(max_gap=126), dict(max_gap=127), dict(max_gap=128),   # either side of "no gap can fall inside a step"
    dict(max_gap=0), dict(max_gap=-1), dict(max_gap=-200),     # unsigned arithmetic of kguts.cc:821
    dict(min_hits=1), dict(min_hits=1, max_gap=3),       # runs of one hit (no carry below two stored hits)
    dict(min_hits=8, max_gap=2000), dict(min_hits=40, min_weighted_hits=100),
]


def _open(img, names, chain):
    os.environ.update(CKM_CHAIN=chain, CKM_OCCUPANCY_BITMAP="1")
    try:
        return api.KmerGuts(image=img, function_names=names)
    finally:
        os.environ.pop("CKM_CHAIN", None)
        os.environ.pop("CKM_OCCUPANCY_BITMAP", None)


def _check(g, orc, batch, prm, what):
    orc.set_params(**prm)
    g.set_parameters(prm)
    want = orc.call_batch(batch, api.WANT_CALLS | api.WANT_BEST)
    for flags in (api.WANT_BEST, api.WANT_CALLS, api.WANT_CALLS | api.WANT_BEST):
        keys = [k for k, f in (("call_offsets", api.WANT_CALLS), ("calls", api.WANT_CALLS), ("best", api.WANT_BEST)) if flags & f]
        g.set_tuning(api.TUNE_NO_FALLBACK)
        got = g.process_aa_seq_batch(batch.residues, batch.offsets, flags)
        assert g.last_batch_was_fused
        wl.assert_results_equal(got, {k: want[k] for k in keys}, f"{what} fused flags={flags} {prm}")
        assert got["n_probes"] == want["n_probes"]
        g.set_tuning(api.TUNE_NO_FALLBACK | api.TUNE_UNFUSED)
        other = g.process_aa_seq_batch(batch.residues, batch.offsets, flags)
        wl.assert_results_equal(other, {k: want[k] for k in keys}, f"{what} unfused flags={flags} {prm}")
        assert not g.last_batch_was_fused
    g.set_tuning(0)


@pytest.mark.parametrize("chain", ["0", "1"])
def test_fused_scan_bit_exact_all_parameter_sets(checkers, chain):
    protos, sig, img = wl.small_world()
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions), chain)
    try:
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 3000))
        batch = wl.concat_batches(batch, indel_batch(protos, 11, 500))
        for prm in PARAM_SETS:
            _check(g, orc, batch, prm, f"small world chain={chain}")
    finally:
        orc.set_params()
        g.close()
        orc.close()


@pytest.mark.parametrize("chain", ["0", "1"])
def test_fused_scan_few_functions_many_switches(checkers, chain):
    """Three functions over prototypes that share segments: strays of another function inside runs, two-in-a-row switches,
    carried-over hit pairs and sandwiches for find_best_call, in nearly every protein."""
    protos, sig, img = tricky_world()
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions), chain)
    try:
        aa = synth.AA
        own = synth.batch_from_strings([aa[protos.codes[int(protos.offsets[i]):int(protos.offsets[i + 1])]].tobytes()
                                        for i in range(protos.n)])
        # chimeras of three and four prototypes: several calls per protein, switches inside and across steps
        rng = np.random.default_rng(23)
        chim = []
        for _ in range(1500):
            parts = []
            for _k in range(int(rng.integers(2, 5))):
                p = int(rng.integers(0, protos.n))
                s = aa[protos.codes[int(protos.offsets[p]):int(protos.offsets[p + 1])]]
                a = int(rng.integers(0, max(1, len(s) - 20)))
                parts.append(s[a:a + int(rng.integers(9, 140))])
            chim.append(np.concatenate(parts).tobytes())
        batch = wl.concat_batches(wl.concat_batches(wl.edge_batch(protos), own), synth.batch_from_strings(chim))
        batch = wl.concat_batches(batch, synth.make_proteins(5, protos, 2500))
        for prm in PARAM_SETS[:4] + PARAM_SETS[9:11]:
            _check(g, orc, batch, prm, f"tricky world chain={chain}")
    finally:
        orc.set_params()
        g.close()
        orc.close()


def test_fused_scan_long_proteins_and_general_fallback(checkers):
    """Proteins of many steps (state carried over dozens of steps), and a batch holding one protein long enough to
    saturate the 39 998-hit window: that batch is served by K1 + scan_kernel<GENERAL> on its own accord."""
    protos, sig, img = wl.small_world(seed=8, n_protos=150, n_sigs=40_000)
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions), "1")
    try:
        rng = np.random.default_rng(17)
        aa = synth.AA
        seqs = []
        for n_parts in (4, 5, 9, 17, 40, 3, 120):
            parts = [aa[protos.codes[int(protos.offsets[p]):int(protos.offsets[p + 1])]] for p in rng.integers(0, protos.n, n_parts)]
            s = np.concatenate(parts).copy()
            sub = rng.random(len(s)) < 0.04
            s[sub] = aa[rng.integers(0, 20, int(sub.sum()))]
            seqs.append(s.tobytes())
        assert 30_000 < max(len(x) for x in seqs) < 40_000
        batch = wl.concat_batches(synth.batch_from_strings(seqs), synth.make_proteins(6, protos, 300))
        for prm in (dict(), dict(min_hits=3, max_gap=50)):
            _check(g, orc, batch, prm, "long proteins")
        # one protein of 45 000 residues: GENERAL for the whole batch, fused or not asked
        p0 = aa[protos.codes[int(protos.offsets[0]):int(protos.offsets[1])]]
        big = np.tile(p0, 45_000 // len(p0) + 1)[:45_000].tobytes()
        batch2 = wl.concat_batches(synth.batch_from_strings([big]), synth.make_proteins(7, protos, 200))
        orc.set_params()
        g.set_default_parameters()
        want = orc.call_batch(batch2, api.WANT_CALLS | api.WANT_BEST)
        got = g.process_aa_seq_batch(batch2.residues, batch2.offsets, api.WANT_CALLS | api.WANT_BEST)
        assert not g.last_batch_was_fused
        wl.assert_results_equal(got, {k: want[k] for k in ("call_offsets", "calls", "best")}, "saturating protein")
    finally:
        orc.set_params()
        g.close()
        orc.close()


def test_fused_scan_through_the_pipelined_host_path_and_device_entry(checkers):
    protos, sig, img = wl.small_world(seed=6)
    orc = checkers.Oracle().open_image(img)
    os.environ.update(CKM_PIPELINE_MIN_KB="0", CKM_PIPELINE_CHUNK_KB="64")
    try:
        for chain in ("0", "1"):
            g = _open(img, synth.function_names(sig.n_functions), chain)
            batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(9, protos, 5000))
            want = orc.call_batch(batch, api.WANT_BEST)
            got = g.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_BEST)  # ~25 chunks on two streams
            assert got["best"].tobytes() == want["best"].tobytes() and got["n_probes"] == want["n_probes"]
            g.close()
    finally:
        os.environ.pop("CKM_PIPELINE_MIN_KB", None)
        os.environ.pop("CKM_PIPELINE_CHUNK_KB", None)
        orc.close()


def test_packed_residue_entry_matches_ascii_entry(checkers):
    """ckm_call_batch_packed (five bits per residue, unpacked on the device) against the oracle on the ASCII strings: edge
    sequences (empty, shorter than a k-mer, lowercase, ambiguity codes, an embedded NUL), every flag combination, a rebased
    offsets array, and the chunked two-stream path that best-call-only batches take."""
    protos, sig, img = wl.small_world(seed=6)
    orc = checkers.Oracle().open_image(img)
    rng = np.random.default_rng(5)
    lens = [0, 1, 6, 7, 12, 13, 19, 25, 26, 31, 32, 33, 63, 64, 65, 127, 128, 129, 640, 641]  # around 5 L = 0 (mod 32)
    aa = synth.AA
    extra = synth.batch_from_strings([aa[rng.integers(0, 20, L)].tobytes() for L in lens])
    batch = wl.concat_batches(wl.concat_batches(wl.edge_batch(protos), extra), synth.make_proteins(9, protos, 4000))
    packed, woff = api.pack_residues(batch.residues, batch.offsets)
    assert int(woff[-1]) == int(((np.diff(batch.offsets.astype(np.int64)) + 6) // 7).sum())
    try:
        for chain in ("0", "1"):
            g = _open(img, synth.function_names(sig.n_functions), chain)
            for prm in (dict(), dict(order_constraint=1, min_hits=2, max_gap=30)):
                orc.set_params(**prm)
                g.set_parameters(prm)
                want = orc.call_batch(batch, ALL)
                got = g.process_packed_batch(packed, woff, ALL)
                wl.assert_results_equal(got, want, f"packed entry, all flags, chain={chain} {prm}")
                assert got["n_probes"] == want["n_probes"] and got["n_hits"] == len(want["hits"])
                got = g.process_packed_batch(packed, woff, api.WANT_CALLS | api.WANT_BEST)
                wl.assert_results_equal(got, {k: want[k] for k in ("call_offsets", "calls", "best")}, f"packed entry, fused {prm}")
                k = 41
                got = g.process_packed_batch(packed, woff[k:], api.WANT_BEST)
                assert got["best"].tobytes() == want["best"][k:].tobytes()
            g.close()
        os.environ.update(CKM_PIPELINE_MIN_KB="0", CKM_PIPELINE_CHUNK_KB="64")
        g = _open(img, synth.function_names(sig.n_functions), "1")
        orc.set_params()
        want = orc.call_batch(batch, api.WANT_BEST)
        got = g.process_packed_batch(packed, woff, api.WANT_BEST)  # ~25 chunks on two streams, each unpacked behind its copy
        assert got["best"].tobytes() == want["best"].tobytes() and got["n_probes"] == want["n_probes"]
        got = g.process_packed_batch(packed, woff[17:], api.WANT_BEST)
        assert got["best"].tobytes() == want["best"][17:].tobytes()
        g.close()
    finally:
        os.environ.pop("CKM_PIPELINE_MIN_KB", None)
        os.environ.pop("CKM_PIPELINE_CHUNK_KB", None)
        orc.close()


@pytest.mark.parametrize("chain", ["0", "1"])
def test_fused_scan_sparse_signatures_with_foreign_functions(checkers, chain):
    """Signature sets as build_signature_kmers leaves them (synth.make_signatures_sparse): a random subset of every prototype's
    windows, a share of them of some other function, avg_from_end jittered.  Nearly every run carries stray hits of other
    functions -- singly (no effect on the run), in pairs (a run-ending pair, kguts.cc:852-856, with its carry-over), at step
    boundaries -- which is what the parallel form of the run logic in probe_pc_kernel has to get right."""
    for seed, keep, foreign, nf in ((3, 0.7, 0.15, 6), (4, 0.4, 0.3, 3), (5, 1.0, 0.05, 50)):
        protos = synth.make_prototypes(seed, 300, 300, 40.0)
        sig = synth.make_signatures_sparse(protos, 80_000, keep=keep, foreign=foreign, n_functions=nf, seed=seed)
        img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
        orc = checkers.Oracle().open_image(img)
        g = _open(img, synth.function_names(sig.n_functions), chain)
        try:
            batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(seed + 10, protos, 3000))
            for prm in (dict(), dict(min_hits=2, max_gap=30), dict(min_hits=1, max_gap=3), dict(min_hits=3, min_weighted_hits=8, max_gap=127)):
                _check(g, orc, batch, prm, f"sparse world keep={keep} foreign={foreign} functions={nf} chain={chain}")
        finally:
            orc.set_params()
            g.close()
            orc.close()
