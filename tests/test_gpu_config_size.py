"""Parity at the sizes BASELINE.json names, not only on the small worlds of the other tests:

  * C1 -- 10 000 proteins against a 1M-signature-k-mer image, every output the ABI has (calls, hit lists, OTU maps, best calls)
    against the reference's own object code (oracle/_ref) -- the configuration the reference itself runs as is;
  * a 20 000-protein slice of a 20M-k-mer world (table 1 GB, larger than L2: occupancy bitmap, neighbour copy and the fused
    K1 are all in use, as for C2 / C3) -- calls, hits and best calls against the reference's object code, through the copy and
    with plain probing, fused and unfused.

Falls back to the plain-C oracle when oracle/_ref was not built (no /root/reference at build time)."""
import os
import shutil
import tempfile

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

pytestmark = pytest.mark.gpu

ALL = api.WANT_CALLS | api.WANT_HITS | api.WANT_OTU | api.WANT_BEST


def _kmer_dir(img, sig, n_otus):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > img.nbytes * 1.2 else None
    d = tempfile.mkdtemp(prefix="ckm_cfg_", dir=base)
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, n_otus)
    return d


def _checker(checkers, d, img):
    """(checker, is_reference): the reference's object code over the directory, else the C port over the image."""
    if os.path.exists(checkers.REF_SO):
        ref = checkers.Ref().open(d)
        ref.set_params()
        return ref, True
    return checkers.Oracle().open_image(img), False


def _compare(got, want, is_ref, guts, what):
    wl.assert_results_equal(got, {k: want[k] for k in got if k in want}, what, check_ambig_indices=not is_ref)
    if is_ref and "best" in got:  # the reference reports an ambiguous call as a string (kguts.cc:1176-1196)
        assert [guts.best_function(r) for r in got["best"]] == want["best_function"], what


def test_c1_all_outputs_against_reference(checkers):
    """BASELINE configs[0]: 10k synthetic proteins vs a 1M-signature-k-mer image."""
    n_sigs = 1_000_000
    protos = synth.make_prototypes(12345, -(-n_sigs // 293) + 8, 300, 0.0)
    sig = synth.make_signatures(protos, n_sigs, otu_mode="mixed")
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins_parallel(12346, protos, 10_000))
    d = _kmer_dir(img, sig, 12)
    try:
        chk, is_ref = _checker(checkers, d, img)
        guts = api.KmerGuts(kmer_dir=d)
        guts.set_default_parameters()
        want = chk.call_batch(batch, ALL)
        assert want["n_hits"] > 1_500_000 and len(want["calls"]) > 9_000
        _compare(guts.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, is_ref, guts, "c1 all flags")
        for flags in (api.WANT_CALLS, api.WANT_HITS, api.WANT_OTU, api.WANT_BEST):
            _compare(guts.process_aa_seq_batch(batch.residues, batch.offsets, flags), want, is_ref, guts, f"c1 flags={flags}")
        pk, woff = api.pack_residues(batch.residues, batch.offsets)
        _compare(guts.process_packed_batch(pk, woff, api.WANT_BEST | api.WANT_CALLS), want, is_ref, guts, "c1 packed entry")
        guts.close()
        chk.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_20m_kmer_world_slice_through_the_neighbour_copy(checkers):
    """A table larger than L2 (64,000,031 buckets x 16 B = 1 GB): bitmap + neighbour copy + fused K1, as for C2 / C3."""
    n_sigs = 20_000_000
    protos = synth.make_prototypes(12345, -(-n_sigs // 293) + 8, 300, 60.0)
    sig = synth.make_signatures(protos, n_sigs, dedupe=False)  # (the 0.2 % repeated 8-mers stay: first in probe order wins)
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins_parallel(777, protos, 20_000))
    d = _kmer_dir(img, sig, 0)
    try:
        chk, is_ref = _checker(checkers, d, img)
        del img
        guts = api.KmerGuts(kmer_dir=d)
        guts.set_default_parameters()
        assert guts.slot_bytes == 16 and guts.has_occupancy_bitmap and guts.chain_info["entries"] > 19_000_000
        flags3 = api.WANT_CALLS | api.WANT_HITS | api.WANT_BEST
        want = chk.call_batch(batch, flags3)
        assert want["n_hits"] > 3_000_000
        for name, tuning in (("copy", api.TUNE_NO_FALLBACK), ("plain", api.TUNE_NO_FALLBACK | api.TUNE_PLAIN_PROBE)):
            guts.set_tuning(tuning)
            got = guts.process_aa_seq_batch(batch.residues, batch.offsets, flags3)  # hit lists asked for: K1 + scan_kernel
            assert not guts.last_batch_was_fused
            _compare(got, want, is_ref, guts, f"20M world, {name}, calls + hits + best")
            if name == "copy":
                assert guts.chain_info["hits_from_copy"] > 0.8 * want["n_hits"]
            got = guts.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_CALLS | api.WANT_BEST)  # the fused K1
            assert guts.last_batch_was_fused
            _compare(got, want, is_ref, guts, f"20M world, {name}, fused calls + best")
            pk, woff = api.pack_residues(batch.residues, batch.offsets)
            got = guts.process_packed_batch(pk, woff, api.WANT_BEST)
            _compare(got, want, is_ref, guts, f"20M world, {name}, packed entry")
        guts.set_tuning(0)
        guts.close()
        chk.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)
