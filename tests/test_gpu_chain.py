"""probe_chain_kernel (K1 through the neighbour-ordered copy of the table, csrc/ckm_chain.cuh) against the oracle and
against plain hash probing.  The copy only changes WHERE a probe is answered from, never the answer: every test here
demands bit-identical hits, calls, OTU maps and best calls."""
import os

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

pytestmark = pytest.mark.gpu

ALL = api.WANT_CALLS | api.WANT_HITS | api.WANT_OTU | api.WANT_BEST
NO_CHAIN = api.TUNE_PLAIN_PROBE | api.TUNE_NO_FALLBACK
FUSED_FLAGS = api.WANT_CALLS | api.WANT_BEST


def _open(img, names, bitmap="1", chain="1"):
    os.environ.update(CKM_CHAIN=chain, CKM_OCCUPANCY_BITMAP=bitmap)
    try:
        return api.KmerGuts(image=img, function_names=names)
    finally:
        os.environ.pop("CKM_CHAIN", None)
        os.environ.pop("CKM_OCCUPANCY_BITMAP", None)


def _codes(s: bytes) -> np.ndarray:
    lut = np.full(256, 255, np.uint8)
    lut[synth.AA] = np.arange(20, dtype=np.uint8)
    c = lut[np.frombuffer(s, np.uint8)]
    assert (c < 20).all()
    return c


def tricky_world(seed=21):
    """Prototypes that stress the chain builder: shared segments between prototypes of different functions (duplicate
    k-mers, inserted twice: dedupe=False), three functions only (several same-function successors), homopolymers and
    short-period repeats (self-successors and cycles), and a prototype that revisits its own k-mers."""
    rng = np.random.default_rng(seed)
    base = synth.make_prototypes(seed, 300, 260, 40.0)
    parts = [base.codes[int(base.offsets[i]):int(base.offsets[i + 1])].copy() for i in range(base.n)]
    for k in range(0, 120, 3):  # copy a 30..60-residue segment of prototype k into prototype k+1 (another function)
        a, b = parts[k], parts[k + 1]
        L = int(rng.integers(30, 60))
        sa, sb = int(rng.integers(0, len(a) - L)), int(rng.integers(0, len(b) - L))
        b[sb:sb + L] = a[sa:sa + L]
    parts += [_codes(b"A" * 40), _codes(b"AC" * 30), _codes(b"ACD" * 25), _codes(b"ACDE" * 20), _codes(b"W" * 9),
              _codes(b"MKVLAAGIW" * 12), _codes(b"ACDEFGHIKLMNPQRSTVWY" * 6)]
    p = parts[7]
    parts.append(np.concatenate([p[:100], p[40:160], p[10:90]]))  # revisits its own windows
    offsets = np.zeros(len(parts) + 1, np.int64)
    np.cumsum([len(x) for x in parts], out=offsets[1:])
    protos = synth.Prototypes(np.concatenate(parts), offsets)
    n_win = int(sum(max(0, len(x) - 7) for x in parts))
    sig = synth.make_signatures(protos, n_win, n_functions=3, otu_mode="mixed", dedupe=False)
    nb = synth.bucket_count(len(sig.keys))
    img = api.build_image(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    return protos, sig, img


def indel_batch(protos, seed, n):
    """Prototype-derived queries with insertions and deletions (the chain offset changes mid-protein), inversions of
    segment order and tandem duplications."""
    rng = np.random.default_rng(seed)
    aa = synth.AA
    seqs = []
    for k in range(n):
        p = int(rng.integers(0, protos.n))
        s = aa[protos.codes[int(protos.offsets[p]):int(protos.offsets[p + 1])]].copy()
        mode = k % 5
        if len(s) < 60:
            seqs.append(s.tobytes())
            continue
        c = int(rng.integers(20, len(s) - 20))
        if mode == 0:
            s = np.concatenate([s[:c], s[c + int(rng.integers(1, 9)):]])
        elif mode == 1:
            s = np.concatenate([s[:c], aa[rng.integers(0, 20, int(rng.integers(1, 9)))], s[c:]])
        elif mode == 2:
            s = np.concatenate([s[c:], s[:c]])
        elif mode == 3:
            s = np.concatenate([s[:c], s[c - 15:c], s[c - 15:]])
        else:
            q = int(rng.integers(0, protos.n))
            t = aa[protos.codes[int(protos.offsets[q]):int(protos.offsets[q + 1])]]
            s = np.concatenate([s[:c], t[: len(t) // 2], s[c:]])
        sub = rng.random(len(s)) < 0.03
        s[sub] = aa[rng.integers(0, 20, int(sub.sum()))]
        seqs.append(s.tobytes())
    return synth.batch_from_strings(seqs)


def _compare(g, orc, batch, what, prms=(dict(), dict(order_constraint=1, min_hits=2, max_gap=600))):
    for prm in prms:
        orc.set_params(**prm)
        g.set_parameters(prm)
        want = orc.call_batch(batch, ALL)
        g.set_tuning(api.TUNE_NO_FALLBACK)
        got = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
        wl.assert_results_equal(got, want, f"{what} chain {prm}")
        assert got["n_probes"] == want["n_probes"] and got["n_hits"] == len(want["hits"])
        from_copy = g.chain_info["hits_from_copy"]
        # the same request served by the fused K1 (calls + best call without hit records, ckm_warp_scan.cuh)
        fused = g.process_aa_seq_batch(batch.residues, batch.offsets, FUSED_FLAGS)
        wl.assert_results_equal(fused, {k: want[k] for k in ("call_offsets", "calls", "best")}, f"{what} fused {prm}")
        assert fused["n_probes"] == want["n_probes"] and fused["n_hits"] == len(want["hits"])
        # CKM_EXPERIMENTS builds only: the other builds of probe_hint_kernel (block shape, hit payload in registers or shared
        # memory; ckm_set_tuning bits 16-18), without the evict_first policy / L2 prefetches, and the walking
        # probe_chain_kernel (bits 7, 6)
        for tuning in ((1 << 16, 2 << 16, 3 << 16, 4 << 16, 5 << 16, 1 | 0x80000, 128, 64) if api.experiments_enabled() else ()):
            g.set_tuning(tuning)
            other = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
            wl.assert_results_equal(other, want, f"{what} tuning {tuning:#x} {prm}")
            assert g.chain_info["hits_from_copy"] > 0 or from_copy == 0
        g.set_tuning(NO_CHAIN)
        plain = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
        wl.assert_results_equal(plain, want, f"{what} plain {prm}")
        assert g.chain_info["hits_from_copy"] == 0
        fused = g.process_aa_seq_batch(batch.residues, batch.offsets, FUSED_FLAGS)
        wl.assert_results_equal(fused, {k: want[k] for k in ("call_offsets", "calls", "best")}, f"{what} plain fused {prm}")
        g.set_tuning(0)
    orc.set_params()
    g.set_default_parameters()
    return from_copy, len(want["hits"])


@pytest.mark.parametrize("bitmap", ["1", "0"])
def test_chain_probe_bit_exact_small_world(checkers, bitmap):
    protos, sig, img = wl.small_world()
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions), bitmap=bitmap)
    try:
        info = g.chain_info
        assert info["entries"] == len(sig.keys)  # every k-mer of a deduplicated image is reachable
        assert 0 < info["chains"] < info["entries"] // 50, info  # prototypes of 300 residues: chains of ~290
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 3000))
        from_copy, n_hits = _compare(g, orc, batch, f"small world bitmap={bitmap}")
        assert from_copy > 0.8 * n_hits, (from_copy, n_hits)  # mutated prototypes follow their chains
        _compare(g, orc, indel_batch(protos, 3, 600), "indels", prms=(dict(),))
        # best-only batches take the chunked two-stream host path
        os.environ.update(CKM_PIPELINE_MIN_KB="0", CKM_PIPELINE_CHUNK_KB="64")
        try:
            want = orc.call_batch(batch, ALL)
            got = g.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_BEST)
            assert got["best"].tobytes() == want["best"].tobytes() and got["n_probes"] == want["n_probes"]
        finally:
            os.environ.pop("CKM_PIPELINE_MIN_KB", None)
            os.environ.pop("CKM_PIPELINE_CHUNK_KB", None)
    finally:
        g.close()
        orc.close()


def test_chain_probe_duplicates_cycles_and_ambiguous_successors(checkers):
    protos, sig, img = tricky_world()
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions))
    try:
        info = g.chain_info
        n_distinct = len(np.unique(sig.keys))
        assert n_distinct < len(sig.keys)  # the image really holds unreachable duplicates
        assert info["entries"] == n_distinct, (info, n_distinct)
        aa = synth.AA
        own = synth.batch_from_strings([aa[protos.codes[int(protos.offsets[i]):int(protos.offsets[i + 1])]].tobytes()
                                        for i in range(protos.n)])
        low = synth.batch_from_strings([b"A" * 300, b"AC" * 150, b"CA" * 150, b"ACD" * 90, b"ACDE" * 70, b"W" * 17,
                                        b"MKVLAAGIW" * 30, b"ACDEFGHIKLMNPQRSTVWY" * 20, b"A" * 129 + b"C" * 9 + b"AC" * 70])
        batch = wl.concat_batches(wl.concat_batches(wl.edge_batch(protos), own), low)
        batch = wl.concat_batches(batch, synth.make_proteins(5, protos, 2500))
        _compare(g, orc, batch, "tricky world")
        _compare(g, orc, indel_batch(protos, 4, 800), "tricky world indels", prms=(dict(min_hits=3, max_gap=50),))
    finally:
        g.close()
        orc.close()


def test_chain_probe_hits_without_chains_fall_back(checkers):
    """Signature k-mers with no neighbours (every k-mer its own chain) and random queries: anchors keep hitting without
    ever being followed, so proteins switch themselves back to plain hash probing -- same results."""
    rng = np.random.default_rng(9)
    protos = synth.make_prototypes(31, 300, 300, 0.0)
    sig = synth.make_signatures(protos, 80_000, otu_mode="mixed")
    keep = np.arange(0, len(sig.keys), 9)  # every ninth window: no two kept k-mers overlap by seven residues
    nb = synth.bucket_count(len(keep))
    img = api.build_image(nb, sig.keys[keep], sig.fI[keep], sig.oI[keep], sig.avg[keep], sig.wt[keep])
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions))
    try:
        info = g.chain_info
        assert info["entries"] == len(keep)
        assert info["chains"] > 0.99 * len(keep), info
        batch = synth.make_proteins(int(rng.integers(100)), protos, 2000)
        from_copy, n_hits = _compare(g, orc, batch, "isolated k-mers")
        assert n_hits > 10_000 and from_copy < 0.05 * n_hits
    finally:
        g.close()
        orc.close()


def test_chain_copy_is_shared_by_clones_and_skipped_for_raw_slots(checkers):
    protos, sig, img = wl.small_world(seed=4, n_protos=100, n_sigs=20_000)
    names = synth.function_names(sig.n_functions)
    orc = checkers.Oracle().open_image(img)
    g = _open(img, names)
    os.environ["CKM_FORCE_RAW_SLOTS"] = "1"
    try:
        raw = _open(img, names)
    finally:
        del os.environ["CKM_FORCE_RAW_SLOTS"]
    try:
        assert raw.slot_bytes == 24 and raw.chain_info["entries"] == 0
        batch = synth.make_proteins(12, protos, 800)
        want = orc.call_batch(batch, ALL)
        wl.assert_results_equal(raw.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, "raw slots, no copy")
        clone = g.clone()
        try:
            assert clone.chain_info["entries"] == g.chain_info["entries"] > 0
            wl.assert_results_equal(clone.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, "clone over the copy")
            assert clone.chain_info["hits_from_copy"] > 0
        finally:
            clone.close()
    finally:
        raw.close()
        g.close()
        orc.close()


def test_chain_probe_long_proteins_reload_hints(checkers):
    """Proteins longer than 32 hint segments (2048 windows at one sample per 64): probe_hint_kernel reloads its hint
    registers every 32 segments; concatenated prototypes also change chain every few hundred windows."""
    protos, sig, img = wl.small_world(seed=8, n_protos=150, n_sigs=40_000)
    orc = checkers.Oracle().open_image(img)
    g = _open(img, synth.function_names(sig.n_functions))
    try:
        rng = np.random.default_rng(17)
        aa = synth.AA
        seqs = []
        for n_parts in (4, 5, 9, 17, 40, 3, 130):
            parts = []
            for _ in range(n_parts):
                p = int(rng.integers(0, protos.n))
                parts.append(aa[protos.codes[int(protos.offsets[p]):int(protos.offsets[p + 1])]])
            s = np.concatenate(parts).copy()
            sub = rng.random(len(s)) < 0.04
            s[sub] = aa[rng.integers(0, 20, int(sub.sum()))]
            seqs.append(s.tobytes())
        for cut in (1031, 1033, 2055, 2057, 4103, 4105):  # one window short of / past a reload for 32-, 64- and 128-window segments
            seqs.append(seqs[3][:cut])
        batch = wl.concat_batches(synth.batch_from_strings(seqs), synth.make_proteins(6, protos, 300))
        assert max(len(x) for x in seqs) > 30_000
        from_copy, n_hits = _compare(g, orc, batch, "long proteins")
        assert from_copy > 0.8 * n_hits, (from_copy, n_hits)
    finally:
        g.close()
        orc.close()


def test_chain_forced_on_empty_and_tiny_tables(checkers):
    """CKM_CHAIN=1 on a table without a k-mer (no copy is built) and on one with fewer than 64 buckets (fast_mod35 does
    not apply: the kernels fall back to the 64-bit reduction)."""
    protos, sig, _ = wl.small_world(seed=5, n_protos=20, n_sigs=2_000)
    batch = synth.make_proteins(3, protos, 200)
    z64, z32, z16, zf = np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.uint16), np.zeros(0, np.float32)
    for nb, k in ((3769, 0), (61, 20)):
        img = api.build_image(nb, sig.keys[:k] if k else z64, sig.fI[:k] if k else z32, sig.oI[:k] if k else z32,
                              sig.avg[:k] if k else z16, sig.wt[:k] if k else zf)
        orc = checkers.Oracle().open_image(img)
        g = _open(img, synth.function_names(sig.n_functions))
        try:
            assert g.chain_info["entries"] == k
            want = orc.call_batch(batch, ALL)
            got = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
            wl.assert_results_equal(got, want, f"{k} k-mers in {nb} buckets")
            assert got["n_probes"] == want["n_probes"]
        finally:
            g.close()
            orc.close()
