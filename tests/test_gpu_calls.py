"""GPU parity tests proper: the CUDA path, called through the C ABI (libckm.so), must agree bit for bit
with the plain-C oracle -- and with the reference's own object code when oracle/_ref is present."""
import os

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, synth

pytestmark = pytest.mark.gpu

ALL = api.WANT_CALLS | api.WANT_HITS | api.WANT_OTU | api.WANT_BEST
PARAM_SETS = [
    dict(),
    dict(order_constraint=1),
    dict(min_hits=3, max_gap=50),
    dict(min_hits=2, min_weighted_hits=20, max_gap=10),
    dict(order_constraint=1, min_hits=2, max_gap=600),
]


@pytest.fixture(scope="module")
def world(checkers):
    protos, sig, img = wl.small_world()
    orc = checkers.Oracle().open_image(img)
    names = synth.function_names(sig.n_functions)
    guts = api.KmerGuts(image=img, function_names=names)
    os.environ["CKM_FORCE_RAW_SLOTS"] = "1"
    guts_raw = api.KmerGuts(image=img, function_names=names)
    del os.environ["CKM_FORCE_RAW_SLOTS"]
    assert guts.slot_bytes == 16 and guts_raw.slot_bytes == 24
    assert guts.stream != 0 and guts.stream != guts_raw.stream  # every context works on a stream of its own, never the default one
    yield protos, sig, img, orc, guts, guts_raw
    guts.close()
    guts_raw.close()
    orc.close()


@pytest.mark.parametrize("prm", PARAM_SETS)
def test_call_batch_bit_exact(checkers, world, prm):
    protos, sig, _, orc, guts, guts_raw = world
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(2, protos, 3000))
    orc.set_params(**prm)
    want = orc.call_batch(batch, ALL)
    for g, nm in ((guts, "packed16"), (guts_raw, "raw24")):
        g.set_parameters(prm)
        got = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
        wl.assert_results_equal(got, want, f"cuda[{nm}] vs oracle {prm}")
        assert got["n_probes"] == want["n_probes"]
        assert got["n_hits"] == len(want["hits"])
    assert len(want["calls"]) > 100


@pytest.mark.parametrize("group", ["4", "8", "16"])
def test_group_probe_kernel_bit_exact(checkers, world, group):
    """probe_group_kernel (4 / 8 / 16 lanes per sequence, chosen for batches of short sequences) against the oracle: forced on
    the mixed edge + protein batch, with and without the occupancy bitmap, and picked by itself on short peptides."""
    protos, sig, img, orc, guts, guts_raw = world
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(4, protos, 3000))
    rng = np.random.default_rng(8)
    full = synth.make_proteins(14, protos, 4000)
    cuts = [full.seq(i)[int(a):int(a) + int(l)] for i, (a, l) in enumerate(zip(rng.integers(0, 200, full.n), rng.integers(0, 70, full.n)))]
    short = synth.batch_from_strings(cuts)  # mean ~35 residues, many under 9 (no window) and some empty
    os.environ["CKM_PROBE_GROUP"] = group
    try:
        for prm in (dict(), dict(order_constraint=1, min_hits=2, max_gap=30)):
            orc.set_params(**prm)
            for b in (batch, short):
                want = orc.call_batch(b, ALL)
                for g, nm in ((guts, "packed16"), (guts_raw, "raw24")):
                    g.set_parameters(prm)
                    got = g.process_aa_seq_batch(b.residues, b.offsets, ALL)
                    wl.assert_results_equal(got, want, f"group {group} [{nm}] {prm}")
                    assert got["n_probes"] == want["n_probes"] and got["n_hits"] == len(want["hits"])
        os.environ["CKM_OCCUPANCY_BITMAP"] = "1"
        g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
        g.set_parameters(prm)
        wl.assert_results_equal(g.process_aa_seq_batch(short.residues, short.offsets, ALL), orc.call_batch(short, ALL), "group + bitmap")
        g.close()
    finally:
        os.environ.pop("CKM_PROBE_GROUP", None)
        os.environ.pop("CKM_OCCUPANCY_BITMAP", None)
        orc.set_params()
        guts.set_default_parameters()
        guts_raw.set_default_parameters()
    # without the override the short batch selects a group kernel on its own (mean length <= 48), the proteins do not
    want = orc.call_batch(short, ALL)
    wl.assert_results_equal(guts.process_aa_seq_batch(short.residues, short.offsets, ALL), want, "auto group")
    assert len(want["hits"]) > 1000


def test_occupancy_bitmap_and_pipelined_host_path(checkers, world):
    """The two large-table mechanisms, forced on a small table: the L2-resident slot-occupancy bitmap (skips the
    DRAM read of empty slots) and the chunked two-stream host path used for find_best_call-only batches."""
    protos, sig, img, orc, _, _ = world
    os.environ.update(CKM_OCCUPANCY_BITMAP="1", CKM_PIPELINE_MIN_KB="0", CKM_PIPELINE_CHUNK_KB="64")
    try:
        for raw in ("0", "1"):
            os.environ["CKM_FORCE_RAW_SLOTS"] = raw
            g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
            assert g.has_occupancy_bitmap
            batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(6, protos, 4000))
            for prm in (dict(), dict(order_constraint=1, min_hits=3)):
                orc.set_params(**prm)
                g.set_parameters(prm)
                want = orc.call_batch(batch, ALL)
                wl.assert_results_equal(g.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, f"bitmap raw={raw} {prm}")
                # best-only -> pipelined path (~20 chunks on two streams); also with a rebased offsets array
                got = g.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_BEST)
                assert got["best"].tobytes() == want["best"].tobytes() and got["n_probes"] == want["n_probes"]
                k = 37
                sub_off = batch.offsets[k:]
                got2 = g.process_aa_seq_batch(batch.residues, sub_off, api.WANT_BEST)
                assert got2["best"].tobytes() == want["best"][k:].tobytes()
            g.close()
    finally:
        for k in ("CKM_OCCUPANCY_BITMAP", "CKM_PIPELINE_MIN_KB", "CKM_PIPELINE_CHUNK_KB", "CKM_FORCE_RAW_SLOTS"):
            os.environ.pop(k, None)


def test_each_flag_alone(checkers, world):
    """The handlers pass different subsets of (calls, hit_cb, otu_stats): every subset must agree."""
    protos, _, _, orc, guts, _ = world
    batch = synth.make_proteins(9, protos, 500)
    orc.set_params()
    guts.set_default_parameters()
    for flags in (api.WANT_CALLS, api.WANT_HITS, api.WANT_OTU, api.WANT_BEST, api.WANT_CALLS | api.WANT_OTU,
                  api.WANT_HITS | api.WANT_BEST):
        want = orc.call_batch(batch, flags)
        got = guts.process_aa_seq_batch(batch.residues, batch.offsets, flags)
        wl.assert_results_equal(got, want, f"flags={flags}")


def test_empty_and_degenerate_batches(checkers, world):
    protos, _, _, orc, guts, _ = world
    guts.set_default_parameters()
    orc.set_params()
    for seqs in ([], [b""], [b"", b"", b"A"], [b"ACDEFGHIK"]):
        batch = synth.batch_from_strings(seqs)
        want = orc.call_batch(batch, ALL)
        got = guts.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
        wl.assert_results_equal(got, want, f"degenerate {seqs}")


def test_against_reference_object_code(checkers, world, tmp_path):
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so absent")
    protos, sig, img, _, guts, _ = world
    d = str(tmp_path)
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    ref = checkers.Ref().open(d)
    ref.set_params()
    guts.set_default_parameters()
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(4, protos, 2000))
    want = ref.call_batch(batch, ALL)
    got = guts.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
    wl.assert_results_equal(got, want, "cuda vs reference object code", check_ambig_indices=False)
    assert [guts.best_function(r) for r in got["best"]] == want["best_function"]
    ref.close()


def test_open_from_directory_and_validation(checkers, world, tmp_path):
    """T2: ckm_open reads the reference's files unchanged and applies its three validations."""
    protos, sig, img, orc, _, _ = world
    d = str(tmp_path)
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 3)
    g = api.KmerGuts(kmer_dir=d)
    assert g.num_sigs == synth.bucket_count(len(sig.keys))
    assert g.function_at_index(5) == "function 5" and g.function_at_index(-1) == "INVALID_OFFSET"
    assert g.function_at_index(10**7) == "INVALID_OFFSET" and g.otu_at_index(2) == "otu 2"
    batch = synth.make_proteins(12, protos, 300)
    orc.set_params()
    wl.assert_results_equal(g.process_aa_seq_batch(batch.residues, batch.offsets, ALL), orc.call_batch(batch, ALL), "dir-open")
    g.close()
    for mutate in ("size", "version", "entry"):
        bad = img.copy()
        hdr = bad[:24].view(np.uint64)
        if mutate == "size":
            bad = bad[:-24]
        elif mutate == "version":
            hdr[2] = 2
        else:
            hdr[1] = 32
        with pytest.raises(api.CkmError) as ei:
            api.KmerGuts(image=bad)
        assert ei.value.code == -3
        assert orc.try_open_image(bad) == -3


def test_window_saturation_long_protein(checkers):
    """> 39998 stored hits in one run (kguts.cc:850-851): single-function image, one 45k-residue protein."""
    protos, sig, img = wl.small_world(seed=21, n_protos=200, n_sigs=58_000, n_functions=1, otu_mode="minus1")
    orc = checkers.Oracle().open_image(img)
    guts = api.KmerGuts(image=img)
    seq = synth.AA[protos.codes[:45_000]].tobytes()
    batch = synth.batch_from_strings([seq, seq[:500], seq[100:41_000]])
    for prm in (dict(), dict(order_constraint=1), dict(max_gap=3)):
        orc.set_params(**prm)
        guts.set_parameters(prm)
        want = orc.call_batch(batch, ALL)
        got = guts.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
        assert want["n_hits"] > 40_000
        wl.assert_results_equal(got, want, f"saturation {prm}")
    # the chunked host path decides per chunk whether the general scan kernel is needed
    os.environ.update(CKM_PIPELINE_MIN_KB="0", CKM_PIPELINE_CHUNK_KB="16")
    try:
        g2 = api.KmerGuts(image=img)
        big = wl.concat_batches(synth.make_proteins(5, protos, 300), wl.concat_batches(batch, synth.make_proteins(6, protos, 300)))
        orc.set_params()
        want = orc.call_batch(big, api.WANT_BEST)
        assert g2.process_aa_seq_batch(big.residues, big.offsets, api.WANT_BEST)["best"].tobytes() == want["best"].tobytes()
        g2.close()
    finally:
        os.environ.pop("CKM_PIPELINE_MIN_KB")
        os.environ.pop("CKM_PIPELINE_CHUNK_KB")
    guts.close()
    orc.close()


def test_randomised_parameters_and_shapes(checkers, world):
    """Seeded random sweeps over engine parameters, flag subsets and batch shapes (many tiny sequences, a few very
    long ones, high-bit bytes, runs of one residue), each against the oracle."""
    protos, sig, img, orc, guts, guts_raw = world
    rng = np.random.default_rng(2024)
    aa = synth.AA
    for trial in range(12):
        seqs = []
        for _ in range(int(rng.integers(1, 400))):
            kind = int(rng.integers(0, 7))
            p = int(rng.integers(0, protos.n))
            s = aa[protos.codes[int(protos.offsets[p]):int(protos.offsets[p + 1])]].copy()
            if kind == 0:
                s = s[: int(rng.integers(0, 30))]
            elif kind == 1:  # several prototypes glued: long, many runs
                parts = [aa[protos.codes[int(protos.offsets[q]):int(protos.offsets[q + 1])]] for q in rng.integers(0, protos.n, int(rng.integers(2, 30)))]
                s = np.concatenate(parts)
            elif kind == 2:
                s[rng.integers(0, len(s), int(rng.integers(1, 40)))] = rng.integers(0, 256, 1, dtype=np.uint8)[0] or 1
            elif kind == 3:
                s = np.full(int(rng.integers(1, 300)), aa[int(rng.integers(0, 20))], np.uint8)
            elif kind == 4:
                s = s[int(rng.integers(0, 200)):]
            elif kind == 5:
                s = np.concatenate([s[:100], s[:100], s[50:150]])
            seqs.append(s.tobytes())
        batch = synth.batch_from_strings(seqs)
        prm = dict(order_constraint=int(rng.integers(0, 2)), min_hits=int(rng.integers(1, 9)),
                   min_weighted_hits=int(rng.integers(0, 30)), max_gap=int(rng.choice([0, 1, 7, 50, 200, 100000])))
        flags = int(rng.integers(1, 16))
        orc.set_params(**prm)
        want = orc.call_batch(batch, flags)
        for g in (guts, guts_raw):
            g.set_parameters(prm)
            wl.assert_results_equal(g.process_aa_seq_batch(batch.residues, batch.offsets, flags), want, f"trial {trial} {prm} flags={flags}")
    guts.set_default_parameters()
    guts_raw.set_default_parameters()


def test_image_built_on_gpu_is_byte_identical(checkers, world):
    """N4: priority linear probing on the GPU reproduces the sequential insert_kmer loop's file bytes (the host builder is
    pinned to the reference's own builder in tests/test_oracle_vs_ref.py)."""
    protos, sig, img, orc, guts, _ = world
    nb = int(img[:8].view(np.uint64)[0])
    assert api.build_image_device(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt).tobytes() == img.tobytes()
    rng = np.random.default_rng(77)
    cases = []
    # duplicates, keys that are not inserted, and long collision chains: many keys congruent modulo the bucket count
    n, nbk = 40_000, 101533
    keys = rng.integers(0, 20**8, n).astype(np.uint64)
    keys[::7] = keys[3]                                  # the same key many times: each takes its own slot, in order
    keys[5::11] = (rng.integers(0, 1000, len(keys[5::11])).astype(np.uint64) * np.uint64(nbk) + np.uint64(17))  # one home bucket
    keys[9::13] = np.uint64(20**8 + 5)                   # > MAX_ENCODED: skipped
    cases.append((nbk, keys))
    # probe chains that wrap around the end of the table
    nbk2 = 3769
    k2 = (rng.integers(0, 50, 1500).astype(np.uint64) * np.uint64(nbk2) + np.uint64(nbk2) - rng.integers(1, 40, 1500).astype(np.uint64))
    cases.append((nbk2, k2))
    cases.append((3769, np.zeros(0, np.uint64)))
    for nbk_, k_ in cases:
        m = len(k_)
        fI, oI = rng.integers(0, 5000, m).astype(np.int32), rng.integers(-1, 300, m).astype(np.int32)
        avg, wt = rng.integers(0, 65536, m).astype(np.uint16), rng.random(m).astype(np.float32)
        for _ in range(3):  # whatever the thread schedule
            assert api.build_image_device(nbk_, k_, fI, oI, avg, wt).tobytes() == api.build_image(nbk_, k_, fI, oI, avg, wt).tobytes()
    with pytest.raises(api.CkmError):  # half full: refused like kguts.cc:213-216
        api.build_image_device(16, np.arange(9, dtype=np.uint64), [0] * 9, [0] * 9, [0] * 9, [0.0] * 9)
    # opened straight from the device-built table: same calls as the context opened on the image
    g2 = api.KmerGuts(built=(nb, sig.keys, sig.fI, sig.oI, sig.avg, sig.wt), function_names=synth.function_names(sig.n_functions))
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(21, protos, 2000))
    orc.set_params()
    wl.assert_results_equal(g2.process_aa_seq_batch(batch.residues, batch.offsets, ALL), orc.call_batch(batch, ALL), "ckm_open_built")
    assert g2.slot_bytes == guts.slot_bytes
    g2.close()


@pytest.mark.parametrize("chain", ["0", "1"])
def test_reference_hash_layout_and_own_table_agree(checkers, world, chain):
    """The default table is the library's own (a power of two of 16-byte slots under a multiplicative hash, filled from the image
    at load); CKM_REFERENCE_HASH=1 keeps the image's slot order and key % num_sigs.  Same answers, with and without the copy."""
    protos, sig, img, orc, guts, _ = world
    assert guts.table_buckets >= guts.num_sigs and guts.table_buckets & (guts.table_buckets - 1) == 0
    os.environ.update(CKM_REFERENCE_HASH="1", CKM_CHAIN=chain, CKM_OCCUPANCY_BITMAP="1")
    try:
        g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
    finally:
        for k in ("CKM_REFERENCE_HASH", "CKM_CHAIN", "CKM_OCCUPANCY_BITMAP"):
            os.environ.pop(k, None)
    try:
        assert g.table_buckets == g.num_sigs == guts.num_sigs and g.slot_bytes == 16
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(31, protos, 2500))
        orc.set_params()
        g.set_default_parameters()
        want = orc.call_batch(batch, ALL)
        wl.assert_results_equal(g.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, f"reference hash, chain={chain}")
        few = g.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_CALLS | api.WANT_BEST)
        assert g.last_batch_was_fused
        wl.assert_results_equal(few, {k: want[k] for k in ("call_offsets", "calls", "best")}, f"reference hash fused, chain={chain}")
    finally:
        g.close()


def test_kmers_the_reference_cannot_reach_stay_unreachable(checkers):
    """A hand-made image: signature k-mers moved two buckets past their home with an empty bucket in between.  lookup_hash_entry
    (kguts.cc:585-602) stops at the empty home bucket and never sees them; the library's own table (rehash_kernel re-inserts
    every k-mer of the image) must not find them either."""
    protos, sig, img = wl.small_world(seed=29, n_protos=120, n_sigs=30_000)
    img = img.copy()
    slots = np.frombuffer(img, api.SLOT_DT, offset=24)  # a writable view
    nb = len(slots)
    occ = slots["which_kmer"] <= 20**8
    moved = 0
    for s in np.flatnonzero(occ)[::7]:
        k = int(slots["which_kmer"][s])
        if k % nb != s or s + 2 >= nb or occ[s + 1] or occ[s + 2]:
            continue
        slots[s + 2] = slots[s]
        slots[s]["which_kmer"] = 20**8 + 1
        occ[s], occ[s + 2] = False, True
        moved += 1
    assert moved > 500
    orc = checkers.Oracle().open_image(img)
    batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(30, protos, 1500, mix=(1.0, 0.0, 0.0, 0.0), sub_rate=0.0))
    want = orc.call_batch(batch, ALL)
    for env in ({}, {"CKM_REFERENCE_HASH": "1"}, {"CKM_CHAIN": "1", "CKM_OCCUPANCY_BITMAP": "1"}):
        os.environ.update(env)
        try:
            g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
        finally:
            for k in env:
                os.environ.pop(k, None)
        try:
            wl.assert_results_equal(g.process_aa_seq_batch(batch.residues, batch.offsets, ALL), want, f"unreachable k-mers {env}")
            few = g.process_aa_seq_batch(batch.residues, batch.offsets, api.WANT_CALLS | api.WANT_BEST)
            wl.assert_results_equal(few, {k: want[k] for k in ("call_offsets", "calls", "best")}, f"unreachable k-mers fused {env}")
        finally:
            g.close()
    orc.close()
