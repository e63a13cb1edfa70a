"""Randomised parity stress: many seeded rounds of GPU vs the plain-C oracle over everything the C ABI computes (calls, hits,
OTU maps, best calls, family matches and score lists, fastq best frames), with random engine parameters, sequence length
mixes (proteins, peptides for the group probe kernels, junk) and family-list shapes.  python tools/stress_parity.py [rounds]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import cpu_checkers as cc
import workloads as wl
from close_kmers_b200 import api, synth

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cc.ensure_built()
ALL = api.WANT_CALLS | api.WANT_HITS | api.WANT_OTU | api.WANT_BEST
t_start = time.time()
stats = dict(rounds=0, proteins=0, hits=0, calls=0, family_entries=0, reads=0, hits_from_copy=0, fused_batches=0)
for rnd in range(rounds):
    rng = np.random.default_rng(1000 + rnd)
    mean_len = int(rng.choice([60, 150, 300, 600]))
    protos, sig, img = wl.small_world(seed=50 + rnd, n_protos=int(rng.integers(100, 500)), n_sigs=int(rng.integers(20_000, 120_000)),
                                      n_functions=int(rng.choice([3, 40, 400])), otu_mode=str(rng.choice(["mixed", "minus1"])),
                                      mean_len=mean_len, sd=float(rng.choice([0.0, 40.0])))
    if rnd % 3 == 1:  # signatures a sparse subset of the windows, a share of them of another function (stray hits inside runs)
        sig = synth.make_signatures_sparse(protos, len(sig.keys), keep=float(rng.choice([0.4, 0.7, 1.0])), foreign=float(rng.choice([0.05, 0.2])),
                                           n_functions=sig.n_functions, seed=rnd)
        img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    fam = synth.make_families(rnd, sig, fams_per_function=int(rng.choice([2, 4, 40])), max_list=int(rng.choice([3, 8, 40])),
                              coverage=float(rng.choice([0.5, 0.9, 1.0])))
    if rnd % 3 == 0:  # wide lists over many families: the vote kernel's overflow / global-scratch paths
        cnt = rng.integers(1, 12, len(fam.kmers))
        fam.fam_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        owner = np.repeat(np.arange(len(cnt)), cnt)
        rank = np.arange(int(fam.fam_off[-1])) - fam.fam_off[:-1].astype(np.int64)[owner]
        fam.fam_ids = ((rng.integers(0, fam.n_fams, len(cnt))[owner] + rank * 7) % fam.n_fams).astype(np.uint32)
        # lists must hold distinct ids: 7 is coprime to every n_fams used here unless n_fams % 7 == 0
        if fam.n_fams % 7 == 0 or fam.n_fams < 12 * 7:
            fam.fam_ids = ((rng.integers(0, fam.n_fams, len(cnt))[owner] + rank) % fam.n_fams).astype(np.uint32)
            if fam.n_fams < 12:
                continue
    os.environ["CKM_OCCUPANCY_BITMAP"] = str(rnd % 2)
    os.environ["CKM_FORCE_RAW_SLOTS"] = str((rnd // 2) % 2)
    os.environ["CKM_CHAIN"] = str((rnd // 4) % 2)  # neighbour copy + hinted probe (packed slots only)
    g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
    g.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    orc = cc.Oracle().open_image(img)
    orc.family_load(fam)
    prm = dict(order_constraint=int(rng.integers(0, 2)), min_hits=int(rng.integers(1, 8)), min_weighted_hits=int(rng.integers(0, 3)),
               max_gap=int(rng.choice([20, 200, 1000])))
    g.set_parameters(prm)
    orc.set_params(**prm)
    full = synth.make_proteins(rnd, protos, int(rng.integers(500, 3000)))
    pieces = [full.seq(i) for i in range(full.n)]
    if rnd % 2:  # peptides: random cuts, mean a few dozen residues
        pieces = [s[int(a):int(a) + int(l)] for s, a, l in zip(pieces, rng.integers(0, 100, full.n), rng.integers(0, int(rng.choice([30, 60, 120])), full.n))]
    batch = wl.concat_batches(wl.edge_batch(protos), synth.batch_from_strings(pieces))
    want = orc.call_batch(batch, ALL)
    got = g.process_aa_seq_batch(batch.residues, batch.offsets, ALL)
    from_copy = g.chain_info["hits_from_copy"]
    wl.assert_results_equal(got, want, f"round {rnd} calls {prm}")
    assert got["n_probes"] == want["n_probes"]
    # the same batch asking for calls and best calls only: the scan runs inside K1 (probe_pc_kernel) unless the order constraint
    # is on; through the ASCII and the packed entry point, and on a world whose signatures are a sparse, mixed-function subset
    for flags in (api.WANT_CALLS | api.WANT_BEST, api.WANT_BEST):
        few = g.process_aa_seq_batch(batch.residues, batch.offsets, flags)
        wl.assert_results_equal(few, {k: want[k] for k in ("call_offsets", "calls", "best") if k in few}, f"round {rnd} fused flags {flags} {prm}")
        stats["fused_batches"] += int(g.last_batch_was_fused)
    pk, woff = api.pack_residues(batch.residues, batch.offsets)
    wl.assert_results_equal(g.process_packed_batch(pk, woff, api.WANT_BEST), {"best": want["best"]}, f"round {rnd} packed entry")
    sc, so = orc.family_scores(batch)
    fs = g.family_scores(batch.residues, batch.offsets)
    assert np.array_equal(fs["score_offsets"], so), f"round {rnd} score offsets"
    for f in ("id", "hit_count", "weighted_total"):
        assert np.array_equal(fs["scores"][f], sc[f]), f"round {rnd} scores {f}"
    wl.assert_family_records_equal(fs["matches"], orc.family_batch(batch), f"round {rnd} family matches")
    reads = synth.make_reads(rnd, protos, int(rng.integers(200, 1500)), read_len=int(rng.choice([75, 150, 250])))
    g.set_default_parameters()
    orc.set_params()
    wl.assert_fq_records_equal(g.fq_batch(reads.residues, reads.offsets), orc.fq_batch(reads), f"round {rnd} fq")
    stats["rounds"] += 1
    stats["proteins"] += batch.n
    stats["hits"] += len(want["hits"])
    stats["calls"] += len(want["calls"])
    stats["family_entries"] += len(sc)
    stats["reads"] += reads.n
    stats["hits_from_copy"] += from_copy
    g.close()
    orc.close()
stats["seconds"] = time.time() - t_start
stats["ok"] = True
print(json.dumps(stats), flush=True)
