"""Generates tests/golden/golden_v1.npz from the REFERENCE'S OWN OBJECT CODE (oracle/_ref/libckm_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these pin the path instead: a small signature image
written by the reference builder, seeded proteins / reads, and every output the reference produces for them
(calls, hit lists, OTU stats, best calls and their names, /query /add /matrix /fq_lookup response text, family
matches with exact f32 scores).  tests/test_golden.py replays them against the oracle (CPU) and the CUDA path (GPU).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_checkers as cc  # noqa: E402
import workloads as wl  # noqa: E402
from close_kmers_b200 import synth  # noqa: E402

PARAM_SETS = [dict(), dict(order_constraint=1), dict(min_hits=3, max_gap=50), dict(min_hits=2, min_weighted_hits=20, max_gap=10)]


def main():
    cc.ensure_built()
    protos = synth.make_prototypes(101, 12, 220, 30.0)
    sig = synth.make_signatures(protos, 2400, n_functions=5, otu_mode="mixed")
    fam = synth.make_families(7, sig, fams_per_function=3)
    d = tempfile.mkdtemp(prefix="ckm_golden_")
    ref = cc.Ref()
    ref.build_image(d, synth.bucket_count(len(sig.keys)), sig)  # the reference's own builder
    synth.write_index_files(d, sig.n_functions, 12)
    image = np.fromfile(os.path.join(d, "kmer.table.mem_map"), np.uint8)
    ref.open(d)
    proteins = wl.concat_batches(wl.edge_batch(protos, seed=9), synth.make_proteins(5, protos, 60, mix=(0.6, 0.3, 0.05, 0.05)))
    ids = [f"fig|1.1.peg.{i}" for i in range(proteins.n)]
    out = dict(image=image, n_functions=sig.n_functions, n_otus=12, residues=proteins.residues, offsets=proteins.offsets,
               ids=np.array(ids))
    flags = cc.WANT_CALLS | cc.WANT_HITS | cc.WANT_OTU | cc.WANT_BEST
    for k, prm in enumerate(PARAM_SETS):
        ref.set_params(**prm)
        r = ref.call_batch(proteins, flags)
        out[f"p{k}_params"] = np.array([prm.get("order_constraint", 0), prm.get("min_hits", 5), prm.get("min_weighted_hits", 0),
                                        prm.get("max_gap", 200)])
        for key in ("call_offsets", "calls", "hit_offsets", "hits", "otu_offsets", "otus", "otus_sorted", "best"):
            out[f"p{k}_{key}"] = r[key]
        out[f"p{k}_best_function"] = np.array(r["best_function"])
    ref.set_params()
    for details, fbc in ((0, 0), (1, 0), (0, 1)):
        out[f"query_text_{details}{fbc}"] = np.array(ref.query_text(ids, proteins, details, fbc))
    # family voting + fastq
    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    for key in ("kmers", "fam_off", "fam_ids", "fam_func_sid", "fam_pgf", "func_sid"):
        out[f"fam_{key}"] = getattr(fam, key)
    out["fam_pgf_names"] = np.array(fam.pgf_names)
    for key in ("pgf", "plf", "function"):
        out[f"fam_{key}"] = np.array(getattr(fam, key))
    fb = ref.family_batch(proteins)
    for key in ("gscore", "lscore", "score"):
        out[f"family_{key}"] = fb[key]
    for key in ("gfam", "lfam", "function"):
        out[f"family_{key}"] = np.array(fb[key])
    from test_oracle_family_fq_matrix import DNA_EDGE
    reads = wl.concat_batches(synth.batch_from_strings(DNA_EDGE), synth.make_reads(3, protos, 80))
    rids = [f"read{i}" for i in range(reads.n)]
    out.update(read_bases=reads.residues, read_offsets=reads.offsets, read_ids=np.array(rids))
    out["fq_text"] = np.array(ref.fq_text(rids, reads))
    out["six_frames"] = np.array([ref.six_frames(reads.seq(i)) if len(reads.seq(i)) >= 3 else "" for i in range(reads.n)])
    # /add (two chunks, second silent) then /matrix
    ref.mapping_new()
    half = proteins.n // 2
    adds = []
    for lo, hi, silent in ((0, half, 0), (half, proteins.n, 1)):
        part = synth.Batch(proteins.residues[int(proteins.offsets[lo]):int(proteins.offsets[hi])],
                           proteins.offsets[lo:hi + 1] - proteins.offsets[lo])
        adds.append(ref.add_text(ids[lo:hi], part, silent))
    out["add_text_0"], out["add_text_1"], out["add_split"] = np.array(adds[0]), np.array(adds[1]), np.array(half)
    out["matrix_text"] = np.array(ref.matrix_text(ids, proteins))
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", proteins.n, "proteins,", reads.n, "reads,",
          len(out["p0_calls"]), "calls,", len(out["p0_hits"]), "hits")


if __name__ == "__main__":
    main()
