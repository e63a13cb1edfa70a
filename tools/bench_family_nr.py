"""N2 measurement: family-mode start-up load (families.nr proteins -> k-mer -> family lists) through ckm_family_nr_* with
host buffers, beside the reference's thread_load / add_fam_mapping (oracle/_ref, one thread) or the C port on a bounded
sample.  python tools/bench_family_nr.py [n_proteins] [n_sigs]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from close_kmers_b200 import api, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
protos = synth.make_prototypes(4242, -(-n_sigs // 293) + 8, 300, 60.0)
sig = synth.make_signatures(protos, n_sigs, dedupe=True)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
fam = synth.make_families(7, sig)
chunk = 250_000
parts = [synth.make_proteins(300 + k, protos, min(chunk, n_prot - k * chunk)) for k in range(-(-n_prot // chunk))]
rng = np.random.default_rng(9)
fam_ids = [rng.integers(0, fam.n_fams, p.n).astype(np.uint32) for p in parts]
g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))


def load():
    g.family_nr_begin()
    for f, p in zip(fam_ids, parts):
        g.family_nr_add(f, p.residues, p.offsets)
    return g.family_nr_finish(fam.pgf, fam.plf, fam.function)


load()  # grows the work buffers
t0 = time.perf_counter()
nk, ne = load()
dt = time.perf_counter() - t0
out = dict(proteins=n_prot, residues=int(sum(int(p.offsets[-1]) for p in parts)), signature_kmers=len(sig.keys), table_kmers=nk,
           table_entries=ne, gpu_e2e_s=dt, gpu_proteins_per_s=n_prot / dt)
import cpu_checkers as cc
cc.ensure_built()
m = min(parts[0].n, 40_000)
sub = synth.Batch(parts[0].residues[: int(parts[0].offsets[m])], parts[0].offsets[: m + 1])
g.family_nr_begin()
g.family_nr_add(fam_ids[0][:m], sub.residues, sub.offsets)
k2, e2 = g.family_nr_finish(fam.pgf, fam.plf, fam.function)
got = g.family_export(k2, e2)
if os.path.exists(cc.REF_SO):
    d = tempfile.mkdtemp(prefix="ckm_nr_")
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 0)
    ref = cc.Ref().open(d)
    ref.set_params()
    t0 = time.perf_counter()
    ref.family_nr_add(fam_ids[0][:m], sub)
    dtc = time.perf_counter() - t0
    want = ref.family_table()
    kind = "reference"
else:
    orc = cc.Oracle().open_image(img)
    t0 = time.perf_counter()
    want = orc.family_nr_build([(fam_ids[0][:m], sub)])
    dtc = time.perf_counter() - t0
    kind = "port"
out.update(cpu_kind=kind, cpu_sample=m, cpu_proteins_per_s_1thread=m / dtc,
           parity_on_sample=bool(all(np.array_equal(a, b) for a, b in zip(got, want))))
print(json.dumps(out), flush=True)
