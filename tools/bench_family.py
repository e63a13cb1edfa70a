"""Family voting on proteins (SURVEY 8 rows F1-F3, and N3's score lists): FamilyMapper::find_best_family_match for a batch
through ckm_family_batch / ckm_family_scores with host buffers, beside the reference's FamilyMapper (oracle/_ref, one thread)
or the C port on a bounded sample.  python tools/bench_family.py [n_proteins] [n_sigs]"""
import ctypes as C
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
res = np.concatenate([p.residues for p in parts])
off = np.concatenate([[0]] + [p.offsets[1:].astype(np.uint64) + np.uint64(sum(int(q.offsets[-1]) for q in parts[:k]))
                               for k, p in enumerate(parts)]).astype(np.uint64)
g = api.KmerGuts(image=img, function_names=synth.function_names(sig.n_functions))
g.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
L = api.lib()


def timed(call):
    call()  # grows the work buffers
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        assert call() == 0
        ts.append(time.perf_counter() - t0)
    return min(ts)


m_ptr = C.c_void_p()
dt_batch = timed(lambda: L.ckm_family_batch(g._h, res.ctypes.data, off.ctypes.data, n_prot, C.byref(m_ptr)))
L.ckm_family_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(api.FamilyScoresC)]
fs = api.FamilyScoresC()
dt_scores = timed(lambda: L.ckm_family_scores(g._h, res.ctypes.data, off.ctypes.data, n_prot, C.byref(fs)))
n_scores = int(api._arr(fs.score_offsets, n_prot + 1, np.dtype("<u8"))[-1])
out = dict(proteins=n_prot, residues=int(off[-1]), signature_kmers=len(sig.keys), family_batch_s=dt_batch,
           family_batch_proteins_per_s=n_prot / dt_batch, family_scores_s=dt_scores, family_scores_proteins_per_s=n_prot / dt_scores,
           score_entries=n_scores)
import cpu_checkers as cc
import workloads as wl
cc.ensure_built()
m = min(parts[0].n, 20_000)
sub = synth.Batch(parts[0].residues[: int(parts[0].offsets[m])], parts[0].offsets[: m + 1])
got = g.find_best_family_match_batch(sub.residues, sub.offsets)
if os.path.exists(cc.REF_SO):
    d = tempfile.mkdtemp(prefix="ckm_fam_")
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 0)
    ref = cc.Ref().open(d)
    ref.set_params()
    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    t0 = time.perf_counter()
    want = ref.family_batch(sub)
    dtc = time.perf_counter() - t0
    wl.assert_family_equal(got, want, fam, synth.function_names(sig.n_functions), "cuda vs reference")
    kind = "reference"
else:
    orc = cc.Oracle().open_image(img)
    orc.family_load(fam)
    t0 = time.perf_counter()
    want = orc.family_batch(sub)
    dtc = time.perf_counter() - t0
    wl.assert_family_records_equal(got, want, "cuda vs oracle")
    kind = "port"
out.update(cpu_kind=kind, cpu_sample=m, cpu_proteins_per_s_1thread=m / dtc, parity_on_sample=True)
print(json.dumps(out), flush=True)
