import sys, os, numpy as np, torch, json
sys.path[:0]=['/root/repo','/root/repo/tests']
from close_kmers_b200 import api, parallel, synth
n=int(sys.argv[1]) if len(sys.argv)>1 else 5000
protos = synth.make_prototypes(777, max(n // 10, 8), 300, 60.0)
sig = synth.make_signatures(protos, min(1_000_000, int(protos.offsets[-1]) - 8 * protos.n), dedupe=True)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
batch = synth.make_proteins_parallel(778, protos, n, mix=(0.9, 0.1, 0.0, 0.0))
eids = np.arange(batch.n, dtype=np.uint32)
g = api.KmerGuts(image=img)
job = parallel.MatrixJob(g, eids, batch, 0, 1)
for _ in range(3):
    merged, stats = job.run()
print(json.dumps(stats["phase_ms"]), len(merged), stats["walked"])
g2 = api.KmerGuts(image=img)
g2.postings_add(eids, batch.residues, batch.offsets)
whole = api.merge_pairs(g2.matrix_rows(eids, batch.residues, batch.offsets))
print("equal:", whole.tobytes()==merged.tobytes(), len(whole))
