"""N1 measurement: proteins/s through the running kser_b200 server over loopback HTTP (FASTA body in, response text out),
for /query in its three output modes.  python tools/bench_server.py [n_proteins] [n_sigs] [connections]"""
import json
import os
import socket
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from close_kmers_b200 import api, build, synth

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n_sigs = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
n_conn = int(sys.argv[3]) if len(sys.argv) > 3 else 4
protos = synth.make_prototypes(4242, -(-n_sigs // 293) + 8, 300, 60.0)
sig = synth.make_signatures(protos, n_sigs)
img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
d = tempfile.mkdtemp(prefix="ckm_srv_")
api.save_kmer_hash_table(img, d)
synth.write_index_files(d, sig.n_functions, 0)
batch = synth.make_proteins(5, protos, n_prot, mix=(0.9, 0.1, 0.0, 0.0))
res = batch.residues
off = batch.offsets.astype(np.int64)
parts = []
for i in range(batch.n):
    s = res[off[i]:off[i + 1]].tobytes()
    parts.append(b">fig|83333.1.peg.%d\n" % i + b"\n".join(s[k:k + 60] for k in range(0, len(s), 60)) + b"\n")
build.build()
pf = os.path.join(d, "port")
proc = subprocess.Popen([os.path.join(os.path.dirname(build.LIB), "kser_b200"), "--listen-port-file", pf, "0", d],
                        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
while not (os.path.exists(pf) and open(pf).read().strip()):
    if proc.poll() is not None:
        sys.exit("server died")
    time.sleep(0.1)
port = int(open(pf).read())


def request(path, body):
    s = socket.create_connection(("127.0.0.1", port))
    head = b"POST %s HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % (path.encode(), len(body))
    t = threading.Thread(target=lambda: s.sendall(head + body))
    t.start()
    n = 0
    while True:
        b = s.recv(1 << 22)
        if not b:
            break
        n += len(b)
    t.join()
    s.close()
    return n


def run(path, conns):
    shard = -(-len(parts) // conns)
    bodies = [b"".join(parts[k * shard:(k + 1) * shard]) for k in range(conns)]
    out = [0] * conns
    def one(k):
        out[k] = request(path, bodies[k])
    times = []
    for _ in range(2 if "details" in path else 7):  # single runs of ~0.1 s with a Python client are noisy: repeat
        t0 = time.perf_counter()
        ts = [threading.Thread(target=one, args=(k,)) for k in range(conns)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        times.append(time.perf_counter() - t0)
    times.sort()
    dt, med = times[0], times[len(times) // 2]
    return dict(path=path, connections=conns, best_seconds=dt, proteins_per_s=n_prot / dt, median_proteins_per_s=n_prot / med,
                runs=len(times), request_mb=sum(map(len, bodies)) / 1e6, response_mb=sum(out) / 1e6)


results = []
request("/query?find_best_call=1", b"".join(parts[:2000]))  # warm up
for path in ("/query?find_best_call=1", "/query", "/query?details=1"):
    for conns in (1, n_conn):
        results.append(run(path, conns))
        print(json.dumps(results[-1]), flush=True)
request_quit = socket.create_connection(("127.0.0.1", port))
request_quit.sendall(b"GET /quit HTTP/1.1\r\n\r\n")
request_quit.recv(1000)
proc.wait(timeout=60)
