"""GPU parity for the "next" rows N1 / N3: the /lookup computation through the C ABI, and the kser_b200 server end to end
(start-up family load from families.dat + families.nr on the GPU, every route, bytes on the wire) against the reference's
handler loops run by oracle/_ref on the same inputs."""
import dataclasses
import gzip
import os
import signal
import socket
import subprocess
import threading
import time

import numpy as np
import pytest

import workloads as wl
from close_kmers_b200 import api, build, synth
from test_oracle_family_fq_matrix import DNA_EDGE, assert_lookup_listing_equal, family_rows, parse_lookup_blocks

pytestmark = pytest.mark.gpu

OK_HEADER = "HTTP/1.1 200 OK\nContent-type: text/plain\n\n"


@pytest.fixture(scope="module")
def world(checkers, tmp_path_factory):
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so absent")
    protos, sig, img = wl.small_world(otu_mode="mixed")
    fam = synth.make_families(7, sig)
    d = str(tmp_path_factory.mktemp("kmerdir"))
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    yield protos, sig, img, fam, d


def fasta(ids, batch, width=60, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i, sid in enumerate(ids):
        s = batch.seq(i)
        out.append(b">" + sid.encode() + b" some definition" + nl)
        out += [s[k:k + width] + nl for k in range(0, len(s), width)]
    return b"".join(out)


def clean_batch(protos, seed, n):
    """Proteins a FASTA body can carry unchanged: letters only (the parser drops anything else), no empty records."""
    b = synth.make_proteins(seed, protos, n)
    seqs = [bytes(c for c in b.seq(i) if chr(c).isalpha()) or b"M" for i in range(b.n)]
    return synth.batch_from_strings(seqs)


# ---- /lookup through the C ABI ----------------------------------------------------------------------------------------

def test_family_scores_and_lookup_text(checkers, world):
    protos, sig, img, fam, d = world
    rows = family_rows(fam)
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    ref = checkers.Ref().open(d)
    ref.set_params()
    ref.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    ref.family_set_extra([r[3] for r in rows], [r[4] for r in rows], [r[5] for r in rows])
    guts = api.KmerGuts(kmer_dir=d)
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    try:
        batch = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(41, protos, 1500))
        ids = [f"s{i}" for i in range(batch.n)]
        want_sc, want_off = orc.family_scores(batch)
        got = guts.family_scores(batch.residues, batch.offsets)
        np.testing.assert_array_equal(got["score_offsets"], want_off)
        for f in ("id", "hit_count", "weighted_total"):  # weighted_total bit-exact: f32 sums in hit order
            np.testing.assert_array_equal(got["scores"][f], want_sc[f], err_msg=f)
        wl.assert_family_records_equal(got["matches"], orc.family_batch(batch), "matches from ckm_family_scores")
        assert len(want_sc) > 10_000
        # listings
        for thr in (0, 3, 7):
            mine = guts.lookup_text(ids, batch.residues, batch.offsets, rows, kmer_hit_threshold=thr, find_reps=thr == 7)
            assert_lookup_listing_equal(mine, ref.lookup_text(ids, batch, kmer_hit_threshold=thr, find_reps=thr == 7))
        # FamilyMapper::find_all_matches: the reference's own function (its threshold is fixed at three hits)
        assert_lookup_listing_equal(guts.find_all_matches_text(ids, batch.residues, batch.offsets, rows), ref.find_all_matches_text(ids, batch))
        # best match per sequence: strings exact, the rolled-up PGF score within 1e-6 (sum order), rest exact
        for ambig, genus in ((0, 0), (1, 1), (0, 2)):
            mine = guts.lookup_text(ids, batch.residues, batch.offsets, rows, find_best_match=True, allow_ambiguous_functions=bool(ambig),
                                    target_genus_id=genus).splitlines()
            want = ref.lookup_text(ids, batch, find_best_match=True, allow_ambiguous_functions=bool(ambig), target_genus_id=genus).splitlines()
            assert len(mine) == len(want) == batch.n
            n_fam = 0
            for a, b in zip(mine, want):
                fa, fb = a.split("\t"), b.split("\t")
                assert len(fa) == len(fb) == 8
                assert fa[0] == fb[0] and fa[4:] == fb[4:] and (fa[3] == "") == (fb[3] == ""), (a, b)
                assert abs(float(fa[2]) - float(fb[2])) <= 2e-6 * max(1.0, abs(float(fb[2]))), (a, b)
                if fa[1] != fb[1]:  # two PGFs with the same rolled-up score: the reference takes unordered_map order
                    assert float(fa[2]) == float(fb[2])
                n_fam += fa[3] != ""
            assert n_fam > 100
    finally:
        guts.close()
        ref.close()
        orc.close()


def test_peg_mode_lookup(checkers, world):
    protos, sig, img, fam, d = world
    ref = checkers.Ref().open(d)
    ref.set_params()
    ref.mapping_new()
    guts = api.KmerGuts(kmer_dir=d)
    pegs = api.KmerPegMapping()
    try:
        sub = synth.Prototypes(protos.codes[: int(protos.offsets[60])], protos.offsets[:61])
        added = synth.make_proteins(12, sub, 300, mix=(0.9, 0.1, 0.0, 0.0))
        add_ids = [f"fig|{i}.peg.1" for i in range(added.n)]
        ref.add_text(add_ids, added, silent=1)
        guts.add_text(pegs, add_ids, added.residues, added.offsets, silent=1)
        q = wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(13, sub, 200))
        qids = [f"q{i}" for i in range(q.n)]
        # the default threshold lists nothing in this mode (hit_total is never incremented)
        assert guts.lookup_text(qids, q.residues, q.offsets, mapping=pegs, family_mode=False) == \
            ref.lookup_text(qids, q, family_mode=False) == "".join(f"{i}\n//\n" for i in qids)
        mine = parse_lookup_blocks(guts.lookup_text(qids, q.residues, q.offsets, mapping=pegs, family_mode=False, kmer_hit_threshold=0))
        want = parse_lookup_blocks(ref.lookup_text(qids, q, family_mode=False, kmer_hit_threshold=0))
        assert [(i, sorted(r)) for i, r in mine] == [(i, sorted(r)) for i, r in want]
        assert sum(len(r) for _, r in want) > 1000
        # a second mapping is independent
        guts.postings_select(1)
        assert all(len(r) == 0 for _, r in parse_lookup_blocks(
            guts.lookup_text(qids, q.residues, q.offsets, mapping=api.KmerPegMapping(), family_mode=False, kmer_hit_threshold=0)))
        guts.postings_select(0)
        pairs, off = guts.postings_scores(q.residues, q.offsets)
        assert int(off[-1]) == sum(len(r) for _, r in want)
    finally:
        guts.close()
        ref.close()


def test_clones_share_tables_and_run_concurrently(checkers, world):
    protos, sig, img, fam, d = world
    orc = checkers.Oracle().open_image(img)
    orc.family_load(fam)
    guts = api.KmerGuts(kmer_dir=d)
    guts.family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    clones = [guts.clone() for _ in range(3)]
    try:
        batches = [wl.concat_batches(wl.edge_batch(protos), synth.make_proteins(90 + k, protos, 4000)) for k in range(4)]
        params = [dict(), dict(min_hits=3), dict(max_gap=50), dict(order_constraint=1)]
        out = [None] * 4

        def work(k):
            g = ([guts] + clones)[k]
            g.set_parameters(params[k])  # parameters are per engine
            for _ in range(5):
                calls = g.process_aa_seq_batch(batches[k].residues, batches[k].offsets, api.WANT_CALLS | api.WANT_BEST)
                fams = g.find_best_family_match_batch(batches[k].residues, batches[k].offsets)
            out[k] = (calls, fams)

        ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        for k in range(4):
            orc.set_params(**params[k])
            wl.assert_results_equal(out[k][0], orc.call_batch(batches[k], checkers.WANT_CALLS | checkers.WANT_BEST), f"clone {k}")
            wl.assert_family_records_equal(out[k][1], orc.family_batch(batches[k]), f"clone {k} families")
        with pytest.raises(api.CkmError):
            clones[0].family_load(fam.kmers, fam.fam_off, fam.fam_ids, fam.pgf, fam.plf, fam.function)
    finally:
        for c in clones:
            c.close()
        guts.close()
        orc.close()


# ---- the server --------------------------------------------------------------------------------------------------------

def http(port, head: bytes, body: bytes = b"", piecewise=0):
    s = socket.create_connection(("127.0.0.1", port), timeout=120)

    def send():  # the server answers while the body is still arriving: write and read concurrently, like a real client
        try:
            s.sendall(head)
            if piecewise:
                for k in range(0, len(body), piecewise):
                    s.sendall(body[k:k + piecewise])
            else:
                s.sendall(body)
        except OSError:
            pass  # the server may answer and close without reading a body (errors)

    try:
        t = threading.Thread(target=send)
        t.start()
        out = []
        while True:
            b = s.recv(1 << 20)
            if not b:
                break
            out.append(b)
        t.join()
        return b"".join(out).decode()
    finally:
        s.close()


def post(port, path, body, extra=b"", **kw):
    return http(port, b"POST %s HTTP/1.1\r\nHost: x\r\n%sContent-Length: %d\r\n\r\n" % (path.encode(), extra, len(body)), body, **kw)


@pytest.fixture(scope="module")
def server(checkers, world, tmp_path_factory):
    """A data directory as kser expects it (VERSION, families.dat, families.genus_map, families.nr/) and a running server."""
    protos, sig, img, fam, d = world
    build.build()
    rng = np.random.default_rng(99)
    # families.nr: ~1.6 M residues -> two load chunks (max_size_ = 1,000,000); one protein of the first chunk has no family
    nr = clean_batch(protos, 50, 5200)
    # most proteins sit in a family annotated with the function they are called with (4 families per function), the rest anywhere
    orc = checkers.Oracle().open_image(img)
    called = orc.call_batch(nr, checkers.WANT_CALLS | checkers.WANT_BEST)["best"]["function_index"].astype(np.int64)
    orc.close()
    nr_fam = rng.integers(0, fam.n_fams, nr.n).astype(np.uint32)
    consistent = (called >= 0) & (rng.random(nr.n) < 0.8)
    nr_fam[consistent] = (called[consistent] * 4 + rng.integers(0, 4, int(consistent.sum()))).astype(np.uint32)
    nr_ids = [f"fig|{1000 + i}.peg.{i % 7}" for i in range(nr.n)]
    orphan = 1500
    nr_fam[orphan] = 0xFFFFFFFF
    sizes = np.diff(nr.offsets.astype(np.int64))
    chunk_end = int(np.searchsorted(np.cumsum(sizes), 1_000_000)) + 1  # the protein that reaches max_size_ closes chunk 1
    assert orphan < chunk_end < nr.n
    os.makedirs(f"{d}/families.nr", exist_ok=True)
    with open(f"{d}/families.nr/nr.0", "wb") as f:
        f.write(fasta(nr_ids, nr))
    with open(f"{d}/VERSION", "w") as f:
        f.write("kmers-r1\nsecond line ignored\n")
    with open(f"{d}/families.version", "w") as f:
        f.write("fams-r7\n")
    with open(f"{d}/families.genus_map", "w") as f:
        for g in range(7):
            f.write(f"Genus{g}\t{1000 + g}\n")
    members = [[] for _ in range(fam.n_fams)]
    for i, f_ in enumerate(nr_fam):
        if f_ != 0xFFFFFFFF:
            members[int(f_)].append(i)
    total_size, count = np.zeros(fam.n_fams, np.uint64), np.zeros(fam.n_fams, np.uint64)
    with open(f"{d}/families.dat", "w") as f:
        for fid in range(fam.n_fams):  # ids are assigned in order of first appearance: keep file order = id order
            rows = [(nr_ids[i], int(sizes[i])) for i in members[fid]] or [(f"fig|unused.{fid}", 100 + fid)]
            for peg, ln in rows:
                f.write("\t".join([f"GF{fam.pgf[fid][4:]}", "x", "y", peg, str(ln), fam.function[fid], "z", f"Genus{fid % 7}", str(fid)]) + "\n")
                total_size[fid] += ln
                count[fid] += 1
    # what the reference builds from the same files
    ref = checkers.Ref().open(d)
    ref.set_params()
    first = synth.Batch(nr.residues[: int(nr.offsets[chunk_end])], nr.offsets[: chunk_end + 1])
    second = synth.Batch(nr.residues[int(nr.offsets[chunk_end]):], nr.offsets[chunk_end:] - nr.offsets[chunk_end])
    ref.family_nr_add(nr_fam[:chunk_end], first)
    ref.family_nr_add(nr_fam[chunk_end:], second)
    ref.family_set_data(fam.pgf, fam.plf, fam.function)
    genus = np.array([1000 + f % 7 for f in range(fam.n_fams)], np.uint64)
    ref.family_set_extra(genus, total_size, count.astype(np.uint16))
    pf = str(tmp_path_factory.mktemp("run") / "port")
    log = open(pf + ".log", "w")
    proc = subprocess.Popen([os.path.join(os.path.dirname(build.LIB), "kser_b200"), "--listen-port-file", pf, "--batch-mb", "1", "0", d],
                            stdout=log, stderr=subprocess.STDOUT)
    port = None
    for _ in range(1200):
        if proc.poll() is not None:
            break
        if os.path.exists(pf) and open(pf).read().strip():
            port = int(open(pf).read())
            break
        time.sleep(0.1)
    if port is None:
        proc.kill()
        pytest.fail("kser_b200 did not start:\n" + open(pf + ".log").read()[-4000:])
    yield port, ref, proc, pf + ".log"
    if proc.poll() is None:
        proc.send_signal(signal.SIGTERM)
        proc.wait(timeout=60)
    ref.close()


def test_server_get_routes_and_errors(server):
    port = server[0]
    body = "kmer\tkmers-r1\nfamilies\tfams-r7\nfamily-mode\t1\n"
    assert http(port, b"GET /version HTTP/1.1\r\n\r\n") == f"HTTP/1.1 200 OK\nContent-type: text/plain\nContent-length: {len(body)}\n\n{body}"
    assert http(port, b"GET /version HTTP/1.0\n\n").startswith("HTTP/1.0 200 OK\n")
    assert http(port, b"GET /nothing HTTP/1.1\r\n\r\n") == "HTTP/1.1 404 Not found\nContent-type: text/plain\nContent-length: 15\n\npath not found\n"
    assert http(port, b"GET /genus_lookup/Genus3 HTTP/1.1\r\n\r\n").endswith("\n\n1003\n")
    assert http(port, b"GET /genus_lookup/Nope HTTP/1.1\r\n\r\n") == \
        "HTTP/1.1 404 Not Found\nContent-type: text/plain\nContent-length: 16\n\ngenus not found\n"
    assert http(port, b"POST /query HTTP/1.1\r\n\r\n") == \
        "HTTP/1.1 500 Missing content length\nContent-type: text/plain\nContent-length: 30\n\nMissing content length header\n"
    assert http(port, b"POST /query HTTP/1.1\r\nTransfer-Encoding: chunked\r\n\r\n").startswith("HTTP/1.1 501 Chunked encoding not implemented\n")
    assert http(port, b"POST /query HTTP/1.1\r\nContent-Length: zz\r\n\r\n") == \
        "HTTP/1.1 500 Failed\nContent-type: text/plain\nContent-length: 23\n\nCaught exception stoul\n"
    assert http(port, b"POST /nothing HTTP/1.1\r\nContent-Length: 0\r\n\r\n").startswith("HTTP/1.1 404 Not found\n")
    assert http(port, b"nonsense\r\n\r\n") == ""
    assert http(port, b"POST /fq_lookup HTTP/1.1\r\nContent-Length: 0\r\n\r\n").endswith("\n\ndata done\n")


def test_server_query(world, server):
    protos = world[0]
    port, ref = server[0], server[1]
    batch = clean_batch(protos, 61, 3000)  # ~0.9 M residues: several 1 MB socket reads, one or two GPU batches
    ids = [f"fig|83333.1.peg.{i}" for i in range(batch.n)]
    body = fasta(ids, batch)
    assert post(port, "/query", body) == OK_HEADER + ref.query_text(ids, batch, 0, 0)
    assert post(port, "/query?details=1", body, piecewise=70_001) == OK_HEADER + ref.query_text(ids, batch, 1, 0)
    assert post(port, "/query?find_best_call=1", fasta(ids, batch, crlf=True)) == OK_HEADER + ref.query_text(ids, batch, 0, 1)
    # request parameters reach the engine: set_parameters per chunk
    ref.set_params(min_hits=3, max_gap=100)
    assert post(port, "/query?min_hits=3&max_gap=100", body) == OK_HEADER + ref.query_text(ids, batch, 0, 0)
    ref.set_params()
    # Expect: 100-continue, and an empty body (parse_complete emits one empty record)
    r = post(port, "/query", body[:5000] + b"\n", extra=b"Expect: 100-continue\r\n")
    assert r.startswith("HTTP/1.1 100 Continue\n\n" + OK_HEADER)
    assert post(port, "/query", b"") == OK_HEADER + ref.query_text([""], synth.batch_from_strings([b""]), 0, 0)


def test_server_survives_bad_clients(world, server):
    """Truncated bodies, oversized request lines, clients that hang up: the server answers what it can and keeps serving."""
    protos = world[0]
    port, ref = server[0], server[1]
    batch = clean_batch(protos, 63, 50)
    ids = [f"t{i}" for i in range(batch.n)]
    body = fasta(ids, batch)
    # body shorter than Content-Length, then the client stops sending: handled like eof (parse_complete on what arrived)
    s = socket.create_connection(("127.0.0.1", port), timeout=60)
    s.sendall(b"POST /query HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % (len(body) + 1000) + body)
    s.shutdown(socket.SHUT_WR)
    got = b""
    while True:
        b = s.recv(1 << 20)
        if not b:
            break
        got += b
    s.close()
    assert got.decode() == OK_HEADER + ref.query_text(ids, batch, 0, 0)
    # a request line that never ends, a client that connects and leaves, binary junk
    for junk in (b"GET /" + b"a" * (3 << 20), b"", b"\x00\xff\x16\x03\x01" * 100 + b"\n\n"):
        s = socket.create_connection(("127.0.0.1", port), timeout=60)
        try:
            s.sendall(junk)
        except OSError:
            pass
        s.close()
    # a client that posts and hangs up without reading the answer
    big = clean_batch(protos, 64, 3000)
    s = socket.create_connection(("127.0.0.1", port), timeout=60)
    bb = fasta([f"h{i}" for i in range(big.n)], big)
    s.sendall(b"POST /query?details=1 HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % len(bb) + bb[: len(bb) // 2])
    s.close()
    # concurrent requests of different kinds
    out = [None] * 6
    def one(k):
        if k % 3 == 0:
            out[k] = post(port, "/query", body)
        elif k % 3 == 1:
            out[k] = post(port, f"/mapping/stress{k}/add?silent=1", body)
        else:
            out[k] = http(port, b"GET /version HTTP/1.1\r\n\r\n")
    ts = [threading.Thread(target=one, args=(k,)) for k in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    want = OK_HEADER + ref.query_text(ids, batch, 0, 0)
    assert out[0] == want and out[3] == want
    assert out[1] == OK_HEADER and out[4] == OK_HEADER  # silent /add: header only
    assert out[2].startswith("HTTP/1.1 200 OK\n") and "family-mode\t1" in out[5]


def fastq(ids, batch, crlf=False):
    out = []
    for i, sid in enumerate(ids):
        s = batch.seq(i)
        out.append(b"@" + sid.encode() + b" 1:N:0\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    return b"".join(out)


def test_server_family_paths(checkers, world, server):
    """fq_lookup and lookup run on the family table the server built from families.dat + families.nr at start-up."""
    protos, sig, img, fam, d = world
    port, ref = server[0], server[1]
    reads = synth.make_reads(3, protos, 3000)
    keep = [i for i in range(reads.n) if reads.seq(i).isalpha() and len(reads.seq(i)) > 0]
    reads = synth.batch_from_strings([reads.seq(i) for i in keep])
    rids = [f"read{i}" for i in range(reads.n)]
    want = ref.fq_text(rids, reads)
    if want.count("\n") < 500:
        pytest.fail("reference produced too little fq output; the start-up family load is not in place")
    body = fastq(rids, reads)
    got = post(port, "/fq_lookup", body)
    assert got.startswith(OK_HEADER)
    assert wl.assert_fq_text_equal(got[len(OK_HEADER):], want) < want.count("\n") // 2
    gz = post(port, "/fq_lookup", gzip.compress(body), piecewise=50_000)
    assert gz == got
    # /lookup, family mode
    prot = clean_batch(protos, 71, 600)
    pids = [f"p{i}" for i in range(prot.n)]
    pbody = fasta(pids, prot)
    got = post(port, "/lookup", pbody)
    assert got.startswith(OK_HEADER)
    assert_lookup_listing_equal(got[len(OK_HEADER):], ref.lookup_text(pids, prot))
    got = post(port, "/lookup?find_best_match=1&target_genus=Genus2", pbody)
    want_lines = ref.lookup_text(pids, prot, find_best_match=True, target_genus_id=1002).splitlines()
    got_lines = got[len(OK_HEADER):].splitlines()
    assert len(got_lines) == len(want_lines) == prot.n
    hits = 0
    for a, b in zip(got_lines, want_lines):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[0] == fb[0] and fa[3:] == fb[3:], (a, b)
        assert abs(float(fa[2]) - float(fb[2])) <= 2e-6 * max(1.0, abs(float(fb[2])))
        hits += fa[3] != ""
    assert hits > 20, hits


def test_server_two_devices_concurrent_connections(checkers, world, server):
    """--device 0,1: the family table built on device 0 is replicated to device 1 (export + load), engines on both devices
    serve concurrent connections; every response still equals the reference's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    protos, sig, img, fam, d = world
    ref = server[1]
    pf = os.path.join(d, "port2")
    log = open(pf + ".log", "w")
    proc = subprocess.Popen([os.path.join(os.path.dirname(build.LIB), "kser_b200"), "--listen-port-file", pf, "--device", "0,1",
                             "--n-kmer-threads", "2", "--batch-mb", "1", "0", d], stdout=log, stderr=subprocess.STDOUT)
    try:
        port = None
        for _ in range(1200):
            if proc.poll() is not None:
                break
            if os.path.exists(pf) and open(pf).read().strip():
                port = int(open(pf).read())
                break
            time.sleep(0.1)
        assert port is not None, open(pf + ".log").read()[-3000:]
        batch = clean_batch(protos, 62, 2000)
        ids = [f"fig|9.9.peg.{i}" for i in range(batch.n)]
        body = fasta(ids, batch)
        want = OK_HEADER + ref.query_text(ids, batch, 0, 0)
        prot = clean_batch(protos, 72, 400)
        pids = [f"p{i}" for i in range(prot.n)]
        want_lookup = ref.lookup_text(pids, prot)
        out = [None] * 8

        def one(k):
            out[k] = post(port, "/query", body) if k % 2 == 0 else post(port, "/lookup", fasta(pids, prot))

        ts = [threading.Thread(target=one, args=(k,)) for k in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        for k in range(8):
            if k % 2 == 0:
                assert out[k] == want
            else:
                assert out[k].startswith(OK_HEADER)
                assert_lookup_listing_equal(out[k][len(OK_HEADER):], want_lookup)
        assert http(port, b"GET /quit HTTP/1.1\r\n\r\n").endswith("OK, quitting\n")
        assert proc.wait(timeout=60) == 0
    finally:
        if proc.poll() is None:
            proc.kill()


def test_server_add_and_matrix(world, server):
    protos = world[0]
    port, ref = server[0], server[1]
    sub = synth.Prototypes(protos.codes[: int(protos.offsets[40])], protos.offsets[:41])
    batch = synth.batch_from_strings([bytes(c for c in s if chr(c).isalpha()) or b"M" for s in
                                      (synth.make_proteins(11, sub, 240, mix=(0.9, 0.1, 0.0, 0.0)).seq(i) for i in range(240))])
    ids = [f"fig|{i % 230}.peg.{i % 230}" for i in range(batch.n)]
    ref.mapping_new()  # /mapping/m1 starts empty on both sides (this drops the reference's family tables: keep this test last)
    half = batch.n // 2
    for lo, hi, silent in ((0, half, 0), (half, batch.n, 1)):
        part = synth.Batch(batch.residues[int(batch.offsets[lo]):int(batch.offsets[hi])], batch.offsets[lo:hi + 1] - batch.offsets[lo])
        want = ref.add_text(ids[lo:hi], part, silent=silent)
        assert post(port, f"/mapping/m1/add?silent={silent}", fasta(ids[lo:hi], part)) == OK_HEADER + want
    order = np.random.default_rng(1).permutation(batch.n)[:200]
    req = synth.batch_from_strings([batch.seq(i) for i in order])
    req_ids = [ids[i] for i in order]
    got = post(port, "/mapping/m1/matrix", fasta(req_ids, req))
    assert got == OK_HEADER + ref.matrix_text(req_ids, req) and got.count("\n") > 500
    # a different key is a different KmerPegMapping: nothing was added there
    assert post(port, "/mapping/m2/matrix", fasta(req_ids, req)) == OK_HEADER


def test_server_quit(server):
    port, _, proc, log = server
    assert http(port, b"GET /quit HTTP/1.1\r\n\r\n") == "HTTP/1.1 200 OK\nContent-type: text/plain\nContent-length: 13\n\nOK, quitting\n"
    assert proc.wait(timeout=60) == 0
    text = open(log, errors="replace").read()
    assert "Listening on 0.0.0.0:" in text and "NO FAM FOR id='fig|2500.peg.2'" in text
