"""Front-end data formats, CPU only: the streaming FASTA / FASTQ parsers against the reference's own
FastaParser / FastqParser object code fed packet by packet, and the request-head parsing / dispatch of KmerRequest2
against known answers derived from krequest2.cc."""
import os

import numpy as np
import pytest

from close_kmers_b200 import api

FASTA_CASES = [
    b"",
    b">",
    b">a",
    b">a\n",
    b">a\nACDE",
    b">a\nACDE\n",
    b">a\nACDE\nFGHI\n>b desc here\nKLMN\n",
    b">a\tdef with tab\nACDE\n\n\nFGH\n>b\n\n>c\nAAA",
    b">a\r\nACDE\r\nFG\r\n>b\r\nHI\r\n",
    b"garbage before\n>a\nACD\n",
    b">a\nAC*DE\n*FG\n>b\nacde\nXYZ123\n",
    b">a\nAC DE\n>b\nF-G.H\n",
    b">a\n>b\n>c\nA\n",
    b">id|with|pipes.peg.1 some function [organism]\nMKV\nLLA\n>second\nMMM\n",
    b">a\nACD\n>",
    b"\n\n>a\nACD\n",
    b">a\x80\xff\nAC\xc3\xa9DE\n",
]

FASTQ_CASES = [
    b"",
    b"@r1\nACGT\n+\nIIII\n",
    b"@r1 desc\nACGT\n+r1\nIIII\n@r2\nGGCC\n+\nJJJJ\n",
    b"@r1\nACGT\n+\nIIII",
    b"@r1\nACGT\n+\n@III\n@r2\nTTTT\n+\n++++\n",
    b">r1\nACGT\n+\nIIII\n",
    b"@r1\nAC GT\n+\nIIII\n\n@r2\nAAAA\n+\nIIII\n",
    b"@r1\r\nACGT\r\n+\r\nIIII\r\n",
    b"@r1\nACGT\nIIII\n+\nJJJJ\n@r2\nCC\n+\nII\n",
    b"@r1\nacgtnN\n+\n!!!!!!\n",
    b"@\n\n+\n\n",
]


def mine(fastq, text, cuts=()):
    p = api.SeqParser(fastq)
    out, pos = [], 0
    for c in list(sorted(cuts)) + [len(text)]:
        p.feed(text[pos:c])
        pos = c
        out += p.take()  # taking between packets must not disturb the sequence in progress
    p.complete()
    return out + p.take()


@pytest.fixture(scope="module")
def ref(checkers):
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so not built (needs /root/reference)")
    return checkers.Ref()


@pytest.mark.parametrize("fastq,cases", [(False, FASTA_CASES), (True, FASTQ_CASES)])
def test_parser_cases_match_reference(ref, fastq, cases):
    for text in cases:
        want = ref.parse_text(fastq, text)
        assert mine(fastq, text) == want, text
        for cut in range(1, len(text)):  # every packet boundary
            assert mine(fastq, text, [cut]) == want, (text, cut)


@pytest.mark.parametrize("fastq", [False, True])
def test_parser_fuzz_matches_reference(ref, fastq):
    rng = np.random.default_rng(17 + fastq)
    alphabet = np.frombuffer(b">>\n\n\n\r \tACGTacdeKLMNWY**x19@@++-|.", np.uint8)
    for trial in range(300):
        n = int(rng.integers(0, 400))
        text = bytes(rng.choice(alphabet, n))
        cuts = sorted(set(int(c) for c in rng.integers(0, n + 1, int(rng.integers(0, 5))))) if n else []
        assert mine(fastq, text, cuts) == ref.parse_text(fastq, text, cuts), (text, cuts)


def test_parser_well_formed_bulk(ref):
    rng = np.random.default_rng(3)
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
    recs = []
    for i in range(2000):
        seq = bytes(rng.choice(aa, int(rng.integers(1, 700))))  # an empty record would swallow the next header (s_data sees ">")
        lines = [seq[k:k + 60] for k in range(0, len(seq), 60)]
        recs.append(b">fig|%d.peg.%d function %d\n" % (i, i, i) + b"".join(l + b"\n" for l in lines))
    text = b"".join(recs)
    cuts = list(range(65536, len(text), 65536))
    got = mine(False, text, cuts)
    assert got == ref.parse_text(False, text, cuts) and len(got) == 2000
    p = api.SeqParser(False)
    p.feed(text)
    assert p.pending() == sum(len(s) for _, s in got[:-1])  # the last record is still open
    p.complete()
    assert len(p.take()) == 2000 and p.n_errors == 0


def test_parser_reports_errors():
    p = api.SeqParser(False)
    p.feed(b"x>a\nAC1\n")
    p.complete()
    assert p.take() == [(b"a", b"AC")] and p.n_errors == 2
    assert p.last_error() == "Error found: Bad data character '1' at line 2 id='a'"
    q = api.SeqParser(True)
    q.feed(b">r\n")
    q.take()
    assert q.last_error() == "Error found: Missing @ at line 2 id=''" and q.n_errors == 3  # '>', 'r', and the newline


def test_request_line_regex():
    d = api.http_describe(b"POST /query?details=1&find_best_call=0;min_hits=3&junk&=x#frag HTTP/1.1\r\nContent-Length: 12\r\n\r\n")
    assert d["type"] == "POST" and d["path"] == "/query" and d["version"] == "1.1" and d["fragment"] == "frag"
    assert d["parameters"] == "details=1&find_best_call=0;min_hits=3&junk&=x"
    assert d["param.details"] == "1" and d["param.find_best_call"] == "0" and d["param.min_hits"] == "3" and d["param."] == "x"
    assert "param.junk" not in d
    assert d["header.content-length"] == "12" and d["decision"] == "post /query key= length=12"
    # the version is the LAST " HTTP/x.y"; a path may contain spaces; '#' before '?' makes everything a fragment
    d = api.http_describe(b"GET /a b HTTP/1.0 HTTP/1.1\n\n")
    assert d["path"] == "/a b HTTP/1.0" and d["version"] == "1.1"
    d = api.http_describe(b"GET /p#f?x=1 HTTP/1.0\n\n")
    assert d["path"] == "/p" and d["fragment"] == "f?x=1" and d["parameters"] == "" and d["version"] == "1.0"
    for bad in (b"get / HTTP/1.1\n\n", b"GET HTTP/1.1\n\n", b"GET / HTTP/1\n\n", b"GET / HTTP/1.1.1\n\n", b"GET /  http/1.1\n\n",
                b" GET / HTTP/1.1\n\n", b"GET / HTTP/1.1 \n\n", b"\n"):
        assert api.http_describe(bad)["decision"] == "invalid", bad
    assert api.http_describe(b"GET  HTTP/1.1\n\n")["path"] == ""  # an empty path is allowed by ([^?#]*)


def test_headers_and_dispatch():
    D = lambda head: api.http_describe(head)["decision"]
    assert D(b"GET /quit HTTP/1.1\n\n") == "get quit"
    assert D(b"GET /version HTTP/1.1\n\n") == "get version"
    assert D(b"GET /genus_lookup/Escherichia HTTP/1.1\n\n") == "get genus_lookup Escherichia"
    assert D(b"GET /genus_lookup/ HTTP/1.1\n\n") == "respond 404 Not found"
    assert D(b"GET /genus_lookup/a/b HTTP/1.1\n\n") == "respond 404 Not found"
    assert D(b"GET /nothing HTTP/1.1\n\n") == "respond 404 Not found"
    assert D(b"POST /query HTTP/1.1\n\n") == "respond 500 Missing content length"
    assert D(b"POST /query HTTP/1.1\nContent-Length: abc\n\n") == "respond 500 Failed"
    assert D(b"POST /query HTTP/1.1\nContent-Length:   77xyz\n\n") == "post /query key= length=77"
    assert D(b"POST /nothing HTTP/1.1\nContent-Length: 1\n\n") == "respond 404 Not found"
    assert D(b"POST /query HTTP/1.1\nTransfer-Encoding: chunked\n\n") == "respond 501 Chunked encoding not implemented"
    assert D(b"PUT /query HTTP/1.1\nContent-Length: 1\n\n") == "none"
    for action in ("add", "matrix", "lookup"):
        assert D(b"POST /mapping/k1/%s HTTP/1.1\ncontent-length: 5\n\n" % action.encode()) == f"post /{action} key=k1 length=5"
    assert D(b"POST /mapping/k1/query HTTP/1.1\ncontent-length: 5\n\n") == "respond 404 Not found"
    assert D(b"POST /mapping//add HTTP/1.1\ncontent-length: 5\n\n") == "respond 404 Not found"
    assert D(b"POST /mapping/a/b/add HTTP/1.1\ncontent-length: 5\n\n") == "respond 404 Not found"
    assert D(b"POST /fq_lookup HTTP/1.1\ncontent-length: 0\n\n") == "post /fq_lookup key= length=0"
    d = api.http_describe(b"POST /add HTTP/1.1\r\nExpect: 100-continue\r\nContent-Length: 3\r\nX-Odd\r\nA:  b : c\r\n\r\n")
    assert d["continue"] == "1" and d["decision"] == "post /add key= length=3"
    assert d["header.x-odd"] == "X-Odd" and d["header.a"] == "b : c" and d["header.expect"] == "100-continue"
