"""Seeded synthetic workloads for the signature-k-mer calling path.

Nothing here is on the product path: it only manufactures inputs (signature images in the
reference's file format, proteins, reads, family tables) for tests/ and bench.py.  The image bytes are
produced by ``ckm_image_build`` (the library's mirror of the reference builder, kguts.cc:77-115,
188-234), and tests/test_image_build.py checks they are byte-identical to what the reference's own
``insert_kmer`` / ``save_kmer_hash_table`` write.
"""
from __future__ import annotations

import dataclasses

import numpy as np

AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)  # kguts.cc:30-32
MAX_ENCODED = 20**8
# build_signature_kmers.cc:862-865: bucket count = first entry > 3 * n_kmers
PRIMES = [3769, 6337, 12791, 24571, 51043, 101533, 206933, 400187, 821999, 2000003, 4000037, 8000009, 16000057,
          32000011, 64000031, 128000003, 248000009, 508000037, 1073741824, 1400303159, 2147483648, 1190492993,
          3559786523, 6461346257]
POW20 = (20 ** np.arange(7, -1, -1)).astype(np.uint64)


def bucket_count(n_kmers: int) -> int:
    for p in PRIMES:
        if p > 3 * n_kmers:
            return p
    raise ValueError(f"no prime for {n_kmers} k-mers")


@dataclasses.dataclass
class Prototypes:
    codes: np.ndarray    # uint8 0..19, concatenated
    offsets: np.ndarray  # int64, P+1

    @property
    def n(self) -> int:
        return len(self.offsets) - 1


@dataclasses.dataclass
class Signatures:
    keys: np.ndarray  # uint64, distinct, insertion order
    fI: np.ndarray    # int32
    oI: np.ndarray    # int32
    avg: np.ndarray   # uint16
    wt: np.ndarray    # float32
    n_functions: int


@dataclasses.dataclass
class Batch:
    residues: np.ndarray  # uint8 ASCII, concatenated (no separators)
    offsets: np.ndarray   # uint64, n+1

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def seq(self, i: int) -> bytes:
        return self.residues[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()


def make_prototypes(seed: int, n_protos: int, mean_len: int = 300, sd: float = 0.0) -> Prototypes:
    rng = np.random.default_rng(seed)
    if sd > 0:
        lens = np.clip(np.rint(rng.normal(mean_len, sd, n_protos)), 50, 1200).astype(np.int64)
    else:
        lens = np.full(n_protos, mean_len, dtype=np.int64)
    offsets = np.zeros(n_protos + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    codes = rng.integers(0, 20, int(offsets[-1]), dtype=np.uint8)
    return Prototypes(codes, offsets)


def window_keys(codes: np.ndarray) -> np.ndarray:
    """key[p] = sum_i codes[p+i] * 20^(7-i) for every p in [0, len-8] (kguts.cc:438-455)."""
    n = len(codes) - 7
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    k = np.zeros(n, dtype=np.uint64)
    for i in range(8):
        k = k * np.uint64(20) + codes[i:i + n].astype(np.uint64)
    return k


def make_signatures(protos: Prototypes, n_sigs: int, n_functions: int | None = None, otu_mode: str = "minus1",
                    seed: int = 0, dedupe: bool = True) -> Signatures:
    """First ``n_sigs`` distinct 8-mers prototype by prototype, position by position.

    ``dedupe=False`` (used for the 1e8-k-mer bench image, where a global sort would dominate set-up) keeps
    the ~0.2% repeated 8-mers: the reference builder never checks for duplicates either (kguts.cc:202-222)
    and the first one in probe order wins on lookup, for the oracle and the CUDA path alike."""
    P = protos.n
    F = n_functions if n_functions is not None else min(50_000, P)
    keys = window_keys(protos.codes)
    lens = np.diff(protos.offsets)
    proto = np.repeat(np.arange(P, dtype=np.int32), lens)[:len(keys)]
    pos = (np.arange(len(keys), dtype=np.int64) - protos.offsets[:-1].repeat(lens)[:len(keys)]).astype(np.int32)
    plen = lens.astype(np.int32).repeat(lens)[:len(keys)]
    valid = pos <= plen - 8  # windows that do not straddle two prototypes
    keys, proto, pos, plen = keys[valid], proto[valid], pos[valid], plen[valid]
    if dedupe:
        _, first = np.unique(keys, return_index=True)
        first.sort()
        first = first[:n_sigs]
        keys, proto, pos, plen = keys[first], proto[first], pos[first], plen[first]
    else:
        keys, proto, pos, plen = keys[:n_sigs], proto[:n_sigs], pos[:n_sigs], plen[:n_sigs]
    proto = proto.astype(np.int64)
    pos = pos.astype(np.int64)
    fI = (proto % F).astype(np.int32)
    if otu_mode == "minus1":  # what build_signature_kmers.cc:708 writes
        oI = np.full(len(keys), -1, dtype=np.int32)
    else:  # exercise the OTU map with a handful of ids
        oI = ((proto * 7 + pos // 64) % 11).astype(np.int32) - 1
    avg = (plen - pos).astype(np.uint16)
    wt = (1.0 + ((31 * proto + pos) % 500) / 100.0).astype(np.float32)
    return Signatures(keys, fI, oI, avg, wt, F)


def make_signatures_sparse(protos: Prototypes, n_sigs: int, keep: float = 0.4, jitter: int = 12, foreign: float = 0.1,
                           n_functions: int | None = None, seed: int = 1, dedupe: bool = True) -> Signatures:
    """Signature sets as build_signature_kmers leaves them, not as make_signatures idealises them: only a random ``keep``
    fraction of every prototype's windows are signatures (the discriminating ones), ``avg_from_end`` is an average over the
    proteins a k-mer was seen in (here: the true distance jittered by up to ``jitter``), and a ``foreign`` fraction of a
    prototype's signatures belongs to some other function.  Consecutive signatures of a prototype therefore overlap by
    seven residues only with probability ``keep``: the worst case for the neighbour-ordered copy of the table."""
    rng = np.random.default_rng(seed)
    dense = make_signatures(protos, 1 << 62, n_functions=n_functions, dedupe=dedupe)
    take = np.flatnonzero(rng.random(len(dense.keys)) < keep)[:n_sigs]
    keys, fI, avg, wt = dense.keys[take], dense.fI[take].copy(), dense.avg[take].astype(np.int64), dense.wt[take]
    avg = np.clip(avg + rng.integers(-jitter, jitter + 1, len(avg)), 0, 65535).astype(np.uint16)
    other = rng.random(len(fI)) < foreign
    fI[other] = rng.integers(0, dense.n_functions, int(other.sum()), dtype=np.int32)
    return Signatures(keys, fI, dense.oI[take], avg, wt, dense.n_functions)


def function_names(n: int) -> list[str]:
    return [f"function {i}" for i in range(n)]


def write_index_files(kmer_dir: str, n_functions: int, n_otus: int = 0) -> None:
    """function.index / otu.index: '<idx>\\t<text>\\n', dense, in order (kguts.cc:544-575)."""
    with open(f"{kmer_dir}/function.index", "w") as f:
        for i, name in enumerate(function_names(n_functions)):
            f.write(f"{i}\t{name}\n")
    with open(f"{kmer_dir}/otu.index", "w") as f:
        for i in range(n_otus):
            f.write(f"{i}\totu {i}\n")


def _to_ascii(codes: np.ndarray) -> np.ndarray:
    return AA[codes]


def make_proteins(seed: int, protos: Prototypes, n: int, mix=(0.80, 0.10, 0.05, 0.05), sub_rate: float = 0.05) -> Batch:
    """80% mutated prototype / 10% chimera of two prototypes / 5% random / 5% prototype with X's and a
    lowercase stretch.  Vectorised: every protein is built from at most two prototype
    slices, then noise is applied on the concatenated array."""
    rng = np.random.default_rng(seed)
    P = protos.n
    kind = rng.choice(4, size=n, p=np.asarray(mix) / np.sum(mix))
    a = rng.integers(0, P, n)
    b = rng.integers(0, P, n)
    la = protos.offsets[a + 1] - protos.offsets[a]
    lb = protos.offsets[b + 1] - protos.offsets[b]
    cut = np.where(kind == 1, rng.integers(20, np.maximum(la - 20, 21)), la)
    # second part: for chimeras take prototype b from `cut` (clipped to its length) to its end
    start_b = np.minimum(cut, lb)
    len_b = np.where(kind == 1, lb - start_b, 0)
    lens = cut + len_b
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    # source index for every output residue
    idx = np.arange(total, dtype=np.int64)
    owner = np.repeat(np.arange(n), lens)
    rel = idx - offsets[owner]
    in_a = rel < cut[owner]
    src = np.where(in_a, protos.offsets[a[owner]] + rel, protos.offsets[b[owner]] + start_b[owner] + (rel - cut[owner]))
    codes = protos.codes[src].copy()
    # substitutions on every non-random protein
    mut = rng.random(total) < sub_rate
    codes[mut] = rng.integers(0, 20, int(mut.sum()), dtype=np.uint8)
    rnd = kind[owner] == 2
    codes[rnd] = rng.integers(0, 20, int(rnd.sum()), dtype=np.uint8)
    res = _to_ascii(codes)
    amb = kind[owner] == 3
    xs = amb & (rng.random(total) < 0.01)
    res[xs] = ord("X")
    # a lowercase stretch of 12 residues somewhere in each "ambiguous" protein
    amb_ids = np.nonzero(kind == 3)[0]
    if len(amb_ids):
        st = offsets[amb_ids] + rng.integers(0, np.maximum(lens[amb_ids] - 12, 1))
        for k in range(12):
            p = np.minimum(st + k, offsets[amb_ids + 1] - 1)
            res[p] = res[p] | 0x20
    return Batch(res, offsets.astype(np.uint64))


def make_proteins_parallel(seed: int, protos: Prototypes, n: int, chunk: int = 32768, workers: int | None = None,
                           **kw) -> Batch:
    """make_proteins in independent seeded chunks on a thread pool (numpy releases the GIL in the heavy ops).
    Chunk c uses seed ``seed * 1_000_003 + c``, so the result does not depend on the worker count."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(0, n, chunk))
    workers = workers or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(lambda c: make_proteins(seed * 1_000_003 + c, protos, min(chunk, n - starts[c]), **kw),
                            range(len(starts))))
    sizes = np.array([int(p.offsets[-1]) for p in parts], dtype=np.uint64)
    base = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    offsets = np.concatenate([parts[0].offsets[:1]] + [p.offsets[1:] + base[i] for i, p in enumerate(parts)]).astype(np.uint64)
    return Batch(np.concatenate([p.residues for p in parts]), offsets)


@dataclasses.dataclass
class FamilyTables:
    """What NRLoader / load_families leave in KmerPegMapping (kmer.h:118-127), plus interned string ids."""
    kmers: np.ndarray      # uint64, distinct
    fam_off: np.ndarray    # uint64, len(kmers)+1
    fam_ids: np.ndarray    # uint32, lists deduped
    pgf: list
    plf: list
    function: list
    # interning used by the oracle (the product interns on its own in ckm_family_load)
    fam_func_sid: np.ndarray
    fam_pgf: np.ndarray
    pgf_names: list
    func_sid: np.ndarray
    hypo_sid: int

    @property
    def n_fams(self) -> int:
        return len(self.pgf)

    @property
    def n_pgf(self) -> int:
        return len(self.pgf_names)


def make_families(seed: int, sig: Signatures, fams_per_function: int = 4, max_list: int = 8, coverage: float = 0.9,
                  hypothetical_every: int = 17) -> FamilyTables:
    """each signature k-mer maps to 1+Geom(0.5) families (cap 8) drawn from the families of its own
    function (4 per function, two PGFs per function); ``coverage`` of the k-mers have a list at all.  Every
    ``hypothetical_every``-th function's families are annotated "hypothetical protein"."""
    rng = np.random.default_rng(seed)
    F = sig.n_functions
    n_fams = F * fams_per_function
    fam_function_idx = np.arange(n_fams) // fams_per_function
    names = function_names(F)
    function = [("hypothetical protein" if (f % hypothetical_every == 0) else names[f]) for f in fam_function_idx]
    pgf = [f"PGF_{(f // 2):08d}" for f in range(n_fams)]
    plf = [f"PLF_{1000 + f % 7}_{f:08d}" for f in range(n_fams)]
    pick = rng.random(len(sig.keys)) < coverage
    kmers = sig.keys[pick].astype(np.uint64)
    kf = sig.fI[pick].astype(np.int64)
    cnt = np.minimum(rng.geometric(0.5, len(kmers)), min(max_list, 2 * fams_per_function)).astype(np.int64)  # ids in a list are distinct
    fam_off = np.zeros(len(kmers) + 1, dtype=np.uint64)
    fam_off[1:] = np.cumsum(cnt)
    owner = np.repeat(np.arange(len(kmers)), cnt)
    rank = np.arange(int(fam_off[-1])) - fam_off[:-1].astype(np.int64)[owner]
    # distinct ids within a list: own function's families first, then the neighbouring function's
    start = rng.integers(0, fams_per_function, len(kmers))[owner]
    slot = (start + rank) % (2 * fams_per_function)
    fam_ids = ((kf[owner] * fams_per_function + slot) % n_fams).astype(np.uint32)
    # interning
    sid = {"hypothetical protein": 0}
    func_sid = np.array([sid.setdefault(nm, len(sid)) for nm in names], dtype=np.uint32)
    fam_func_sid = np.array([sid.setdefault(nm, len(sid)) for nm in function], dtype=np.uint32)
    pg = {}
    fam_pgf = np.array([pg.setdefault(nm, len(pg)) for nm in pgf], dtype=np.uint32)
    return FamilyTables(kmers, fam_off, fam_ids, pgf, plf, function, fam_func_sid, fam_pgf, list(pg.keys()), func_sid, 0)


NCBI11_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"  # NCBI genetic code 11, TCAG order


def _codon_tables():
    """aa code (0..19) -> (codons[6,3] uint8 ASCII, count)."""
    order = "TCAG"
    by_aa = {}
    for pos, aa in enumerate(NCBI11_AAS):
        codon = order[pos // 16] + order[(pos // 4) % 4] + order[pos % 4]
        by_aa.setdefault(aa, []).append(codon)
    cod = np.zeros((20, 6, 3), dtype=np.uint8)
    cnt = np.zeros(20, dtype=np.int64)
    for i, a in enumerate(AA.tobytes().decode()):
        cs = by_aa[a]
        cnt[i] = len(cs)
        for k in range(6):
            cod[i, k] = np.frombuffer(cs[k % len(cs)].encode(), np.uint8)
    return cod, cnt


_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtN", b"TGCAtgcaN"):
    _COMP[_a] = _b


def make_reads(seed: int, protos: Prototypes, n: int, read_len: int = 150, n_rate: float = 0.005) -> Batch:
    """Reads cut from back-translated prototypes (synonymous codons uniform), random strand and phase,
    0.5% N.  All reads have ``read_len`` bases."""
    rng = np.random.default_rng(seed)
    cod, cnt = _codon_tables()
    naa = read_len // 3 + 2
    P = protos.n
    p = rng.integers(0, P, n)
    plen = protos.offsets[p + 1] - protos.offsets[p]
    start = (rng.random(n) * np.maximum(plen - naa, 1)).astype(np.int64)
    idx = protos.offsets[p][:, None] + np.minimum(start[:, None] + np.arange(naa)[None, :], (plen - 1)[:, None])
    aa = protos.codes[idx]                                  # n x naa
    pick = (rng.random((n, naa)) * cnt[aa]).astype(np.int64)
    dna = cod[aa, pick].reshape(n, naa * 3)                 # n x 3*naa
    phase = rng.integers(0, 3, n)
    cols = phase[:, None] + np.arange(read_len)[None, :]
    reads = np.take_along_axis(dna, cols, axis=1)
    rev = rng.random(n) < 0.5
    reads[rev] = _COMP[reads[rev][:, ::-1]]
    reads[rng.random(reads.shape) < n_rate] = ord("N")
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len))
    return Batch(reads.reshape(-1).copy(), offsets)


def batch_from_strings(seqs) -> Batch:
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    residues = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return Batch(residues, offsets)


def n_probes_expected(batch: Batch) -> int:
    """Number of probed windows: starts p < len-8 whose 8 residues are all valid."""
    valid = np.isin(batch.residues, AA)
    total = 0
    # windowed all-valid via cumulative sum of invalid flags
    bad = np.concatenate([[0], np.cumsum(~valid)])
    for i in range(batch.n):
        lo, hi = int(batch.offsets[i]), int(batch.offsets[i + 1])
        L = hi - lo
        if L <= 8:
            continue
        p = np.arange(lo, hi - 8)
        total += int(np.count_nonzero(bad[p + 8] - bad[p] == 0))
    return total
