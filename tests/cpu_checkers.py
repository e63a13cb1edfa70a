"""ctypes bindings of the two CPU checkers (TEST INFRASTRUCTURE):

* ``Ref``    -- oracle/_ref/libckm_ref.so: the reference's own object code behind ref_driver.cc.
* ``Oracle`` -- oracle/libckm_oracle.so: the plain-C restatement (ckm_oracle.c).

Both return results in the ckm.h record layout as numpy structured arrays, so the parity tests compare
them (and the CUDA path) field by field.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libckm_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libckm_oracle.so")

WANT_CALLS, WANT_HITS, WANT_OTU, WANT_BEST = 1, 2, 4, 8
BEST_HAS_CALLS, BEST_AMBIG = 1, 2

CALL_DT = np.dtype([("start", "<u4"), ("end", "<u4"), ("count", "<i4"), ("function_index", "<u4"),
                    ("weighted_hits", "<f4")])
HIT_DT = np.dtype([("which_kmer", "<u8"), ("offset", "<u4"), ("otu_index", "<i4"), ("function_index", "<i4"),
                   ("function_wt", "<f4"), ("avg_from_end", "<u2"), ("pad_", "<u2"), ("pad2_", "<u4")])
OTU_DT = np.dtype([("otu_index", "<i4"), ("count", "<i4")])
BEST_DT = np.dtype([("function_index", "<i4"), ("ambig_a", "<i4"), ("ambig_b", "<i4"), ("flags", "<u4"),
                    ("score", "<f4"), ("weighted_score", "<f4"), ("score_offset", "<f4")])
SLOT_DT = np.dtype([("which_kmer", "<u8"), ("otu_index", "<i4"), ("avg_from_end", "<u2"), ("pad_", "<u2"),
                    ("function_index", "<i4"), ("function_wt", "<f4")])
FAMILY_DT = np.dtype([("gfam", "<i4"), ("lfam", "<i4"), ("gfam_score", "<f4"), ("lfam_score", "<f4"), ("score", "<f4"),
                      ("function_index", "<i4")])
assert CALL_DT.itemsize == 20 and HIT_DT.itemsize == 32 and BEST_DT.itemsize == 28 and SLOT_DT.itemsize == 24


class BatchOutC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("call_offsets", C.c_void_p), ("calls", C.c_void_p), ("hit_offsets", C.c_void_p),
                ("hits", C.c_void_p), ("otu_offsets", C.c_void_p), ("otus", C.c_void_p), ("best", C.c_void_p),
                ("n_probes", C.c_uint64), ("n_hits", C.c_uint64)]


class FqOutC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("best_frame", C.c_void_p), ("best_score", C.c_void_p), ("match_offsets", C.c_void_p),
                ("matches", C.c_void_p), ("n_fragments", C.c_uint64), ("n_probes", C.c_uint64)]


PAIR_DT = np.dtype([("eid_i", "<u4"), ("eid_j", "<u4"), ("count", "<u8")])
FQ_MATCH_DT = np.dtype([("length", "<u4"), ("gfam", "<i4"), ("lfam", "<i4"), ("gfam_score", "<f4"), ("lfam_score", "<f4"),
                        ("score", "<f4"), ("function_index", "<i4")])
assert FQ_MATCH_DT.itemsize == 28


def unpack_fq_out(o) -> dict:
    n = o.n
    off = _arr(o.match_offsets, n + 1, np.uint64)
    return {"n": n, "best_frame": _arr(o.best_frame, n, np.int32), "best_score": _arr(o.best_score, n, np.float64),
            "match_offsets": off, "matches": _arr(o.matches, int(off[-1]) if n + 1 else 0, FQ_MATCH_DT),
            "n_fragments": o.n_fragments, "n_probes": o.n_probes}


def _arr(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def unpack_batch_out(o: BatchOutC) -> dict:
    """Copy a ckm_batch_out_t into numpy arrays (the C side may free / reuse its buffers afterwards)."""
    n = o.n
    r = {"n": n, "n_probes": o.n_probes, "n_hits": o.n_hits}
    for name, dt in (("call", CALL_DT), ("hit", HIT_DT), ("otu", OTU_DT)):
        offp = getattr(o, f"{name}_offsets")
        if offp:
            off = _arr(offp, n + 1, np.uint64)
            r[f"{name}_offsets"] = off
            r[f"{name}s"] = _arr(getattr(o, f"{name}s"), int(off[-1]) if n + 1 else 0, dt)
    if o.best:
        r["best"] = _arr(o.best, n, BEST_DT)
    return r


def ensure_built() -> None:
    """(Re)build the checkers when sources are newer; the reference .so only when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, stdout=subprocess.DEVNULL)


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _cstr_array(strs):
    bs = [s.encode() if isinstance(s, str) else s for s in strs]
    arr = (C.c_char_p * len(bs))(*bs)
    return arr


def parse_dump(raw: bytes, n: int):
    out, pos = [], 0
    for _ in range(n):
        nl = raw.index(b"\n", pos)
        a, b = (int(x) for x in raw[pos:nl].split())
        pos = nl + 1
        out.append((raw[pos:pos + a], raw[pos + a:pos + a + b]))
        pos += a + b + 1
    assert pos == len(raw)
    return out


class Ref:
    """The reference's own code.  Chatty on stdout/stderr (its own prints)."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.L = C.CDLL(REF_SO)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p, C.c_int]
        L.ref_close.argtypes = [C.c_void_p]
        L.ref_call_batch.restype = C.c_void_p
        L.ref_call_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_out_view.restype = C.POINTER(BatchOutC)
        L.ref_out_view.argtypes = [C.c_void_p]
        L.ref_out_best_function.restype = C.c_char_p
        L.ref_out_best_function.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_out_otus_sorted.restype = C.c_void_p
        L.ref_out_otus_sorted.argtypes = [C.c_void_p]
        L.ref_out_free.argtypes = [C.c_void_p]
        L.ref_encoded_aa_kmer.restype = C.c_uint64
        L.ref_encoded_aa_kmer.argtypes = [C.c_char_p]
        L.ref_encoder_encoded_aa_kmer.restype = C.c_uint64
        L.ref_encoder_encoded_aa_kmer.argtypes = [C.c_char_p]
        L.ref_decoded_kmer.argtypes = [C.c_uint64, C.c_char_p]
        L.ref_build_image.argtypes = [C.c_char_p, C.c_longlong, C.c_uint64] + [C.c_void_p] * 5
        L.ref_build_image_str.argtypes = [C.c_char_p, C.c_longlong, C.c_uint64] + [C.c_void_p] * 5
        L.ref_set_params_kv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_get_params.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.ref_function_count.argtypes = [C.c_void_p]
        L.ref_function_at_index.restype = C.c_char_p
        L.ref_function_at_index.argtypes = [C.c_void_p, C.c_int]
        L.ref_bench_calls.restype = C.c_double
        L.ref_bench_calls.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        for f in ("ref_query_text", "ref_add_text"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int] + (
                [C.c_int] if f == "ref_query_text" else [])
        L.ref_matrix_text.restype = C.c_void_p
        L.ref_matrix_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.ref_mapping_new.argtypes = [C.c_void_p]
        L.ref_family_load.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.ref_family_nr_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_family_set_data.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_family_table_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.ref_family_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_family_clear.argtypes = [C.c_void_p]
        L.ref_family_set_extra.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_lookup_text.restype = C.c_void_p
        L.ref_lookup_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint, C.c_int, C.c_int,
                                      C.c_int, C.c_uint64]
        L.ref_parse_text.restype = C.c_void_p
        L.ref_parse_text.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.ref_family_text.restype = C.c_void_p
        L.ref_family_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_family_batch.restype = C.c_void_p
        L.ref_family_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_fq_text.restype = C.c_void_p
        L.ref_fq_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_fq_batch.restype = C.c_void_p
        L.ref_fq_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.ref_six_frames.restype = C.c_void_p
        L.ref_six_frames.argtypes = [C.c_char_p]
        L.ref_free_text.argtypes = [C.c_void_p]
        self.h = None

    # -- builder / statics
    def build_image(self, kmer_dir, nbuckets, sig):
        keys = np.ascontiguousarray(sig.keys, np.uint64)
        self.L.ref_build_image(kmer_dir.encode(), nbuckets, len(keys), keys.ctypes.data,
                               np.ascontiguousarray(sig.fI, np.int32).ctypes.data,
                               np.ascontiguousarray(sig.oI, np.int32).ctypes.data,
                               np.ascontiguousarray(sig.avg, np.uint16).ctypes.data,
                               np.ascontiguousarray(sig.wt, np.float32).ctypes.data)

    def build_image_str(self, kmer_dir, nbuckets, kmers: bytes, fI, oI, avg, wt):
        n = len(kmers) // 8
        self.L.ref_build_image_str(kmer_dir.encode(), nbuckets, n, kmers,
                                   np.ascontiguousarray(fI, np.int32).ctypes.data,
                                   np.ascontiguousarray(oI, np.int32).ctypes.data,
                                   np.ascontiguousarray(avg, np.uint16).ctypes.data,
                                   np.ascontiguousarray(wt, np.float32).ctypes.data)

    def encoded_aa_kmer(self, s: bytes) -> int:
        return self.L.ref_encoded_aa_kmer(s)

    def encoder_encoded_aa_kmer(self, s: bytes) -> int:
        return self.L.ref_encoder_encoded_aa_kmer(s)

    def decoded_kmer(self, k: int) -> bytes:
        b = C.create_string_buffer(9)
        self.L.ref_decoded_kmer(k, b)
        return b.value

    # -- engine
    def open(self, kmer_dir, threads=1):
        self.h = self.L.ref_open(kmer_dir.encode(), threads)
        return self

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None

    def set_params(self, **kv):
        keys = _cstr_array(list(kv.keys()))
        vals = _cstr_array([str(v) for v in kv.values()])
        self.L.ref_set_params_kv(self.h, len(kv), keys, vals)

    def get_params(self):
        v = [C.c_int() for _ in range(4)]
        self.L.ref_get_params(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def function_at_index(self, i):
        return self.L.ref_function_at_index(self.h, i).decode()

    def call_batch(self, batch, flags):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        o = self.L.ref_call_batch(self.h, res.ctypes.data, off.ctypes.data, batch.n, flags)
        r = unpack_batch_out(self.L.ref_out_view(o).contents)
        if flags & WANT_BEST:
            r["best_function"] = [self.L.ref_out_best_function(o, i).decode() for i in range(batch.n)]
        if flags & WANT_OTU:
            r["otus_sorted"] = _arr(self.L.ref_out_otus_sorted(o), len(r["otus"]), OTU_DT)
        self.L.ref_out_free(o)
        return r

    def bench_calls(self, batch, want_best=True):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        tot = C.c_uint64()
        return self.L.ref_bench_calls(self.h, res.ctypes.data, off.ctypes.data, batch.n, int(want_best), C.byref(tot))

    def _text(self, p):
        s = C.string_at(p).decode()
        self.L.ref_free_text(p)
        return s

    def query_text(self, ids, batch, details=0, find_best_call=0):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        return self._text(self.L.ref_query_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n,
                                                details, find_best_call))

    def mapping_new(self):
        self.L.ref_mapping_new(self.h)

    def add_text(self, ids, batch, silent=0):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        return self._text(self.L.ref_add_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n, silent))

    def matrix_text(self, ids, batch):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        npairs = C.c_uint64()
        return self._text(self.L.ref_matrix_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n,
                                                 C.byref(npairs)))

    def family_load(self, kmers, fam_off, fam_ids, pgf, plf, function):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        fam_off = np.ascontiguousarray(fam_off, np.uint64)
        fam_ids = np.ascontiguousarray(fam_ids, np.uint32)
        self.L.ref_family_load(self.h, len(kmers), kmers.ctypes.data, fam_off.ctypes.data, fam_ids.ctypes.data, len(pgf),
                               _cstr_array(pgf), _cstr_array(plf), _cstr_array(function))

    def family_set_extra(self, genus_id, total_size, count):
        g = np.ascontiguousarray(genus_id, np.uint64)
        t = np.ascontiguousarray(total_size, np.uint64)
        c = np.ascontiguousarray(count, np.uint16)
        self.L.ref_family_set_extra(self.h, len(g), g.ctypes.data, t.ctypes.data, c.ctypes.data)

    def lookup_text(self, ids, batch, family_mode=True, kmer_hit_threshold=3, find_best_match=False, find_reps=False,
                    allow_ambiguous_functions=False, target_genus_id=0):
        """LookupRequest's worker loop (lookup_request.cc:138-400) over the reference engine."""
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        return self._text(self.L.ref_lookup_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n, int(family_mode),
                                                 kmer_hit_threshold, int(find_best_match), int(find_reps),
                                                 int(allow_ambiguous_functions), target_genus_id))

    def find_all_matches_text(self, ids, batch):
        """FamilyMapper::find_all_matches (family_mapper.cc:207-285), the reference's own code."""
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        self.L.ref_find_all_matches_text.restype = C.c_void_p
        self.L.ref_find_all_matches_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        return self._text(self.L.ref_find_all_matches_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n))

    def parse_text(self, fastq: bool, text: bytes, cuts=()):
        """Reference FastaParser / FastqParser over `text` fed packet by packet; returns [(id, seq)] as bytes."""
        cuts = np.ascontiguousarray(sorted(cuts), np.uint64)
        n = C.c_uint64()
        p = self.L.ref_parse_text(int(fastq), text, len(text), cuts.ctypes.data, len(cuts), C.byref(n))
        raw = C.string_at(p)  # ids / sequences never contain NUL here
        self.L.ref_free_text(C.c_void_p(p))
        return parse_dump(raw, n.value)

    # NRLoader::thread_load in family mode (+ KmerInserter + add_fam_mapping) for one chunk
    def family_nr_add(self, fam_ids, batch):
        fam_ids = np.ascontiguousarray(fam_ids, np.uint32)
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        self.L.ref_family_nr_add(self.h, fam_ids.ctypes.data, res.ctypes.data, off.ctypes.data, batch.n)

    def family_set_data(self, pgf, plf, function):
        self.L.ref_family_set_data(self.h, len(pgf), _cstr_array(pgf), _cstr_array(plf), _cstr_array(function))

    def family_table(self):
        """kmer_to_family_id_ as (kmers, fam_off, fam_ids), canonical order."""
        from close_kmers_b200.api import canonical_family_csr
        nk, ne = C.c_uint64(), C.c_uint64()
        self.L.ref_family_table_size(self.h, C.byref(nk), C.byref(ne))
        k = np.zeros(nk.value, np.uint64)
        o = np.zeros(nk.value + 1, np.uint64)
        ids = np.zeros(max(ne.value, 1), np.uint32)
        self.L.ref_family_table(self.h, k.ctypes.data, o.ctypes.data, ids.ctypes.data)
        return canonical_family_csr(k, o, ids[:ne.value])

    def family_clear(self):
        self.L.ref_family_clear(self.h)

    def family_text(self, batch):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        return self._text(self.L.ref_family_text(self.h, res.ctypes.data, off.ctypes.data, batch.n))

    def family_batch(self, batch):
        """find_best_family_match per sequence: dict(gfam, lfam, function: lists of str; gscore, lscore, score: f32)."""
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        g, l, s = (np.zeros(batch.n, np.float32) for _ in range(3))
        txt = self._text(self.L.ref_family_batch(self.h, res.ctypes.data, off.ctypes.data, batch.n, g.ctypes.data,
                                                 l.ctypes.data, s.ctypes.data))
        rows = [ln.split("\t") for ln in txt.split("\n")[:batch.n]]
        return dict(gfam=[r[0] for r in rows], lfam=[r[1] for r in rows], function=[r[2] for r in rows], gscore=g, lscore=l,
                    score=s)

    def fq_text(self, ids, batch):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        return self._text(self.L.ref_fq_text(self.h, _cstr_array(ids), res.ctypes.data, off.ctypes.data, batch.n))

    def fq_batch(self, batch):
        """Per read: (best_frame, best_score, [(len, gfam, gscore, lfam, lscore, function, score), ...]), exact floats."""
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        txt = self._text(self.L.ref_fq_batch(self.h, res.ctypes.data, off.ctypes.data, batch.n))
        out = []
        for ln in txt.split("\n")[:batch.n]:
            c = ln.split("\t")
            ms = []
            for k in range(int(c[2])):
                b = 3 + 7 * k
                ms.append((int(c[b]), c[b + 1], float.fromhex(c[b + 2]), c[b + 3], float.fromhex(c[b + 4]), c[b + 5],
                           float.fromhex(c[b + 6])))
            out.append((int(c[0]), float.fromhex(c[1]), ms))
        return out

    def six_frames(self, bases: bytes) -> str:
        return self._text(self.L.ref_six_frames(bases))


class ParamsC(C.Structure):
    _fields_ = [("order_constraint", C.c_int), ("min_hits", C.c_int), ("min_weighted_hits", C.c_int), ("max_gap", C.c_int)]


class Oracle:
    def __init__(self):
        L = self.L = C.CDLL(ORACLE_SO)
        L.orc_open_image.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.orc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_num_sigs.restype = C.c_uint64
        L.orc_num_sigs.argtypes = [C.c_void_p]
        L.orc_to_amino_acid_off.restype = C.c_uint8
        L.orc_to_amino_acid_off.argtypes = [C.c_char]
        L.orc_encoded_aa_kmer.restype = C.c_uint64
        L.orc_encoded_aa_kmer.argtypes = [C.c_char_p]
        L.orc_decoded_kmer.argtypes = [C.c_uint64, C.c_char_p]
        L.orc_lookup_hash_entry.restype = C.c_int64
        L.orc_lookup_hash_entry.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_call_batch.restype = C.POINTER(BatchOutC)
        L.orc_call_batch.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_out_free.argtypes = [C.c_void_p]
        L.orc_find_best_call.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_family_new.restype = C.c_void_p
        L.orc_family_new.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
        L.orc_family_free.argtypes = [C.c_void_p]
        L.orc_family_batch.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_void_p]
        L.orc_translate_frame.restype = C.c_size_t
        L.orc_translate_frame.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p]
        L.orc_fq_batch.restype = C.POINTER(FqOutC)
        L.orc_fq_batch.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_fq_out_free.argtypes = [C.c_void_p]
        L.orc_postings_new.restype = C.c_void_p
        L.orc_postings_free.argtypes = [C.c_void_p]
        L.orc_postings_add.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_postings_count.restype = C.c_uint64
        L.orc_postings_count.argtypes = [C.c_void_p]
        L.orc_matrix_rows.restype = C.c_void_p
        L.orc_matrix_rows.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                      C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_bench_calls.restype = C.c_double
        L.orc_bench_calls.argtypes = [C.c_void_p, C.POINTER(ParamsC), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                      C.POINTER(C.c_uint64)]
        self.t = None
        self.params = ParamsC(0, 5, 0, 200)
        self._keep = None

    def open(self, kmer_dir):
        t = C.c_void_p()
        rc = self.L.orc_open(kmer_dir.encode(), C.byref(t))
        if rc:
            raise RuntimeError(f"orc_open rc={rc}")
        self.t = t
        return self

    def open_image(self, image: np.ndarray):
        t = C.c_void_p()
        rc = self.L.orc_open_image(image.ctypes.data, image.nbytes, C.byref(t))
        if rc:
            raise RuntimeError(f"orc_open_image rc={rc}")
        self._keep = image
        self.t = t
        return self

    def try_open_image(self, image: np.ndarray) -> int:
        t = C.c_void_p()
        rc = self.L.orc_open_image(image.ctypes.data, image.nbytes, C.byref(t))
        if rc == 0:
            self.L.orc_close(t)
        return rc

    def close(self):
        if self.t:
            self.L.orc_close(self.t)
            self.t = None

    def set_params(self, order_constraint=0, min_hits=5, min_weighted_hits=0, max_gap=200):
        self.params = ParamsC(order_constraint, min_hits, min_weighted_hits, max_gap)

    def encoded_aa_kmer(self, s: bytes) -> int:
        return self.L.orc_encoded_aa_kmer(s)

    def decoded_kmer(self, k: int) -> bytes:
        b = C.create_string_buffer(9)
        self.L.orc_decoded_kmer(k, b)
        return b.value

    def lookup(self, key: int) -> int:
        return self.L.orc_lookup_hash_entry(self.t, key)

    def call_batch(self, batch, flags):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        o = self.L.orc_call_batch(self.t, C.byref(self.params), res.ctypes.data, off.ctypes.data, batch.n, flags)
        r = unpack_batch_out(o.contents)
        self.L.orc_out_free(o)
        return r

    def find_best_call(self, calls: np.ndarray):
        calls = np.ascontiguousarray(calls, CALL_DT)
        out = np.zeros(1, BEST_DT)
        self.L.orc_find_best_call(calls.ctypes.data, len(calls), out.ctypes.data)
        return out[0]

    def family_load(self, fam):
        """fam: close_kmers_b200.synth.FamilyTables (strings interned there)."""
        self.fam = self.L.orc_family_new(len(fam.kmers), fam.kmers.ctypes.data, fam.fam_off.ctypes.data,
                                         fam.fam_ids.ctypes.data, fam.n_fams, fam.fam_func_sid.ctypes.data,
                                         fam.fam_pgf.ctypes.data, fam.n_pgf, len(fam.func_sid), fam.func_sid.ctypes.data,
                                         fam.hypo_sid)
        self._fam_keep = fam

    def family_batch(self, batch):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        out = np.zeros(batch.n, FAMILY_DT)
        self.L.orc_family_batch(self.t, C.byref(self.params), self.fam, res.ctypes.data, off.ctypes.data, batch.n,
                                out.ctypes.data)
        return out

    def family_scores(self, batch):
        """(scores[SCORE_DT], score_offsets): LookupRequest::on_hit in family mode, ascending family id."""
        from close_kmers_b200.api import SCORE_DT
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        L = self.L
        L.orc_family_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        ps, po = C.c_void_p(), C.c_void_p()
        L.orc_family_scores(self.t, C.byref(self.params), self.fam, res.ctypes.data, off.ctypes.data, batch.n, C.byref(ps), C.byref(po))
        so = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), (batch.n + 1,)).copy()
        ns = int(so[-1])
        sc = np.frombuffer(C.string_at(ps, ns * SCORE_DT.itemsize), SCORE_DT).copy() if ns else np.zeros(0, SCORE_DT)
        L.orc_free(ps)
        L.orc_free(po)
        return sc, so

    def family_nr_build(self, chunks):
        """chunks: iterable of (fam_ids, batch).  Returns (kmers, fam_off, fam_ids) sorted by (k-mer, family)."""
        L = self.L
        L.orc_postings_new.restype = C.c_void_p
        L.orc_family_nr_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_family_nr_table.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_postings_free.argtypes = [C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        p = C.c_void_p(L.orc_postings_new())
        for fam_ids, batch in chunks:
            fam_ids = np.ascontiguousarray(fam_ids, np.uint32)
            res = np.ascontiguousarray(batch.residues, np.uint8)
            off = np.ascontiguousarray(batch.offsets, np.uint64)
            L.orc_family_nr_add(self.t, C.byref(self.params), p, fam_ids.ctypes.data, res.ctypes.data, off.ctypes.data, batch.n)
        nk, ne = C.c_uint64(), C.c_uint64()
        pk, po, pi = C.c_void_p(), C.c_void_p(), C.c_void_p()
        L.orc_family_nr_table(p, C.byref(nk), C.byref(ne), C.byref(pk), C.byref(po), C.byref(pi))
        k = np.ctypeslib.as_array(C.cast(pk, C.POINTER(C.c_uint64)), (max(nk.value, 1),))[:nk.value].copy()
        o = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), (nk.value + 1,)).copy()
        ids = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint32)), (max(ne.value, 1),))[:ne.value].copy()
        for q in (pk, po, pi):
            L.orc_free(q)
        L.orc_postings_free(p)
        return k, o, ids

    def postings_new(self):
        self.post = C.c_void_p(self.L.orc_postings_new())

    def postings_add(self, eids, batch):
        eids = np.ascontiguousarray(eids, np.uint32)
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        self.L.orc_postings_add(self.t, C.byref(self.params), self.post, eids.ctypes.data, res.ctypes.data, off.ctypes.data,
                                batch.n)

    def postings_export(self):
        self.L.orc_postings_count.restype = C.c_uint64
        n = self.L.orc_postings_count(self.post)
        keys, eids = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        self.L.orc_postings_export(self.post, keys.ctypes.data_as(C.c_void_p), eids.ctypes.data_as(C.c_void_p))
        return keys, eids

    def postings_import(self, keys, eids):
        keys, eids = np.ascontiguousarray(keys, np.uint64), np.ascontiguousarray(eids, np.uint32)
        self.L.orc_postings_import(self.post, keys.ctypes.data_as(C.c_void_p), eids.ctypes.data_as(C.c_void_p), C.c_uint64(len(keys)))

    def matrix_rows(self, eids, batch, row_begin=0, row_end=None):
        eids = np.ascontiguousarray(eids, np.uint32)
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        npairs = C.c_uint64()
        p = self.L.orc_matrix_rows(self.t, C.byref(self.params), self.post, eids.ctypes.data, res.ctypes.data, off.ctypes.data,
                                   batch.n, row_begin, batch.n if row_end is None else row_end, C.byref(npairs))
        out = _arr(p, npairs.value, PAIR_DT)
        self.L.orc_free(p)
        return out

    def translate_frame(self, dna: bytes, frame: int) -> bytes:
        out = C.create_string_buffer(len(dna) // 3 + 2)
        n = self.L.orc_translate_frame(dna, len(dna), frame, out)
        return out.raw[:n]

    def six_frames(self, dna: bytes):
        """[(frame, [fragments...])] with boost::split(token_compress_on) semantics (dna_seq.cc:9-23)."""
        out = []
        for f in (1, 2, 3, -1, -2, -3):
            p = self.translate_frame(dna, f)
            toks, i, n = [], 0, len(p)
            while True:  # maximal stop-free runs; a leading / trailing stop run still yields one empty token
                j = i
                while j < n and p[j:j + 1] != b"*":
                    j += 1
                toks.append(p[i:j])
                if j >= n:
                    break
                while j < n and p[j:j + 1] == b"*":
                    j += 1
                i = j
                if i >= n:
                    toks.append(b"")
                    break
            out.append((f, toks))
        return out

    def fq_batch(self, batch):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        o = self.L.orc_fq_batch(self.t, C.byref(self.params), self.fam, res.ctypes.data, off.ctypes.data, batch.n)
        r = unpack_fq_out(o.contents)
        self.L.orc_fq_out_free(o)
        return r

    def bench_calls(self, batch, threads, want_best=True):
        res = np.ascontiguousarray(batch.residues, np.uint8)
        off = np.ascontiguousarray(batch.offsets, np.uint64)
        tot = C.c_uint64()
        return self.L.orc_bench_calls(self.t, C.byref(self.params), res.ctypes.data, off.ctypes.data, batch.n,
                                      int(want_best), threads, C.byref(tot))


def best_function_string(best_rec, function_at_index) -> str:
    """Host-side naming of a ckm_best_t (kguts.cc:1160, 1176-1196): 'F1 ?? F2' with the greater string first."""
    if best_rec["flags"] & BEST_AMBIG:
        f1 = function_at_index(int(best_rec["ambig_a"]))
        f2 = function_at_index(int(best_rec["ambig_b"]))
        if f2 > f1:
            f1, f2 = f2, f1
        return f1 + " ?? " + f2
    if best_rec["function_index"] >= 0:
        return function_at_index(int(best_rec["function_index"]))
    return ""
