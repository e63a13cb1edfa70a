"""Python host-side mirror of the reference's KmerGuts interface, bound to libckm.so through ctypes.

The class and method names follow the reference (kguts.h:334-372) so that tests read like the
reference's call sites; the unit of work is a *batch* of sequences (one request body chunk) instead
of one sequence.  There is no CPU path here: if libckm.so cannot be loaded, or no sm_100 device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

WANT_CALLS, WANT_HITS, WANT_OTU, WANT_BEST = 1, 2, 4, 8
TUNE_PLAIN_PROBE, TUNE_UNFUSED, TUNE_NO_FALLBACK = 32, 0x100000, 0x200000  # ckm_set_tuning (include/ckm.h)
BEST_HAS_CALLS, BEST_AMBIG = 1, 2
MAX_ENCODED = 20**8

CALL_DT = np.dtype([("start", "<u4"), ("end", "<u4"), ("count", "<i4"), ("function_index", "<u4"),
                    ("weighted_hits", "<f4")])
HIT_DT = np.dtype([("which_kmer", "<u8"), ("offset", "<u4"), ("otu_index", "<i4"), ("function_index", "<i4"),
                   ("function_wt", "<f4"), ("avg_from_end", "<u2"), ("pad_", "<u2"), ("pad2_", "<u4")])
OTU_DT = np.dtype([("otu_index", "<i4"), ("count", "<i4")])
FAMILY_DT = np.dtype([("gfam", "<i4"), ("lfam", "<i4"), ("gfam_score", "<f4"), ("lfam_score", "<f4"), ("score", "<f4"),
                      ("function_index", "<i4")])
BEST_DT = np.dtype([("function_index", "<i4"), ("ambig_a", "<i4"), ("ambig_b", "<i4"), ("flags", "<u4"),
                    ("score", "<f4"), ("weighted_score", "<f4"), ("score_offset", "<f4")])
SLOT_DT = np.dtype([("which_kmer", "<u8"), ("otu_index", "<i4"), ("avg_from_end", "<u2"), ("pad_", "<u2"),
                    ("function_index", "<i4"), ("function_wt", "<f4")])


class CkmError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libckm error {code}: {text}")
        self.code = code


class BatchOutC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("call_offsets", C.c_void_p), ("calls", C.c_void_p), ("hit_offsets", C.c_void_p),
                ("hits", C.c_void_p), ("otu_offsets", C.c_void_p), ("otus", C.c_void_p), ("best", C.c_void_p),
                ("n_probes", C.c_uint64), ("n_hits", C.c_uint64)]


class DeviceOutC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("d_n_hits", C.c_void_p), ("d_n_calls", C.c_void_p), ("d_calls", C.c_void_p),
                ("min_hits_for_call_base", C.c_int32), ("d_best", C.c_void_p), ("d_totals", C.c_void_p)]


class FqOutC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("best_frame", C.c_void_p), ("best_score", C.c_void_p), ("match_offsets", C.c_void_p),
                ("matches", C.c_void_p), ("n_fragments", C.c_uint64), ("n_probes", C.c_uint64)]


class FqFragmentsC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_fragments", C.c_uint64), ("frag_frame_offsets", C.c_void_p),
                ("frag_offsets", C.c_void_p), ("residues", C.c_void_p)]


PAIR_DT = np.dtype([("eid_i", "<u4"), ("eid_j", "<u4"), ("count", "<u8")])
FQ_MATCH_DT = np.dtype([("length", "<u4"), ("gfam", "<i4"), ("lfam", "<i4"), ("gfam_score", "<f4"), ("lfam_score", "<f4"),
                        ("score", "<f4"), ("function_index", "<i4")])

_lib = None


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) libckm.so.  Fails loudly; never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CKM_LIB_PATH") or _build.build()  # CKM_LIB_PATH: a prebuilt variant of the library, for A/B runs
    L = C.CDLL(path)
    L.ckm_last_error.restype = C.c_char_p
    L.ckm_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.ckm_open_image.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_void_p)]
    L.ckm_close.argtypes = [C.c_void_p]
    L.ckm_function_at_index.restype = C.c_char_p
    L.ckm_function_at_index.argtypes = [C.c_void_p, C.c_int32]
    L.ckm_otu_at_index.restype = C.c_char_p
    L.ckm_otu_at_index.argtypes = [C.c_void_p, C.c_int32]
    L.ckm_function_count.argtypes = [C.c_void_p]
    L.ckm_otu_count.argtypes = [C.c_void_p]
    L.ckm_num_sigs.restype = C.c_uint64
    L.ckm_num_sigs.argtypes = [C.c_void_p]
    L.ckm_table_slot_bytes.argtypes = [C.c_void_p]
    L.ckm_l2_fetch_granularity.argtypes = [C.c_void_p]
    L.ckm_has_occupancy_bitmap.argtypes = [C.c_void_p]
    L.ckm_set_tuning.argtypes = [C.c_void_p, C.c_uint32]
    L.ckm_last_batch_was_fused.argtypes = [C.c_void_p]
    L.ckm_table_buckets.restype = C.c_uint64
    L.ckm_table_buckets.argtypes = [C.c_void_p]
    L.ckm_chain_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ckm_copy_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.ckm_set_default_params.argtypes = [C.c_void_p]
    L.ckm_set_params.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ckm_get_params.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    L.ckm_encoded_aa_kmer.restype = C.c_uint64
    L.ckm_encoded_aa_kmer.argtypes = [C.c_char_p]
    L.ckm_decoded_kmer.argtypes = [C.c_uint64, C.c_char_p]
    L.ckm_call_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(BatchOutC)]
    L.ckm_call_batch_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(BatchOutC)]
    L.ckm_packed_words.restype = C.c_uint64
    L.ckm_packed_words.argtypes = [C.c_uint64]
    L.ckm_pack_residues.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    L.ckm_call_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
    L.ckm_device_results.argtypes = [C.c_void_p, C.POINTER(DeviceOutC)]
    L.ckm_read_totals.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.ckm_stream.restype = C.c_void_p
    L.ckm_stream.argtypes = [C.c_void_p]
    L.ckm_launch_count.restype = C.c_uint64
    L.ckm_launch_count.argtypes = [C.c_void_p]
    L.ckm_synchronize.argtypes = [C.c_void_p]
    L.ckm_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.ckm_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.ckm_calibrate_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ckm_family_load.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    L.ckm_family_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.ckm_family_nr_begin.argtypes = [C.c_void_p]
    L.ckm_family_nr_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.ckm_family_nr_finish.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64)]
    L.ckm_family_export.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    for f in ("ckm_family_pgf_name", "ckm_family_plf_name"):
        getattr(L, f).restype = C.c_char_p
        getattr(L, f).argtypes = [C.c_void_p, C.c_int32]
    L.ckm_family_function_name.restype = C.c_char_p
    L.ckm_family_function_name.argtypes = [C.c_void_p, C.c_void_p]
    L.ckm_fq_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(FqOutC)]
    L.ckm_fq_translate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(FqFragmentsC)]
    L.ckm_fq_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.ckm_family_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.ckm_postings_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.ckm_postings_append_last.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.ckm_postings_clear.argtypes = [C.c_void_p]
    L.ckm_postings_count.restype = C.c_uint64
    L.ckm_postings_count.argtypes = [C.c_void_p]
    L.ckm_matrix_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.ckm_postings_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.ckm_postings_import_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.ckm_matrix_rows_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ckm_mapping_new.restype = C.c_void_p
    L.ckm_mapping_free.argtypes = [C.c_void_p]
    L.ckm_mapping_encode_id.restype = C.c_uint32
    L.ckm_mapping_encode_id.argtypes = [C.c_void_p, C.c_char_p]
    L.ckm_mapping_decode_id.restype = C.c_char_p
    L.ckm_mapping_decode_id.argtypes = [C.c_void_p, C.c_uint32]
    L.ckm_add_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                               C.POINTER(C.c_void_p)]
    L.ckm_matrix_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.ckm_matrix_merge_pairs.restype = C.c_uint64
    L.ckm_matrix_merge_pairs.argtypes = [C.c_void_p, C.c_uint64]
    L.ckm_query_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                 C.POINTER(C.c_void_p)]
    L.ckm_free_text.argtypes = [C.c_void_p]
    L.ckm_image_build.argtypes = [C.c_uint64, C.c_uint64] + [C.c_void_p] * 6 + [C.c_size_t]
    L.ckm_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.ckm_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def pack_residues(residues: np.ndarray, offsets: np.ndarray) -> tuple:
    """ASCII batch -> (packed uint32 words, word offsets): seven residues per word (base 22), every sequence from a word
    boundary (csrc/ckm_packed.cuh)."""
    residues = np.ascontiguousarray(residues, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint64)
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64))
    cap = int(((lens + 6) // 7).sum()) if n else 0
    packed = np.zeros(cap + 2, np.uint32)
    woff = np.zeros(n + 1, np.uint64)
    _check(lib().ckm_pack_residues(residues.ctypes.data, offsets.ctypes.data, n, packed.ctypes.data, cap, woff.ctypes.data))
    return packed, woff


def experiments_enabled() -> bool:
    """True when libckm.so was built with -DCKM_EXPERIMENTS (A/B kernels and cache-policy bits of ckm_set_tuning)."""
    return bool(lib().ckm_experiments_enabled())


def _check(rc: int) -> None:
    if rc != 0:
        raise CkmError(rc, lib().ckm_last_error().decode(errors="replace"))


def _arr(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def encoded_aa_kmer(kmer: bytes) -> int:
    """KmerGuts::encoded_aa_kmer (kguts.cc:457-471)."""
    return lib().ckm_encoded_aa_kmer(kmer)


def decoded_kmer(key: int) -> bytes:
    """KmerGuts::decoded_kmer (kguts.cc:473-483)."""
    b = C.create_string_buffer(9)
    lib().ckm_decoded_kmer(key, b)
    return b.value


def build_image(nbuckets: int, keys, fI, oI, avg, wt) -> np.ndarray:
    """KmerGuts(dir, nbuckets) + insert_kmer + save_kmer_hash_table (kguts.cc:77-115, 188-234): returns the
    file bytes (header + slots) as a uint8 array."""
    keys = np.ascontiguousarray(keys, np.uint64)
    fI = np.ascontiguousarray(fI, np.int32)
    oI = np.ascontiguousarray(oI, np.int32)
    avg = np.ascontiguousarray(avg, np.uint16)
    wt = np.ascontiguousarray(wt, np.float32)
    img = np.empty(24 + 24 * nbuckets, dtype=np.uint8)
    _check(lib().ckm_image_build(nbuckets, len(keys), keys.ctypes.data, fI.ctypes.data, oI.ctypes.data, avg.ctypes.data,
                                 wt.ctypes.data, img.ctypes.data, img.nbytes))
    return img


def build_image_device(nbuckets: int, keys, fI, oI, avg, wt, device: int = 0) -> np.ndarray:
    """The same file bytes built on the GPU (write_hashtable, build_signature_kmers.cc:860-898)."""
    keys = np.ascontiguousarray(keys, np.uint64)
    fI = np.ascontiguousarray(fI, np.int32)
    oI = np.ascontiguousarray(oI, np.int32)
    avg = np.ascontiguousarray(avg, np.uint16)
    wt = np.ascontiguousarray(wt, np.float32)
    img = np.empty(24 + 24 * nbuckets, dtype=np.uint8)
    L = lib()
    L.ckm_image_build_device.argtypes = [C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6 + [C.c_size_t]
    _check(L.ckm_image_build_device(device, nbuckets, len(keys), keys.ctypes.data, fI.ctypes.data, oI.ctypes.data, avg.ctypes.data,
                                    wt.ctypes.data, img.ctypes.data, img.nbytes))
    return img


def signature_weight(NSF, KS, NSi, NFj, NSiFj) -> float:
    """compute_weight_of_signature (build_signature_kmers.cc:841-853)."""
    L = lib()
    L.ckm_signature_weight.restype = C.c_float
    L.ckm_signature_weight.argtypes = [C.c_float] * 5
    return L.ckm_signature_weight(NSF, KS, NSi, NFj, NSiFj)


def save_kmer_hash_table(image: np.ndarray, kmer_dir: str) -> None:
    image.tofile(os.path.join(kmer_dir, "kmer.table.mem_map"))


class KmerPegMapping:
    """The peg-id table of KmerPegMapping (kmer.h:109-116, kmer.cc:272-295): ids in order of first encode_id."""

    def __init__(self):
        self._m = C.c_void_p(lib().ckm_mapping_new())

    def encode_id(self, peg) -> int:
        return lib().ckm_mapping_encode_id(self._m, peg.encode() if isinstance(peg, str) else peg)

    def decode_id(self, i: int) -> str:
        return lib().ckm_mapping_decode_id(self._m, i).decode()

    def __del__(self):
        if getattr(self, "_m", None):
            lib().ckm_mapping_free(self._m)
            self._m = None


def merge_pairs(pairs: np.ndarray) -> np.ndarray:
    """Order COO entries by (eid_i, eid_j) like the reference's std::map and sum equal keys."""
    pairs = np.ascontiguousarray(pairs, PAIR_DT).copy()
    n = lib().ckm_matrix_merge_pairs(pairs.ctypes.data, len(pairs))
    return pairs[:n]


SCORE_DT = np.dtype([("id", "<u4"), ("hit_count", "<u4"), ("weighted_total", "<f4")])


class FamilyScoresC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("scores", C.c_void_p), ("score_offsets", C.c_void_p), ("best", C.c_void_p), ("matches", C.c_void_p)]


class FamilyDataC(C.Structure):
    _fields_ = [("pgf", C.c_char_p), ("plf", C.c_char_p), ("function", C.c_char_p), ("genus_id", C.c_uint64), ("total_size", C.c_uint64),
                ("count", C.c_uint16)]


class LookupOptionsC(C.Structure):
    _fields_ = [("family_mode", C.c_int), ("kmer_hit_threshold", C.c_uint), ("find_best_match", C.c_int), ("find_reps", C.c_int),
                ("allow_ambiguous_functions", C.c_int), ("target_genus_id", C.c_uint64)]


class SeqBatchC(C.Structure):
    _fields_ = [("n", C.c_uint32), ("ids", C.POINTER(C.c_char_p)), ("residues", C.c_void_p), ("offsets", C.POINTER(C.c_uint64)),
                ("n_errors", C.c_uint64)]


class SeqParser:
    """Streaming FASTA / FASTQ parser of the request front end (include/ckm_server.h); host-only."""

    def __init__(self, fastq: bool = False):
        L = lib()
        L.ckm_seq_parser_new.restype = C.c_void_p
        L.ckm_seq_parser_new.argtypes = [C.c_int]
        L.ckm_seq_parser_free.argtypes = [C.c_void_p]
        L.ckm_seq_parser_feed.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.ckm_seq_parser_complete.argtypes = [C.c_void_p]
        L.ckm_seq_parser_pending.restype = C.c_uint64
        L.ckm_seq_parser_pending.argtypes = [C.c_void_p]
        L.ckm_seq_parser_take.argtypes = [C.c_void_p, C.POINTER(SeqBatchC)]
        L.ckm_seq_parser_last_error.restype = C.c_char_p
        L.ckm_seq_parser_last_error.argtypes = [C.c_void_p]
        self._p = C.c_void_p(L.ckm_seq_parser_new(int(fastq)))
        self.n_errors = 0

    def feed(self, data: bytes):
        lib().ckm_seq_parser_feed(self._p, data, len(data))

    def complete(self):
        lib().ckm_seq_parser_complete(self._p)

    def pending(self) -> int:
        return lib().ckm_seq_parser_pending(self._p)

    def take(self) -> list:
        """[(id, sequence)] completed since the last take, as bytes."""
        b = SeqBatchC()
        lib().ckm_seq_parser_take(self._p, C.byref(b))
        self.n_errors = b.n_errors
        res = C.string_at(b.residues, b.offsets[b.n]) if b.n else b""
        return [(b.ids[i], res[b.offsets[i]:b.offsets[i + 1]]) for i in range(b.n)]

    def last_error(self) -> str:
        return lib().ckm_seq_parser_last_error(self._p).decode(errors="replace")

    def __del__(self):
        try:
            lib().ckm_seq_parser_free(self._p)
        except Exception:
            pass


def http_describe(head: bytes) -> dict:
    """ckm_http_describe as a dict (repeated keys keep the last value)."""
    L = lib()
    L.ckm_http_describe.restype = C.c_void_p
    L.ckm_http_describe.argtypes = [C.c_char_p, C.c_size_t]
    L.ckm_free_text.argtypes = [C.c_void_p]
    p = L.ckm_http_describe(head, len(head))
    text = C.string_at(p).decode(errors="replace")
    L.ckm_free_text(p)
    return dict(line.split("=", 1) for line in text.split("\n") if "=" in line)


def pinned_empty(nbytes: int) -> np.ndarray:
    """A uint8 array in page-locked host memory (ckm_host_alloc); it is released when the array is collected."""
    p = C.c_void_p()
    _check(lib().ckm_host_alloc(C.byref(p), max(nbytes, 1)))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, np.uint8)[:nbytes]
    import weakref
    weakref.finalize(buf, lib().ckm_host_free, p)
    return arr


def canonical_family_csr(kmers, fam_off, fam_ids) -> tuple:
    """Sort a k-mer -> family-list CSR by k-mer and each list by family id (the lists are sets: kmer.cc:216-230)."""
    kmers = np.asarray(kmers, np.uint64)
    fam_off = np.asarray(fam_off, np.uint64)
    fam_ids = np.asarray(fam_ids, np.uint32)
    cnt = np.diff(fam_off).astype(np.int64)
    owner = np.repeat(kmers, cnt)
    order = np.lexsort((fam_ids, owner))
    ks = np.argsort(kmers, kind="stable")
    off = np.zeros(len(kmers) + 1, np.uint64)
    off[1:] = np.cumsum(cnt[ks])
    return kmers[ks], off, fam_ids[order]


class KmerGuts:
    """Batch mirror of ``KmerGuts(kmer_dir, image)`` (kguts.cc:34-58)."""

    def __init__(self, kmer_dir: str | None = None, image: np.ndarray | None = None, device: int = 0,
                 function_names=None, otu_names=None, built=None):
        """built = (nbuckets, keys, fI, oI, avg, wt): build the table on the GPU and open it (ckm_open_built)."""
        L = lib()
        h = C.c_void_p()
        if built is not None:
            nb, keys, fI, oI, avg, wt = built
            keys = np.ascontiguousarray(keys, np.uint64)
            fI = np.ascontiguousarray(fI, np.int32)
            oI = np.ascontiguousarray(oI, np.int32)
            avg = np.ascontiguousarray(avg, np.uint16)
            wt = np.ascontiguousarray(wt, np.float32)
            fn = [s.encode() for s in (function_names or [])]
            on = [s.encode() for s in (otu_names or [])]
            fa = (C.c_char_p * len(fn))(*fn)
            oa = (C.c_char_p * len(on))(*on)
            L.ckm_open_built.argtypes = [C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 5 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                                                                           C.POINTER(C.c_void_p)]
            _check(L.ckm_open_built(device, nb, len(keys), keys.ctypes.data, fI.ctypes.data, oI.ctypes.data, avg.ctypes.data,
                                    wt.ctypes.data, fa, len(fn), oa, len(on), C.byref(h)))
        elif image is not None:
            fn = [s.encode() for s in (function_names or [])]
            on = [s.encode() for s in (otu_names or [])]
            fa = (C.c_char_p * len(fn))(*fn)
            oa = (C.c_char_p * len(on))(*on)
            _check(L.ckm_open_image(image.ctypes.data, image.nbytes, device, fa, len(fn), oa, len(on), C.byref(h)))
        else:
            _check(L.ckm_open(kmer_dir.encode(), device, C.byref(h)))
        self._h = h
        self.device = device

    def clone(self) -> "KmerGuts":
        """A second engine over the same tables (one KmerGuts per worker thread, threadpool.cc:33); close it first."""
        L = lib()
        L.ckm_clone.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        h = C.c_void_p()
        _check(L.ckm_clone(self._h, C.byref(h)))
        other = object.__new__(KmerGuts)
        other._h = h
        other.device = self.device
        return other

    def close(self):
        if getattr(self, "_h", None) and _lib is not None and lib is not None:  # (module globals are gone at interpreter exit)
            _lib.ckm_close(self._h)
            self._h = None

    __del__ = close

    # -- Q1 ------------------------------------------------------------------------------------------
    _PARAM_NAMES = ("order_constraint", "min_hits", "min_weighted_hits", "max_gap")

    def set_default_parameters(self):
        lib().ckm_set_default_params(self._h)

    def set_parameters(self, params: dict):
        """KmerGuts::set_parameters (kguts.cc:244-268): reset to defaults, then apply the integer-valued
        entries whose key is one of the four engine parameters; non-integers are warned about and ignored."""
        cur = dict(order_constraint=0, min_hits=5, min_weighted_hits=0, max_gap=200)
        for k, v in params.items():
            if k in cur:
                try:
                    cur[k] = _stoi(v)
                except ValueError:
                    import sys
                    print(f"Warning: invalid integer '{v}' passed for parameter {k}", file=sys.stderr)
        _check(lib().ckm_set_params(self._h, cur["order_constraint"], cur["min_hits"], cur["min_weighted_hits"], cur["max_gap"]))

    def get_parameters(self) -> dict:
        v = [C.c_int() for _ in range(4)]
        lib().ckm_get_params(self._h, *[C.byref(x) for x in v])
        return dict(zip(self._PARAM_NAMES, (x.value for x in v)))

    # -- metadata --------------------------------------------------------------------------------------
    def function_at_index(self, i: int) -> str:
        return lib().ckm_function_at_index(self._h, i).decode()

    def otu_at_index(self, i: int) -> str:
        return lib().ckm_otu_at_index(self._h, i).decode()

    @property
    def function_count(self) -> int:
        return lib().ckm_function_count(self._h)

    @property
    def num_sigs(self) -> int:
        return lib().ckm_num_sigs(self._h)

    @property
    def slot_bytes(self) -> int:
        return lib().ckm_table_slot_bytes(self._h)

    @property
    def l2_fetch_granularity(self) -> int:
        return lib().ckm_l2_fetch_granularity(self._h)

    @property
    def has_occupancy_bitmap(self) -> bool:
        return bool(lib().ckm_has_occupancy_bitmap(self._h))

    @property
    def table_buckets(self) -> int:
        """Buckets of the table in HBM (ckm_table_buckets): the image's, or the power of two of the library's own table."""
        return int(lib().ckm_table_buckets(self._h))

    @property
    def copy_state(self) -> dict:
        """Automatic fall-back of K1 from the neighbour copy to plain hash probing (ckm_copy_state)."""
        st = (C.c_uint32 * 3)()
        _check(lib().ckm_copy_state(self._h, st))
        return {"suspended": bool(st[0]), "retry_in": int(st[1]), "suspensions": int(st[2])}

    @property
    def chain_info(self) -> dict:
        """Neighbour-ordered copy of the table (ckm_chain.cuh): entries, chains, build time, hits it answered last batch."""
        info = (C.c_uint64 * 4)()
        _check(lib().ckm_chain_info(self._h, info))
        return {"entries": int(info[0]), "chains": int(info[1]), "build_ms": info[2] / 1000.0, "hits_from_copy": int(info[3])}

    @property
    def stream(self) -> int:
        return lib().ckm_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return lib().ckm_launch_count(self._h)

    def profile_enable(self, on: bool = True):
        lib().ckm_profile_enable(self._h, int(on))

    def profile_read(self):
        """(probe_ms, scan_ms, batches) summed since the last read; device-timed with CUDA events."""
        p, s, n = C.c_double(), C.c_double(), C.c_uint64()
        _check(lib().ckm_profile_read(self._h, C.byref(p), C.byref(s), C.byref(n)))
        return p.value, s.value, n.value

    def calibrate_gather(self, nbytes=16, unroll=4, rounds=64, blocks_per_sm=8):
        """Independent random reads over the resident table: (accesses/s, ms)."""
        r, ms = C.c_double(), C.c_double()
        _check(lib().ckm_calibrate_gather(self._h, nbytes, unroll, rounds, blocks_per_sm, C.byref(r), C.byref(ms)))
        return r.value, ms.value

    @property
    def last_batch_was_fused(self) -> bool:
        return bool(lib().ckm_last_batch_was_fused(self._h))

    def set_tuning(self, bits: int):
        lib().ckm_set_tuning(self._h, bits)

    def synchronize(self):
        _check(lib().ckm_synchronize(self._h))

    # -- S4 (+B1) over a batch ---------------------------------------------------------------------------
    def process_aa_seq_batch(self, residues: np.ndarray, offsets: np.ndarray, flags: int) -> dict:
        """process_aa_seq / process_aa_seq_hits (+ find_best_call) for every sequence of the batch
        (kguts.cc:879-908, 1008-1199).  Host arrays in, numpy copies of the CSR results out."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        o = BatchOutC()
        _check(lib().ckm_call_batch(self._h, residues.ctypes.data, offsets.ctypes.data, n, flags, C.byref(o)))
        r = {"n": n, "n_probes": o.n_probes, "n_hits": o.n_hits}
        for name, dt in (("call", CALL_DT), ("hit", HIT_DT), ("otu", OTU_DT)):
            offp = getattr(o, f"{name}_offsets")
            if offp:
                off = _arr(offp, n + 1, np.uint64)
                r[f"{name}_offsets"] = off
                r[f"{name}s"] = _arr(getattr(o, f"{name}s"), int(off[-1]), dt)
        if o.best:
            r["best"] = _arr(o.best, n, BEST_DT)
        return r

    def process_packed_batch(self, packed: np.ndarray, word_offsets: np.ndarray, flags: int) -> dict:
        """process_aa_seq_batch for residues packed seven to a 32-bit word (pack_residues; ckm_call_batch_packed)."""
        packed = np.ascontiguousarray(packed, np.uint32)
        word_offsets = np.ascontiguousarray(word_offsets, np.uint64)
        n = len(word_offsets) - 1
        o = BatchOutC()
        _check(lib().ckm_call_batch_packed(self._h, packed.ctypes.data, word_offsets.ctypes.data, n, flags, C.byref(o)))
        r = {"n": n, "n_probes": o.n_probes, "n_hits": o.n_hits}
        for name, dt in (("call", CALL_DT), ("hit", HIT_DT), ("otu", OTU_DT)):
            offp = getattr(o, f"{name}_offsets")
            if offp:
                off = _arr(offp, n + 1, np.uint64)
                r[f"{name}_offsets"] = off
                r[f"{name}s"] = _arr(getattr(o, f"{name}s"), int(off[-1]), dt)
        if o.best:
            r["best"] = _arr(o.best, n, BEST_DT)
        return r

    def call_batch_packed_raw(self, packed_ptr: int, word_offsets_ptr: int, n: int, flags: int) -> BatchOutC:
        o = BatchOutC()
        _check(lib().ckm_call_batch_packed(self._h, packed_ptr, word_offsets_ptr, n, flags, C.byref(o)))
        return o

    def call_batch_raw(self, residues_ptr: int, offsets_ptr: int, n: int, flags: int) -> BatchOutC:
        """Same call with caller-owned (ideally pinned) host pointers and no result copies: for timing."""
        o = BatchOutC()
        _check(lib().ckm_call_batch(self._h, residues_ptr, offsets_ptr, n, flags, C.byref(o)))
        return o

    def call_batch_device(self, d_residues: int, d_offsets: int, n: int, total: int, max_len: int, flags: int):
        _check(lib().ckm_call_batch_device(self._h, d_residues, d_offsets, n, total, max_len, flags))

    def device_results(self) -> DeviceOutC:
        o = DeviceOutC()
        _check(lib().ckm_device_results(self._h, C.byref(o)))
        return o

    def read_totals(self):
        """(probes, hits, calls) of the last batch."""
        t = (C.c_uint64 * 3)()
        _check(lib().ckm_read_totals(self._h, t))
        return int(t[0]), int(t[1]), int(t[2])

    # -- request handlers (include/ckm_handlers.h) -----------------------------------------------------
    def _take_text(self, p) -> str:
        s = C.string_at(p).decode()
        lib().ckm_free_text(p)
        return s

    def query_text(self, ids, residues, offsets, details=0, find_best_call=0) -> str:
        """Response text of POST /query for one chunk (query_request.cc:103-152)."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        arr = (C.c_char_p * len(bs))(*bs)
        t = C.c_void_p()
        _check(lib().ckm_query_text(self._h, arr, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, details,
                                    find_best_call, C.byref(t)))
        return self._take_text(t)

    # -- family voting (FamilyMapper) ------------------------------------------------------------------
    def family_load(self, kmers, fam_off, fam_ids, pgf, plf, function):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        fam_off = np.ascontiguousarray(fam_off, np.uint64)
        fam_ids = np.ascontiguousarray(fam_ids, np.uint32)
        mk = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])
        _check(lib().ckm_family_load(self._h, len(kmers), kmers.ctypes.data, fam_off.ctypes.data, fam_ids.ctypes.data, len(pgf),
                                     mk(pgf), mk(plf), mk(function)))

    # family-mode start-up load (NRLoader::thread_load + KmerInserter + add_fam_mapping), on the GPU
    def family_nr_begin(self):
        _check(lib().ckm_family_nr_begin(self._h))

    def family_nr_add(self, fam_ids, residues, offsets):
        """One chunk of families.nr: fam_ids[i] = family of sequence i, 0xFFFFFFFF = none (ends the chunk, nr_loader.cc:159)."""
        fam_ids = np.ascontiguousarray(fam_ids, np.uint32)
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        _check(lib().ckm_family_nr_add(self._h, fam_ids.ctypes.data, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def family_nr_finish(self, pgf, plf, function) -> tuple:
        """Install the collected table; returns (n_kmers, n_entries)."""
        mk = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])
        nk, ne = C.c_uint64(), C.c_uint64()
        _check(lib().ckm_family_nr_finish(self._h, len(pgf), mk(pgf), mk(plf), mk(function), C.byref(nk), C.byref(ne)))
        self._fam_size = (nk.value, ne.value)
        return self._fam_size

    def family_export(self, n_kmers, n_entries) -> tuple:
        """(kmers, fam_off, fam_ids) of the installed table, k-mers ascending and every list ascending."""
        k = np.zeros(n_kmers, np.uint64)
        o = np.zeros(n_kmers + 1, np.uint64)
        ids = np.zeros(max(n_entries, 1), np.uint32)
        _check(lib().ckm_family_export(self._h, n_kmers, n_entries, k.ctypes.data, o.ctypes.data, ids.ctypes.data))
        return canonical_family_csr(k, o, ids[:n_entries])

    def family_scores(self, residues, offsets) -> dict:
        """LookupRequest's per-sequence (family, hit_count, weighted_total) lists + best call + FamilyMapper match."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        L = lib()
        L.ckm_family_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(FamilyScoresC)]
        o = FamilyScoresC()
        _check(L.ckm_family_scores(self._h, residues.ctypes.data, offsets.ctypes.data, n, C.byref(o)))
        off = _arr(o.score_offsets, n + 1, np.dtype("<u8"))
        return dict(score_offsets=off, scores=_arr(o.scores, int(off[-1]), SCORE_DT), best=_arr(o.best, n, BEST_DT),
                    matches=_arr(o.matches, n, FAMILY_DT))

    def postings_scores(self, residues, offsets) -> tuple:
        """(pairs, pair_offsets): per sequence, the pegs of the selected postings sharing hit k-mers, with counts."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        L = lib()
        L.ckm_postings_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        p, po = C.c_void_p(), C.c_void_p()
        _check(L.ckm_postings_scores(self._h, residues.ctypes.data, offsets.ctypes.data, n, C.byref(p), C.byref(po)))
        off = _arr(po.value, n + 1, np.dtype("<u8"))
        return _arr(p.value, int(off[-1]), PAIR_DT), off

    def postings_select(self, key: int):
        L = lib()
        L.ckm_postings_select.argtypes = [C.c_void_p, C.c_uint32]
        _check(L.ckm_postings_select(self._h, key))

    def lookup_text(self, ids, residues, offsets, families=None, mapping=None, family_mode=True, kmer_hit_threshold=3,
                    find_best_match=False, find_reps=False, allow_ambiguous_functions=False, target_genus_id=0) -> str:
        """POST /lookup for one chunk.  families: list of (pgf, plf, function, genus_id, total_size, count)."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        fams = families or []
        arr = (FamilyDataC * max(len(fams), 1))()
        for k, f in enumerate(fams):
            arr[k] = FamilyDataC(f[0].encode(), f[1].encode(), f[2].encode(), f[3], f[4], f[5])
        opt = LookupOptionsC(int(family_mode), kmer_hit_threshold, int(find_best_match), int(find_reps), int(allow_ambiguous_functions),
                             target_genus_id)
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        idarr = (C.c_char_p * max(len(bs), 1))(*bs)
        L = lib()
        L.ckm_lookup_text.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint32, C.POINTER(C.c_void_p)]
        t = C.c_void_p()
        _check(L.ckm_lookup_text(self._h, mapping._m if mapping is not None else None, arr, len(fams), C.byref(opt), idarr,
                                 residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, C.byref(t)))
        return self._take_text(t)

    def find_all_matches_text(self, ids, residues, offsets, families) -> str:
        """FamilyMapper::find_all_matches (family_mapper.cc:207-285) for one chunk; families as for lookup_text."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        arr = (FamilyDataC * max(len(families), 1))()
        for k, f in enumerate(families):
            arr[k] = FamilyDataC(f[0].encode(), f[1].encode(), f[2].encode(), f[3], f[4], f[5])
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        idarr = (C.c_char_p * max(len(bs), 1))(*bs)
        L = lib()
        L.ckm_family_all_matches_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                  C.POINTER(C.c_void_p)]
        t = C.c_void_p()
        _check(L.ckm_family_all_matches_text(self._h, arr, len(families), idarr, residues.ctypes.data, offsets.ctypes.data,
                                             len(offsets) - 1, C.byref(t)))
        return self._take_text(t)

    def find_best_family_match_batch(self, residues, offsets) -> np.ndarray:
        """FamilyMapper::find_best_family_match (family_mapper.cc:65-205) for every sequence."""
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        p = C.c_void_p()
        _check(lib().ckm_family_batch(self._h, residues.ctypes.data, offsets.ctypes.data, n, C.byref(p)))
        return _arr(p.value, n, FAMILY_DT)

    def family_names(self, m) -> tuple:
        """(gfam_id, lfam_id, function) strings of best_match_t for one FAMILY_DT record."""
        L = lib()
        rec = np.array([m], dtype=FAMILY_DT)
        return (L.ckm_family_pgf_name(self._h, int(m["gfam"])).decode(), L.ckm_family_plf_name(self._h, int(m["lfam"])).decode(),
                L.ckm_family_function_name(self._h, rec.ctypes.data).decode())

    def family_text(self, residues, offsets) -> str:
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        t = C.c_void_p()
        _check(lib().ckm_family_text(self._h, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, C.byref(t)))
        return self._take_text(t)

    # -- fastq reads (FqProcessRequest / DNASequence / TranslationTable) ---------------------------------
    def get_possible_proteins_batch(self, bases, offsets, min_len=0):
        """DNASequence::get_possible_proteins (dna_seq.cc:9-23) for every read, fragments longer than min_len:
        list per read of [(frame, [fragment bytes, ...]) x 6]."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        o = FqFragmentsC()
        _check(lib().ckm_fq_translate(self._h, bases.ctypes.data, offsets.ctypes.data, n, min_len, C.byref(o)))
        ffo = _arr(o.frag_frame_offsets, 6 * n + 1, np.uint64)
        fo = _arr(o.frag_offsets, o.n_fragments + 1, np.uint64)
        res = _arr(o.residues, int(fo[-1]) if len(fo) else 0, np.uint8).tobytes()
        out = []
        for r in range(n):
            frames = []
            for s, f in enumerate((1, 2, 3, -1, -2, -3)):
                a, b = int(ffo[6 * r + s]), int(ffo[6 * r + s + 1])
                frames.append((f, [res[int(fo[k]):int(fo[k + 1])] for k in range(a, b)]))
            out.append(frames)
        return out

    def fq_batch(self, bases, offsets) -> dict:
        """FqProcessRequest::on_parsed_seq (fq_process_request.cc:298-365) for every read."""
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        o = FqOutC()
        _check(lib().ckm_fq_batch(self._h, bases.ctypes.data, offsets.ctypes.data, n, C.byref(o)))
        off = _arr(o.match_offsets, n + 1, np.uint64)
        return {"n": n, "best_frame": _arr(o.best_frame, n, np.int32), "best_score": _arr(o.best_score, n, np.float64),
                "match_offsets": off, "matches": _arr(o.matches, int(off[-1]), FQ_MATCH_DT), "n_fragments": o.n_fragments,
                "n_probes": o.n_probes}

    def fq_text(self, ids, bases, offsets) -> str:
        bases = np.ascontiguousarray(bases, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        arr = (C.c_char_p * len(bs))(*bs)
        t = C.c_void_p()
        _check(lib().ckm_fq_text(self._h, arr, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, C.byref(t)))
        return self._take_text(t)

    # -- /add postings and /matrix ------------------------------------------------------------------------
    def postings_add(self, eids, residues, offsets):
        eids = np.ascontiguousarray(eids, np.uint32)
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        _check(lib().ckm_postings_add(self._h, eids.ctypes.data, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def postings_clear(self):
        lib().ckm_postings_clear(self._h)

    @property
    def postings_count(self) -> int:
        return lib().ckm_postings_count(self._h)

    def matrix_rows(self, eids, residues, offsets, row_begin=0, row_end=None) -> np.ndarray:
        """Rows [row_begin, row_end) of MatrixRequest's lower-triangular count matrix as unordered COO entries."""
        eids = np.ascontiguousarray(eids, np.uint32)
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        p, npairs = C.c_void_p(), C.c_uint64()
        _check(lib().ckm_matrix_rows(self._h, eids.ctypes.data, residues.ctypes.data, offsets.ctypes.data, n, row_begin,
                                     n if row_end is None else row_end, C.byref(p), C.byref(npairs)))
        return _arr(p.value, npairs.value, PAIR_DT)

    def postings_device(self) -> tuple:
        """(device pointer of the k-mers, device pointer of the peg ids, count) of the postings this ctx holds."""
        k, e, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(lib().ckm_postings_device(self._h, C.byref(k), C.byref(e), C.byref(n)))
        return k.value or 0, e.value or 0, n.value

    def postings_import_device(self, d_keys: int, d_eids: int, n: int):
        _check(lib().ckm_postings_import_device(self._h, d_keys, d_eids, n))

    def matrix_rows_device(self, eids, residues, offsets, row_begin=0, row_end=None) -> tuple:
        """ckm_matrix_rows with the tile left in HBM, rows ordered by partner id: (device pointer, pairs, postings walked)."""
        eids = np.ascontiguousarray(eids, np.uint32)
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = len(offsets) - 1
        p, npairs, walked = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(lib().ckm_matrix_rows_device(self._h, eids.ctypes.data, residues.ctypes.data, offsets.ctypes.data, n, row_begin,
                                            n if row_end is None else row_end, C.byref(p), C.byref(npairs), C.byref(walked)))
        return p.value or 0, npairs.value, walked.value

    def add_text(self, mapping: KmerPegMapping, ids, residues, offsets, silent=0) -> str:
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        arr = (C.c_char_p * len(bs))(*bs)
        t = C.c_void_p()
        _check(lib().ckm_add_text(self._h, mapping._m, arr, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1, silent,
                                  C.byref(t)))
        return self._take_text(t)

    def matrix_text(self, mapping: KmerPegMapping, ids, residues, offsets) -> str:
        residues = np.ascontiguousarray(residues, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.uint64)
        bs = [s.encode() if isinstance(s, str) else s for s in ids]
        arr = (C.c_char_p * len(bs))(*bs)
        t = C.c_void_p()
        _check(lib().ckm_matrix_text(self._h, mapping._m, arr, residues.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                     C.byref(t)))
        return self._take_text(t)

    def best_function(self, best_rec) -> str:
        """The `function` string find_best_call returns (kguts.cc:1160, 1176-1196)."""
        if best_rec["flags"] & BEST_AMBIG:
            f1 = self.function_at_index(int(best_rec["ambig_a"]))
            f2 = self.function_at_index(int(best_rec["ambig_b"]))
            if f2.encode() > f1.encode():
                f1, f2 = f2, f1
            return f1 + " ?? " + f2
        if best_rec["function_index"] >= 0:
            return self.function_at_index(int(best_rec["function_index"]))
        return ""


def _stoi(v) -> int:
    """std::stoi semantics for the cases set_parameters meets: leading whitespace, optional sign, digits,
    trailing junk ignored; nothing parseable -> ValueError."""
    import re
    m = re.match(r"\s*([+-]?\d+)", str(v))
    if not m:
        raise ValueError(v)
    return int(m.group(1))
