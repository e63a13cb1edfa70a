"""The C-ABI library builds, loads without a GPU, exports every symbol include/*.h declares, and its host-only
entry points (statics, image builder, error paths) behave -- no compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from close_kmers_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("ckm.h", "ckm_handlers.h", "ckm_server.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(ckm_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_every_declared_symbol_is_exported():
    L = api.lib()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_record_layouts_match_header():
    assert api.CALL_DT.itemsize == 20 and api.HIT_DT.itemsize == 32 and api.BEST_DT.itemsize == 28
    assert api.SLOT_DT.itemsize == 24 and api.FAMILY_DT.itemsize == 24 and api.FQ_MATCH_DT.itemsize == 28 and api.PAIR_DT.itemsize == 16


def test_encode_decode_known_answers():
    assert api.encoded_aa_kmer(b"AAAAAAAA") == 0 and api.encoded_aa_kmer(b"AAAAAAAC") == 1
    assert api.encoded_aa_kmer(b"YYYYYYYY") == 20**8 - 1 and api.encoded_aa_kmer(b"CAAAAAAA") == 20**7
    for bad in (b"ACDEFGHX", b"acdefghi", b"ACDE*GHI", b"BJOUZXXX"):
        assert api.encoded_aa_kmer(bad) == 20**8 + 1
    for k in (0, 1, 20**7, 20**8 - 1, 1234567890):
        assert api.encoded_aa_kmer(api.decoded_kmer(k)) == k


def test_image_builder_format_and_limits():
    keys = np.array([api.encoded_aa_kmer(b"ACDEFGHI"), api.encoded_aa_kmer(b"CDEFGHIK"), 20**8 + 1], np.uint64)
    img = api.build_image(3769, keys, [1, 2, 3], [-1, 4, 5], [10, 20, 30], [1.5, 2.5, 3.5])
    hdr = img[:24].view(np.uint64)
    assert tuple(hdr) == (3769, 24, 1) and img.nbytes == 24 + 24 * 3769
    slots = np.frombuffer(img, api.SLOT_DT, offset=24)
    occ = slots[slots["which_kmer"] <= 20**8]
    assert len(occ) == 2  # the invalid key is skipped (kguts.cc:206-210)
    assert np.all(slots["which_kmer"][slots["which_kmer"] > 20**8] == 20**8 + 1)
    s = slots[int(keys[0]) % 3769]
    assert (s["which_kmer"], s["function_index"], s["otu_index"], s["avg_from_end"], s["function_wt"]) == (keys[0], 1, -1, 10, 1.5)
    with pytest.raises(api.CkmError):  # half-full table is refused like kguts.cc:213-216
        api.build_image(16, np.arange(9, dtype=np.uint64), [0] * 9, [0] * 9, [0] * 9, [0.0] * 9)


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a usable device every compute entry point must fail with CKM_ECUDA; on a GPU box this test is moot."""
    protos = synth.make_prototypes(1, 8, 100)
    sig = synth.make_signatures(protos, 500)
    img = api.build_image(synth.bucket_count(len(sig.keys)), sig.keys, sig.fI, sig.oI, sig.avg, sig.wt)
    try:
        g = api.KmerGuts(image=img)
    except api.CkmError as e:
        assert e.code == -4 and "no CPU fallback" in str(e)
    else:
        g.close()


def test_signature_weight_formula(checkers):
    """compute_weight_of_signature: the library's statement against the reference's, compiled by the reference's compiler."""
    import ctypes as C
    if not os.path.exists(checkers.REF_SO):
        pytest.skip("oracle/_ref/libckm_ref.so not built (needs /root/reference)")
    R = checkers.Ref().L
    R.ref_signature_weight.restype = C.c_float
    R.ref_signature_weight.argtypes = [C.c_float] * 5
    rng = np.random.default_rng(4)
    cases = [(1000, 5000, 10, 40, 9), (123456, 9876543, 1, 1, 1), (50, 10, 50, 50, 50), (7, 3, 5, 2, 0)]
    for _ in range(20000):
        nsf = int(rng.integers(1, 10**7))
        nfj = int(rng.integers(1, nsf + 1))
        nsi = int(rng.integers(1, nsf + 1))
        cases.append((nsf, int(rng.integers(1, 10**8)), nsi, nfj, int(rng.integers(0, min(nsi, nfj) + 1))))
    for c in cases:
        got, want = api.signature_weight(*c), R.ref_signature_weight(*c)
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), c


def test_image_header_validation_rejects_crafted_sizes():
    """T2 (kmer_image.cc:87-105) on hostile headers, before any device is touched: a bucket count whose 24 * num_sigs wraps
    around to the file size, and a one-bucket image (floor(2^64 / 1) does not fit the 64-bit multiplier of key % num_sigs)."""
    def header(num_sigs, entry=24, version=1):
        return np.array([num_sigs, entry, version], np.uint64).view(np.uint8)

    body = np.zeros(24 * 4, np.uint8)
    wrapped = (2**64 + 24 * 4) // 24  # 24 * wrapped == 96 (mod 2^64) only if 24 divided 2^64 + 96; search the next one that does
    for k in range(1, 25):
        if (k * 2**64 + 24 * 4) % 24 == 0:
            wrapped = (k * 2**64 + 24 * 4) // 24
            break
    cases = [np.concatenate([header(1), np.zeros(24, np.uint8)]),
             np.concatenate([header(0), np.zeros(0, np.uint8)]),
             np.concatenate([header(5), body]),
             np.concatenate([header(4), body])[:-1]]
    if wrapped < 2**64:
        assert (24 * wrapped) % 2**64 == 24 * 4
        cases.append(np.concatenate([header(wrapped), body]))
    for bad in cases:
        with pytest.raises(api.CkmError) as ei:
            api.KmerGuts(image=np.ascontiguousarray(bad))
        assert ei.value.code == -3, ei.value


def test_packed_residue_format_round_trip():
    """ckm_pack_residues (host code): seven residues to a 32-bit word as base-22 digits, every sequence from a word boundary,
    'end' behind the last residue and from an embedded NUL on (strlen, kguts.cc:791), 'other' for anything that is not one of
    the twenty letters.  Decoded here in numpy."""
    seqs = [b"", b"A", b"ACDEFGH", b"ACDEFGHI", b"ACDEFGHIKLMNPQRSTVWY" * 3, b"MKXacd*-BJOUZ", b"MKV\x00ACDEFGHIKL", b"Y" * 15]
    batch = synth.batch_from_strings(seqs)
    packed, woff = api.pack_residues(batch.residues, batch.offsets)
    assert list(np.diff(woff.astype(np.int64))) == [(len(s) + 6) // 7 for s in seqs] == [api.lib().ckm_packed_words(len(s)) for s in seqs]
    alpha = b"ACDEFGHIKLMNPQRSTVWY"
    for i, s in enumerate(seqs):
        digits = []
        for w in packed[int(woff[i]): int(woff[i + 1])]:
            x = int(w)
            assert x < 22**7
            for _ in range(7):
                digits.append(x % 22)
                x //= 22
        want = []
        ended = False
        for ch in s:
            if ended or ch == 0:
                ended = True
                want.append(21)
            else:
                want.append(alpha.index(bytes([ch])) if bytes([ch]) in alpha else 20)
        want += [21] * (len(digits) - len(want))
        assert digits == want, (s, digits, want)
