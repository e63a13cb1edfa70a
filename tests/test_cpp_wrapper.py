"""include/ckm.hpp (the C++ mirror of KmerGuts / FamilyMapper) compiles, and on a GPU box the reference's handler loop
written against it reproduces the reference's response text."""
import os
import subprocess

import pytest

from close_kmers_b200 import api, build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    lib = build.build()
    exe = os.path.join(str(tmp_path), "wrapper_check")
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "wrapper_check.cc"), "-L", os.path.dirname(lib), "-lckm",
                    f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", exe], check=True)
    return exe


def test_wrapper_compiles_and_reports_errors(tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, str(tmp_path), "0", "0"], input="a\tACDEFGHIKL\n", capture_output=True, text=True)
    assert r.returncode == 1 and "libckm:" in r.stderr  # no image in that directory (and no GPU here): a C++ exception


@pytest.mark.gpu
def test_wrapper_loop_matches_reference_text(tmp_path, checkers):
    import workloads as wl
    exe = _compile(tmp_path)
    protos, sig, img = wl.small_world()
    d = str(tmp_path)
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    batch = synth.make_proteins(5, protos, 300)
    ids = [f"fig|1.1.peg.{i}" for i in range(batch.n)]
    work = "".join(f"{ids[i]}\t{batch.seq(i).decode()}\n" for i in range(batch.n))
    guts = api.KmerGuts(kmer_dir=d)
    ref = checkers.Ref().open(d) if os.path.exists(checkers.REF_SO) else None
    for details, fbc in ((0, 0), (1, 0), (0, 1)):
        r = subprocess.run([exe, d, str(details), str(fbc)], input=work, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == guts.query_text(ids, batch.residues, batch.offsets, details, fbc)
        if ref is not None:
            ref.set_params()
            assert r.stdout == ref.query_text(ids, batch, details, fbc)
    guts.close()


def test_text_builder_prints_like_ostream(tmp_path):
    """The handlers build responses with ckm_text::Text (to_chars); the reference uses std::ostream: same bytes."""
    exe = os.path.join(str(tmp_path), "text_check")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "text_check.cc"), "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
