"""include/*.h are valid C, and a plain-C caller links against libckm.so; on a GPU box it also runs end to end."""
import os
import subprocess

import pytest

from close_kmers_b200 import api, build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    lib = build.build()
    exe = os.path.join(str(tmp_path), "query_chunk")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "query_chunk.c"), "-L", os.path.dirname(lib), "-lckm",
                    f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", exe], check=True)
    return exe


def test_c_caller_compiles_and_links(tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    # no image there (and no GPU on the CPU box): a clean error through ckm_last_error, never a crash
    assert r.returncode == 1 and "ckm_open:" in r.stderr


@pytest.mark.gpu
def test_c_caller_runs(tmp_path, checkers):
    import workloads as wl
    exe = _compile(tmp_path)
    protos, sig, img = wl.small_world()
    d = str(tmp_path)
    api.save_kmer_hash_table(img, d)
    synth.write_index_files(d, sig.n_functions, 12)
    r = subprocess.run([exe, d], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("PROTEIN-ID\t") == 2 and r.stdout.count("OTU-COUNTS\t") == 2
