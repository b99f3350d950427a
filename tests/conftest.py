"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200.

The oracle (oracle/) is the checker in every test; the product (hla_mapper_b200/) is only
ever called through the C-ABI of libhlamap.so."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    import oracle

    oracle.build(ref=os.path.isdir("/root/reference/src"))
    return oracle


@pytest.fixture(scope="session")
def small_db(tmp_path_factory):
    from hla_mapper_b200 import synth

    d = tmp_path_factory.mktemp("db_small")
    return synth.make_db(str(d / "db"), seed=7, n_loci=6, total_alleles=120, len_range=(600, 800),
                         n_others=20)


@pytest.fixture(scope="session")
def small_reads(small_db):
    from hla_mapper_b200 import synth

    return synth.make_reads(small_db, 1500, read_len=150, seed=11, on_target=0.7, frag_mu=300,
                            frag_sd=40)


@pytest.fixture(scope="session")
def hostsim():
    import ctypes as C

    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libhostsim.so"))
    u8p = C.POINTER(C.c_uint8)
    L.hs_myers.argtypes = [u8p, C.c_int, u8p, C.c_int]
    L.hs_dp.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
    L.hs_seed.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.POINTER(C.c_uint32)]
    L.hs_db_open.restype = C.c_void_p
    L.hs_db_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.hs_db_close.argtypes = [C.c_void_p]
    L.hs_db_stat.restype = C.c_int64
    L.hs_db_stat.argtypes = [C.c_void_p, C.c_int]
    L.hs_db_kmer.restype = C.c_uint32
    L.hs_db_kmer.argtypes = [C.c_void_p, C.c_char_p]
    L.hs_db_candidates.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int]
    L.hs_db_candidates_brute.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int]
    L.hs_db_tile_info.restype = C.c_uint32
    L.hs_db_tile_info.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    L.hs_db_eval.argtypes = [C.c_void_p, u8p, C.c_int, C.c_uint32, C.POINTER(C.c_int)]
    return L


