"""The parallel id sort of the file pipeline (csrc/hlm_files.cpp: sort_keys - sample sort with a
stable scatter) against one std::stable_sort, through the library's self-test entry point.  The
order is the reference's std::map order of the read ids (src/preselect.cpp:1138-1149)."""
import ctypes as C

import pytest

from hla_mapper_b200 import api


@pytest.mark.parametrize("n,threads", [(0, 4), (1, 4), (63, 8), (64, 2), (1000, 3), (5000, 16), (200_000, 8),
                                       (300_001, 32)])
def test_sample_sort_equals_stable_sort(n, threads):
    lib = api.lib()
    f = lib.hlm_selftest_sort_keys
    f.argtypes = [C.c_int64, C.c_uint32, C.c_int]
    f.restype = C.c_int
    for seed in (1, 2, 3):
        assert f(n, seed, threads) == 0, (n, threads, seed)
