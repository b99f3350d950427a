"""The two FASTQ sources of the file pipeline (csrc/hlm_files.cpp) against each other: the mapped
source - several parser threads, batch boundaries from a parallel newline count, the two files of
a batch parsed side by side, batch sizes ramped (B/8, B/8, B/4, B/2, B ... B/4, B/8) - and the
sequential reader must hand over the same records in the same order.  Record rules as the
reference's getline loops (src/preselect.cpp:652-700): four lines per record, a last line without
newline completes its record, an incomplete record at the end is dropped."""
import ctypes as C
import os
import random

import pytest

from hla_mapper_b200 import api


def write_fastq(path, n, seed, mate, last_newline=True, ragged=True):
    rng = random.Random(seed)
    with open(path, "w") as f:
        for i in range(n):
            ln = rng.randint(1, 151) if ragged else 100
            seq = "".join(rng.choice("ACGTN") for _ in range(ln))
            qual = "".join(rng.choice("?5+'") for _ in range(ln))
            end = "" if (i == n - 1 and not last_newline) else "\n"
            f.write(f"@read{i:07d}/{mate} extra field\n{seq}\n+\n{qual}{end}")


def sources(r1, r2, batch, threads, seq1=None, seq2=None):
    f = api.lib().hlm_selftest_fastq_sources
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                  C.c_char_p, C.c_char_p]
    f.restype = C.c_int
    nb, nr = C.c_int64(0), C.c_int64(0)
    rc = f(r1.encode(), r2.encode() if r2 else None, batch, threads, C.byref(nb), C.byref(nr),
           seq1.encode() if seq1 else None, seq2.encode() if seq2 else None)
    return rc, nb.value, nr.value


@pytest.mark.parametrize("n,batch,ramp", [(1, 64, 0), (1000, 64, 0), (1000, 64, 8), (5000, 512, 64), (4097, 4096, 64),
                                          (777, 1 << 19, 0), (20000, 800, 800)])
@pytest.mark.parametrize("last_newline", [True, False])
def test_mapped_source_equals_sequential_reader(tmp_path, monkeypatch, n, batch, ramp, last_newline):
    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    write_fastq(r1, n, 1, 1, last_newline)
    write_fastq(r2, n, 2, 2, last_newline)
    if ramp:
        monkeypatch.setenv("HLM_RAMP_MIN_PAIRS", str(ramp))
    rc, nb, nr = sources(r1, r2, batch, 5)
    assert rc == 0 and nr == n
    if ramp and batch >= ramp and n > 2 * batch:
        assert nb > -(-n // batch)  # the ramp cut more batches than a uniform split
    rc, nb, nr = sources(r1, None, batch, 3)  # single end
    assert rc == 0 and nr == n
    monkeypatch.delenv("HLM_RAMP_MIN_PAIRS", raising=False)


def test_incomplete_last_record_is_dropped(tmp_path):
    r1 = str(tmp_path / "r1.fastq")
    write_fastq(r1, 10, 3, 1)
    with open(r1, "a") as f:
        f.write("@partial\nACGT\n+\n")  # no quality line
    rc, nb, nr = sources(r1, None, 4, 2)
    assert rc == 0 and nr == 10


@pytest.mark.parametrize("reader", ["threads", "zlib"])
def test_gzip_input_gives_the_same_records(tmp_path, monkeypatch, reader):
    """plain gzip files go through the multi-threaded inflater (csrc/hlm_pgzip.h; HLM_GZIP_ZLIB=1:
    zlib's gzread): same records as the plain files, here with chunks of 4 KiB of compressed data"""
    import gzip

    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    write_fastq(r1, 30000, 5, 1)
    write_fastq(r2, 30000, 6, 2)
    for p in (r1, r2):
        with open(p, "rb") as f, open(p + ".gz", "wb") as o:
            o.write(gzip.compress(f.read(), 6))
    monkeypatch.setenv("HLM_PGZIP_CHUNK_KB", "4")
    if reader == "zlib":
        monkeypatch.setenv("HLM_GZIP_ZLIB", "1")
    rc, nb, nr = sources(r1, r2, 7000, 4, r1 + ".gz", r2 + ".gz")
    assert rc == 0 and nr == 30000
    # a truncated file is an error of the reader
    data = open(r2 + ".gz", "rb").read()
    open(r2 + ".gz", "wb").write(data[:len(data) // 2])
    rc, _, _ = sources(r1, r2, 7000, 4, r1 + ".gz", r2 + ".gz")
    assert rc != 0


def test_bgzip_input_gives_the_same_records(tmp_path):
    """bgzip-compressed files (BGZF blocks inflated on several threads, ahead of the parser:
    csrc/hlm_bgzf.h): same records as the plain files; a truncated file is an error"""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import bamlite

    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    write_fastq(r1, 30000, 8, 1)
    write_fastq(r2, 30000, 9, 2)
    for p in (r1, r2):
        with open(p, "rb") as f, open(p + ".gz", "wb") as o:
            o.write(bamlite.bgzf_deflate(f.read(), block=20000))
    rc, nb, nr = sources(r1, r2, 7000, 4, r1 + ".gz", r2 + ".gz")
    assert rc == 0 and nr == 30000
    data = open(r1 + ".gz", "rb").read()
    open(r1 + ".gz", "wb").write(data[:len(data) // 3])
    rc, _, _ = sources(r1, r2, 7000, 4, r1 + ".gz", r2 + ".gz")
    assert rc != 0
