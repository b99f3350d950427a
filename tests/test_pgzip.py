"""The parallel gzip reader of the FASTQ front end (csrc/hlm_pgzip.h) against zlib: the same text
for every kind of gzip file - compression levels (stored, fixed and dynamic blocks), several
members, sync-flush points, incompressible data, highly repetitive data - whatever the chunk size
and thread count; a truncated or damaged file is an error.  The reference reads .gz input through
boost's gzip_decompressor (src/preselect.cpp:807-831)."""
import ctypes as C
import gzip
import random
import zlib

import pytest

from hla_mapper_b200 import api


def pgzip(path, threads=4, chunk_kb=0):
    f = api.lib().hlm_selftest_pgzip
    f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64),
                  C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    f.restype = C.c_int
    n, crc, nc, ns, zok, sec = C.c_int64(0), C.c_uint32(0), C.c_int64(0), C.c_int64(0), C.c_int(0), C.c_double(0)
    rc = f(str(path).encode(), threads, chunk_kb, C.byref(n), C.byref(crc), C.byref(nc), C.byref(ns), C.byref(zok),
           C.byref(sec))
    return dict(rc=rc, n=n.value, crc=crc.value, chunks=nc.value, slow=ns.value, zlib_ok=bool(zok.value), s=sec.value)


def fastq_text(n, seed, read_len=150):
    rng = random.Random(seed)
    ref = "".join(rng.choice("ACGT") for _ in range(20000))
    out = []
    for i in range(n):
        p = rng.randrange(len(ref) - read_len)
        seq = list(ref[p:p + read_len])
        for _ in range(rng.randrange(3)):
            seq[rng.randrange(read_len)] = rng.choice("ACGT")
        qual = "".join(rng.choice("?????5+'") for _ in range(read_len))
        out.append(f"@A00296:39:HFH3TDSXX:1:{1101 + i // 5000}:{rng.randrange(30000)}:{rng.randrange(30000)} 1:N:0:ACGT\n"
                   f"{''.join(seq)}\n+\n{qual}\n")
    return "".join(out).encode()


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("chunk_kb,threads", [(16, 4), (64, 8), (1, 3), (0, 2)])
def test_same_text_as_zlib(tmp_path, level, chunk_kb, threads):
    raw = fastq_text(12000, level)
    p = tmp_path / "x.fastq.gz"
    p.write_bytes(gzip.compress(raw, level))
    r = pgzip(p, threads, chunk_kb)
    assert r["rc"] == 0 and r["n"] == len(raw) and r["crc"] == zlib.crc32(raw), r
    if chunk_kb in (16, 64) and level > 0:
        assert r["chunks"] > 4          # really cut into chunks
        assert r["slow"] * 4 <= r["chunks"], r  # and most starts were found by the search


def test_members_sync_flushes_and_odd_data(tmp_path):
    rng = random.Random(5)
    noise = bytes(rng.randrange(256) for _ in range(300_000))
    parts = [fastq_text(3000, 1), b"", noise, b"A" * 2_000_000, fastq_text(2000, 2), b"x"]
    blob = b"".join(gzip.compress(x, lv) for x, lv in zip(parts, [6, 6, 6, 9, 1, 6]))
    # a member written with sync flushes (empty stored blocks, as pigz writes them) and a full flush
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    body = fastq_text(4000, 3)
    z = b""
    for k in range(0, len(body), 70_000):
        z += co.compress(body[k:k + 70_000]) + co.flush(zlib.Z_SYNC_FLUSH if k % 140_000 else zlib.Z_FULL_FLUSH)
    z += co.flush()
    p = tmp_path / "m.gz"
    p.write_bytes(blob + z + b"\0" * 7)
    raw = b"".join(parts) + body
    for chunk_kb, threads in [(8, 4), (32, 16), (1024, 2)]:
        r = pgzip(p, threads, chunk_kb)
        assert r["rc"] == 0 and r["n"] == len(raw) and r["crc"] == zlib.crc32(raw), (chunk_kb, r)


def test_truncated_and_damaged_files_are_errors(tmp_path):
    raw = fastq_text(8000, 7)
    z = gzip.compress(raw, 6)
    for name, data in [("cut_half", z[:len(z) // 2]), ("cut_trailer", z[:-5]), ("cut_final", z[:-9])]:
        p = tmp_path / (name + ".gz")
        p.write_bytes(data)
        r = pgzip(p, 4, 16)
        assert r["rc"] == 2 and not r["zlib_ok"], (name, r)
    rng = random.Random(1)
    n_err = 0
    for k in range(12):
        b = bytearray(z)
        at = rng.randrange(20, len(z) - 8)
        b[at] ^= 1 << rng.randrange(8)
        p = tmp_path / f"flip{k}.gz"
        p.write_bytes(bytes(b))
        r = pgzip(p, 4, 16)
        assert r["rc"] in (0, 2)
        assert (r["rc"] == 2) == (not r["zlib_ok"]), (k, at, r)  # an error exactly when zlib reports one
        n_err += r["rc"] == 2
    assert n_err >= 10
    b = bytearray(z)
    b[-6] ^= 0x10  # CRC-32 of the trailer
    p = tmp_path / "crc.gz"
    p.write_bytes(bytes(b))
    assert pgzip(p, 4, 16)["rc"] == 2


def test_crc32_equals_zlib():
    """csrc/hlm_crc32.h: the carry-less-multiplication CRC-32 (used only after it reproduced zlib on
    a test vector in this process) against zlib's crc32() on random buffers"""
    f = api.lib().hlm_selftest_crc32
    f.argtypes = [C.c_int, C.c_uint32, C.POINTER(C.c_int)]
    f.restype = C.c_int
    fast = C.c_int(-1)
    assert f(4000, 7, C.byref(fast)) == 0
    assert fast.value in (0, 1)
