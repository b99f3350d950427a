"""Host-only measurement: records/s of the sequential-reader FASTQ source (one reader thread per
file; gzip and bgzip inflated on several threads) over r1 [r2].  No GPU needed.
    python tools/reader_bench.py r1.fastq[.gz] [r2.fastq[.gz]]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hla_mapper_b200 import api


def main():
    r1 = sys.argv[1]
    r2 = sys.argv[2] if len(sys.argv) > 2 else None
    f = api.lib().hlm_selftest_reader_speed
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    f.restype = C.c_int
    for rep in range(3):
        s, n, b = C.c_double(0), C.c_int64(0), C.c_int64(0)
        rc = f(r1.encode(), r2.encode() if r2 else None, 1 << 19, C.byref(s), C.byref(n), C.byref(b))
        print(f"rc {rc}: {n.value} records, {b.value / 1e6:.0f} M bases in {s.value:.3f} s = "
              f"{n.value / max(s.value, 1e-9) / 1e6:.2f} M records/s")


if __name__ == "__main__":
    main()
