"""ctypes loader for the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from hla_mapper_b200 import _abi  # noqa: E402  (struct layouts only)

_lib = None


def build(ref=False):
    """compile liborc.so (and, when the reference sources are mounted, oracle/_ref)"""
    subprocess.check_call(["make", "-s", "-C", HERE, os.path.join(HERE, "liborc.so")])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_db_open.restype = C.c_void_p
        L.orc_db_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_db_close.argtypes = [C.c_void_p]
        L.orc_db_n_loci.argtypes = [C.c_void_p]
        L.orc_db_locus_name.restype = C.c_char_p
        L.orc_db_locus_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_db_size.argtypes = [C.c_void_p]
        L.orc_db_n_alleles.restype = C.c_int64
        L.orc_db_n_alleles.argtypes = [C.c_void_p, C.c_int]
        L.orc_mtrim.argtypes = [_abi.u8p, C.c_int, C.c_int, C.c_double, C.c_int,
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_screen_read.argtypes = [C.c_void_p, _abi.u8p, C.c_int, C.c_int, C.c_int]
        L.orc_locus_mask.restype = C.c_uint32
        L.orc_locus_mask.argtypes = [C.c_void_p, _abi.u8p, C.c_int]
        L.orc_align.argtypes = [_abi.u8p, C.c_int, _abi.u8p, C.c_int, C.c_int] + \
            [C.POINTER(C.c_int)] * 4
        L.orc_set_noclip.argtypes = [C.c_int]
        L.orc_banded_ed.argtypes = [_abi.u8p, C.c_int, _abi.u8p, C.c_int, C.c_int]
        L.orc_sam_score_dna.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_sam_score_rna.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_screen_trim.argtypes = [C.c_void_p, C.POINTER(_abi.HlmParams),
                                      C.POINTER(_abi.HlmBatch), C.POINTER(_abi.HlmScreenOut),
                                      C.c_int]
        L.orc_score.argtypes = [C.c_void_p, C.POINTER(_abi.HlmParams), C.POINTER(_abi.HlmBatch),
                                C.POINTER(_abi.HlmScreenOut), C.POINTER(_abi.HlmScoreOut),
                                C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_assign.argtypes = [C.c_void_p, C.POINTER(_abi.HlmParams), C.POINTER(_abi.HlmBatch),
                                 C.POINTER(_abi.HlmScreenOut), C.POINTER(_abi.HlmScoreOut),
                                 C.POINTER(_abi.HlmAssignOut)]
        L.orc_score_alleles.argtypes = [C.c_void_p, C.POINTER(_abi.HlmParams), C.POINTER(_abi.HlmBatch),
                                        C.POINTER(_abi.HlmScreenOut), C.c_int] + [_abi.u8p] * 4 + [C.c_int]
        _lib = L
    return _lib


class OracleDB:
    def __init__(self, path, rna=0):
        err = C.create_string_buffer(512)
        self.h = lib().orc_db_open(str(path).encode(), int(rna), err, 512)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.rna = int(rna)
        self.n_loci = lib().orc_db_n_loci(self.h)
        self.loci = [lib().orc_db_locus_name(self.h, i).decode() for i in range(self.n_loci + 1)]
        self.v_size = lib().orc_db_size(self.h)

    def close(self):
        if self.h:
            lib().orc_db_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(db: OracleDB, params, rb, mode=1, nthreads=0, stages=("screen", "score", "assign"),
        size_override1=None, rna_split1=None, rna_split2=None, buffers=None, rna_blocks=None):
    """screen_trim -> score -> assign with the oracle; returns (_abi.Buffers, n_dp_cells)"""
    L = lib()
    batch = _abi.make_batch(rb, size_override1, rna_split1, rna_split2, paired=bool(params.paired),
                            rna_blocks=rna_blocks)
    buf = buffers or _abi.Buffers(rb.n_pairs, db.n_loci + 1)
    cells = C.c_int64(0)
    if "screen" in stages:
        rc = L.orc_screen_trim(db.h, C.byref(params), C.byref(batch), C.byref(buf.scr), nthreads)
        assert rc == 0
    if "score" in stages:
        rc = L.orc_score(db.h, C.byref(params), C.byref(batch), C.byref(buf.scr), C.byref(buf.sc),
                         mode, nthreads, C.byref(cells))
        if rc != 0:
            raise RuntimeError(f"orc_score rc={rc}")
    if "assign" in stages:
        rc = L.orc_assign(db.h, C.byref(params), C.byref(batch), C.byref(buf.scr), C.byref(buf.sc),
                          C.byref(buf.asg))
        assert rc == 0
    return buf, cells.value


def score_alleles(db: OracleDB, params, rb, buf, locus, nthreads=0, **kw):
    """per-allele matrix of one target set: (nm1, clip1, nm2, clip2) uint8 [n_pairs, n_alleles]"""
    L = lib()
    batch = _abi.make_batch(rb, paired=bool(params.paired), **kw)
    na = int(L.orc_db_n_alleles(db.h, int(locus)))
    arrs = [np.zeros((rb.n_pairs, na), dtype=np.uint8) for _ in range(4)]
    rc = L.orc_score_alleles(db.h, C.byref(params), C.byref(batch), C.byref(buf.scr), int(locus),
                             *[a.ctypes.data_as(_abi.u8p) for a in arrs], nthreads)
    assert rc == 0
    return tuple(arrs)


def mtrim(qual: bytes, err, v_size, slen=None):
    q = np.frombuffer(qual, dtype=np.uint8)
    lo, ln = C.c_int(0), C.c_int(0)
    ok = lib().orc_mtrim(q.ctypes.data_as(_abi.u8p), len(q), len(q) if slen is None else slen,
                         float(err), int(v_size), C.byref(lo), C.byref(ln))
    return ok, lo.value, ln.value


def align(read_codes, ref_codes, d):
    r = np.ascontiguousarray(read_codes, np.uint8)
    t = np.ascontiguousarray(ref_codes, np.uint8)
    o = [C.c_int(0) for _ in range(4)]
    lib().orc_align(r.ctypes.data_as(_abi.u8p), len(r), t.ctypes.data_as(_abi.u8p), len(t), int(d),
                    *[C.byref(x) for x in o])
    return tuple(x.value for x in o)  # S, cost, lead, trail


def banded_ed(read_codes, ref_codes, d):
    r = np.ascontiguousarray(read_codes, np.uint8)
    t = np.ascontiguousarray(ref_codes, np.uint8)
    return lib().orc_banded_ed(r.ctypes.data_as(_abi.u8p), len(r), t.ctypes.data_as(_abi.u8p),
                               len(t), int(d))


def sam_score(line: str, rna=False, clip_mode=0):
    q = C.create_string_buffer(512)
    f = lib().orc_sam_score_rna if rna else lib().orc_sam_score_dna
    v = f(line.encode(), clip_mode, q, 512)
    return v, q.value.decode()
