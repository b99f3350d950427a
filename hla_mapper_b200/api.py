"""Python binding of libhlamap's C-ABI (include/hlamap.h) via ctypes.

This is the host-side mirror of the reference's stage boundaries for the hot path
(`main_preselect` screen+trim -> per-locus sort -> scoring -> compare/assign); every call goes
through the exported C symbols, i.e. exactly what a C/C++ caller would bind.  There is NO CPU
fallback: if the CUDA library or a GPU is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import Buffers, HlmAssignOut, HlmBatch, HlmParams, HlmProfile, HlmScoreOut, HlmScreenOut

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhlamap.so")

# every symbol include/hlamap.h declares
SYMBOLS = [
    "hlm_params_default", "hlm_abi_version", "hlm_db_open", "hlm_db_close", "hlm_db_n_loci",
    "hlm_db_locus_name", "hlm_db_size", "hlm_db_stat", "hlm_ctx_create", "hlm_ctx_destroy",
    "hlm_last_error", "hlm_screen_trim", "hlm_score", "hlm_assign", "hlm_run", "hlm_batch_upload",
    "hlm_batch_free", "hlm_run_device", "hlm_fetch", "hlm_get_profile", "hlm_int32_peak",
    "hlm_map_fastq", "hlm_run_async", "hlm_wait", "hlm_multi_create", "hlm_multi_destroy",
    "hlm_multi_n_devices", "hlm_multi_last_error", "hlm_multi_run", "hlm_dp_layout_ab",
    "hlm_bam_extract", "hlm_reads_paired", "hlm_reads_count", "hlm_reads_free",
    "hlm_reads_write_fastq", "hlm_map_reads", "hlm_int32_peaks", "hlm_db_n_alleles",
    "hlm_db_allele_name", "hlm_score_alleles", "hlm_multi_map_fastq", "hlm_multi_map_reads",
    "hlm_star_sam_read", "hlm_star_free",
]

_lib = None


class HlmError(RuntimeError):
    pass


def lib():
    """load libhlamap.so (in-tree); raises if it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HlmError(f"{LIB_PATH} is missing: run `python -m hla_mapper_b200.build` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.hlm_params_default.argtypes = [C.POINTER(HlmParams), C.c_int]
    L.hlm_db_open.restype = C.c_void_p
    L.hlm_db_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.hlm_db_close.argtypes = [C.c_void_p]
    L.hlm_db_n_loci.argtypes = [C.c_void_p]
    L.hlm_db_locus_name.restype = C.c_char_p
    L.hlm_db_locus_name.argtypes = [C.c_void_p, C.c_int]
    L.hlm_db_size.argtypes = [C.c_void_p]
    L.hlm_db_stat.restype = C.c_int64
    L.hlm_db_stat.argtypes = [C.c_void_p, C.c_char_p]
    L.hlm_ctx_create.restype = C.c_void_p
    L.hlm_ctx_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(HlmParams), C.c_char_p, C.c_size_t]
    L.hlm_ctx_destroy.argtypes = [C.c_void_p]
    L.hlm_last_error.restype = C.c_char_p
    L.hlm_last_error.argtypes = [C.c_void_p]
    L.hlm_screen_trim.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut)]
    L.hlm_score.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut),
                            C.POINTER(HlmScoreOut)]
    L.hlm_assign.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut),
                             C.POINTER(HlmScoreOut), C.POINTER(HlmAssignOut)]
    L.hlm_db_n_alleles.argtypes = [C.c_void_p, C.c_int]
    L.hlm_db_allele_name.restype = C.c_char_p
    L.hlm_db_allele_name.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.hlm_score_alleles.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut), C.c_int,
                                    C.POINTER(_abi.HlmAlleleOut)]
    L.hlm_run.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut),
                          C.POINTER(HlmScoreOut), C.POINTER(HlmAssignOut)]
    L.hlm_run_async.argtypes = L.hlm_run.argtypes
    L.hlm_wait.argtypes = [C.c_void_p]
    L.hlm_multi_create.restype = C.c_void_p
    L.hlm_multi_create.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(HlmParams),
                                   C.c_char_p, C.c_size_t]
    L.hlm_multi_destroy.argtypes = [C.c_void_p]
    L.hlm_multi_n_devices.argtypes = [C.c_void_p]
    L.hlm_multi_last_error.restype = C.c_char_p
    L.hlm_multi_last_error.argtypes = [C.c_void_p]
    L.hlm_multi_run.argtypes = [C.c_void_p, C.POINTER(HlmBatch), C.POINTER(HlmScreenOut),
                                C.POINTER(HlmScoreOut), C.POINTER(HlmAssignOut)]
    L.hlm_dp_layout_ab.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.hlm_batch_upload.restype = C.c_void_p
    L.hlm_batch_upload.argtypes = [C.c_void_p, C.POINTER(HlmBatch)]
    L.hlm_batch_free.argtypes = [C.c_void_p, C.c_void_p]
    L.hlm_run_device.argtypes = [C.c_void_p, C.c_void_p]
    L.hlm_fetch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(HlmScreenOut), C.POINTER(HlmScoreOut),
                            C.POINTER(HlmAssignOut)]
    L.hlm_get_profile.argtypes = [C.c_void_p, C.POINTER(HlmProfile)]
    L.hlm_int32_peak.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 4
    L.hlm_int32_peaks.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
    L.hlm_map_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                C.POINTER(_abi.HlmFileOpts), C.POINTER(_abi.HlmFileStats)]
    L.hlm_multi_map_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p,
                                      C.POINTER(_abi.HlmFileOpts), C.POINTER(_abi.HlmFileStats)]
    L.hlm_multi_map_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p,
                                      C.POINTER(_abi.HlmFileOpts), C.POINTER(_abi.HlmFileStats)]
    L.hlm_star_sam_read.restype = C.c_int64
    L.hlm_star_sam_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.POINTER(_abi.HlmStarRec))]
    L.hlm_star_free.argtypes = [C.POINTER(_abi.HlmStarRec)]
    L.hlm_bam_extract.restype = C.c_void_p
    L.hlm_bam_extract.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                  C.POINTER(_abi.HlmBamStats), C.c_char_p, C.c_size_t]
    L.hlm_reads_paired.argtypes = [C.c_void_p]
    L.hlm_reads_count.restype = C.c_int64
    L.hlm_reads_count.argtypes = [C.c_void_p]
    L.hlm_reads_free.argtypes = [C.c_void_p]
    L.hlm_reads_write_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.hlm_map_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p,
                                C.POINTER(_abi.HlmFileOpts), C.POINTER(_abi.HlmFileStats)]
    _lib = L
    return L


class Database:
    """hlm_db: the reference's on-disk database (db_*.info, motif/, mapper/, others/)"""

    def __init__(self, path, rna=0):
        err = C.create_string_buffer(512)
        self.h = lib().hlm_db_open(str(path).encode(), int(rna), err, 512)
        if not self.h:
            raise HlmError(err.value.decode())
        self.rna = int(rna)
        self.n_loci = lib().hlm_db_n_loci(self.h)
        self.loci = [lib().hlm_db_locus_name(self.h, i).decode() for i in range(self.n_loci + 1)]
        self.v_size = lib().hlm_db_size(self.h)

    def stat(self, key):
        return int(lib().hlm_db_stat(self.h, key.encode()))

    def allele_names(self, locus):
        """record names of the locus' FASTA in file order (locus == n_loci: others)"""
        n = lib().hlm_db_n_alleles(self.h, int(locus))
        return [lib().hlm_db_allele_name(self.h, int(locus), i).decode() for i in range(n)]

    def close(self):
        if self.h:
            lib().hlm_db_close(self.h)
            self.h = None


def star_sam_blocks(sam_path, gene):
    """hlm_star_sam_read: [(qname, flag, seq_len, piece, mate, start, len)] of a STAR SAM file"""
    p = C.POINTER(_abi.HlmStarRec)()
    n = lib().hlm_star_sam_read(str(sam_path).encode(), int(gene), C.byref(p))
    if n < 0:
        raise HlmError(f"hlm_star_sam_read rc={n}")
    out = [(C.string_at(p[i].qname, p[i].qname_len).decode(), p[i].flag, p[i].seq_len, p[i].piece, p[i].mate,
            p[i].start, p[i].len) for i in range(n)]
    if n:
        lib().hlm_star_free(p)
    return out


class BamReads:
    """hlm_reads: what the reference's bam= front end (four samtools calls, preselect.cpp:315-447)
    hands to its screen loop, extracted in-process (host threads only, no GPU involved)"""

    def __init__(self, db: Database, bam, bed=None, skip_unmapped=0, threads=0):
        err = C.create_string_buffer(512)
        st = _abi.HlmBamStats()
        self.h = lib().hlm_bam_extract(db.h, str(bam).encode(), str(bed).encode() if bed else None,
                                       int(skip_unmapped), int(threads), C.byref(st), err, 512)
        if not self.h:
            raise HlmError(err.value.decode())
        self.stats = {k: getattr(st, k) for k, _ in _abi.HlmBamStats._fields_}
        self.paired = bool(lib().hlm_reads_paired(self.h))
        self.n = int(lib().hlm_reads_count(self.h))

    def write_fastq(self, r1, r2=None):
        rc = lib().hlm_reads_write_fastq(self.h, str(r1).encode(), str(r2).encode() if r2 else None)
        if rc != 0:
            raise HlmError(f"hlm_reads_write_fastq rc={rc}")

    def close(self):
        if self.h:
            lib().hlm_reads_free(self.h)
            self.h = None


class Mapper:
    """hlm_ctx: one context per GPU"""

    def __init__(self, db: Database, device=0, params: HlmParams | None = None, **kw):
        self.db = db
        if params is None:
            self.params = _abi.default_params(db.rna, **kw)
        else:  # a private copy with the keyword overrides applied
            self.params = HlmParams.from_buffer_copy(params)
            for k, v in kw.items():
                setattr(self.params, k, v)
        err = C.create_string_buffer(512)
        self.h = lib().hlm_ctx_create(db.h, int(device), C.byref(self.params), err, 512)
        if not self.h:
            raise HlmError(err.value.decode())
        self.T = db.n_loci + 1

    def _check(self, rc):
        if rc != 0:
            raise HlmError(f"libhlamap rc={rc}: {lib().hlm_last_error(self.h).decode()}")

    def _batch(self, rb, **kw):
        return _abi.make_batch(rb, paired=bool(self.params.paired), **kw)

    def run(self, rb, buffers=None, **kw):
        b = self._batch(rb, **kw)
        buf = buffers or Buffers(rb.n_pairs, self.T)
        self._check(lib().hlm_run(self.h, C.byref(b), C.byref(buf.scr), C.byref(buf.sc),
                                  C.byref(buf.asg)))
        return buf

    def run_async(self, rb, buffers=None, **kw):
        """hlm_run_async; returns the buffers, valid after wait()"""
        b = self._batch(rb, **kw)
        buf = buffers or Buffers(rb.n_pairs, self.T)
        self._check(lib().hlm_run_async(self.h, C.byref(b), C.byref(buf.scr), C.byref(buf.sc),
                                        C.byref(buf.asg)))
        self._inflight = (b, buf)  # keep the arrays alive
        return buf

    def wait(self):
        rc = lib().hlm_wait(self.h)
        self._inflight = None
        self._check(rc)

    def screen_trim(self, rb, buffers=None, **kw):
        b = self._batch(rb, **kw)
        buf = buffers or Buffers(rb.n_pairs, self.T)
        self._check(lib().hlm_screen_trim(self.h, C.byref(b), C.byref(buf.scr)))
        return buf

    def score(self, rb, buf, **kw):
        b = self._batch(rb, **kw)
        self._check(lib().hlm_score(self.h, C.byref(b), C.byref(buf.scr), C.byref(buf.sc)))
        return buf

    def score_alleles(self, rb, buf, locus, **kw):
        """hlm_score_alleles: (nm1, clip1, nm2, clip2) uint8 arrays [n_pairs, n_alleles of locus]"""
        import numpy as np

        b = self._batch(rb, **kw)
        na = lib().hlm_db_n_alleles(self.db.h, int(locus))
        arrs = [np.zeros((rb.n_pairs, na), dtype=np.uint8) for _ in range(4)]
        out = _abi.HlmAlleleOut()
        out.n_alleles = na
        u8p = C.POINTER(C.c_uint8)
        out.nm1, out.clip1, out.nm2, out.clip2 = [a.ctypes.data_as(u8p) for a in arrs]
        self._check(lib().hlm_score_alleles(self.h, C.byref(b), C.byref(buf.scr), int(locus), C.byref(out)))
        return tuple(arrs)

    def assign(self, rb, buf, **kw):
        b = self._batch(rb, **kw)
        self._check(lib().hlm_assign(self.h, C.byref(b), C.byref(buf.scr), C.byref(buf.sc),
                                     C.byref(buf.asg)))
        return buf

    # device-resident path (kernel-only timing)
    def upload(self, rb, **kw):
        b = self._batch(rb, **kw)
        h = lib().hlm_batch_upload(self.h, C.byref(b))
        if not h:
            raise HlmError("hlm_batch_upload failed: " + lib().hlm_last_error(self.h).decode())
        return h

    def run_device(self, dev_batch):
        self._check(lib().hlm_run_device(self.h, dev_batch))

    def fetch(self, dev_batch, n_pairs, buffers=None):
        buf = buffers or Buffers(n_pairs, self.T)
        self._check(lib().hlm_fetch(self.h, dev_batch, C.byref(buf.scr), C.byref(buf.sc),
                                    C.byref(buf.asg)))
        return buf

    def free(self, dev_batch):
        lib().hlm_batch_free(self.h, dev_batch)

    def map_fastq(self, r1, r2, sample, out_dir, downsampling=30, batch_pairs=0, write_selected=1,
                  select_only=0, threads=0, star_cmd=None):
        """hlm_map_fastq: FASTQ(.gz) in, the reference's files around the hot path out"""
        o = _abi.HlmFileOpts(C.sizeof(_abi.HlmFileOpts), downsampling, batch_pairs, write_selected,
                             select_only, threads, str(star_cmd).encode() if star_cmd else None)
        st = _abi.HlmFileStats()
        self._check(lib().hlm_map_fastq(self.h, str(r1).encode(), str(r2).encode() if r2 else None,
                                        sample.encode(), str(out_dir).encode(), C.byref(o),
                                        C.byref(st)))
        return {k: getattr(st, k) for k, _ in _abi.HlmFileStats._fields_}

    def map_reads(self, reads: BamReads, sample, out_dir, downsampling=30, batch_pairs=0,
                  write_selected=1, select_only=0):
        """hlm_map_reads: reads extracted from a BAM in, the same files as map_fastq out"""
        o = _abi.HlmFileOpts(C.sizeof(_abi.HlmFileOpts), downsampling, batch_pairs, write_selected,
                             select_only, 0, None)
        st = _abi.HlmFileStats()
        self._check(lib().hlm_map_reads(self.h, reads.h, sample.encode(), str(out_dir).encode(),
                                        C.byref(o), C.byref(st)))
        return {k: getattr(st, k) for k, _ in _abi.HlmFileStats._fields_}

    def profile(self):
        p = HlmProfile()
        lib().hlm_get_profile(self.h, C.byref(p))
        return {k: getattr(p, k) for k, _ in HlmProfile._fields_}

    def close(self):
        if self.h:
            lib().hlm_ctx_destroy(self.h)
            self.h = None


class MultiMapper:
    """hlm_multi: every GPU of the box (or the listed ones), pairs sharded by contiguous range"""

    def __init__(self, db: Database, devices=None, params: HlmParams | None = None, **kw):
        self.db = db
        self.params = HlmParams.from_buffer_copy(params) if params is not None else _abi.default_params(db.rna)
        for k, v in kw.items():
            setattr(self.params, k, v)
        err = C.create_string_buffer(512)
        arr = (C.c_int * len(devices))(*devices) if devices else None
        self.h = lib().hlm_multi_create(db.h, arr, len(devices) if devices else 0,
                                        C.byref(self.params), err, 512)
        if not self.h:
            raise HlmError(err.value.decode())
        self.n_devices = lib().hlm_multi_n_devices(self.h)
        self.T = db.n_loci + 1

    def run(self, rb, buffers=None, **kw):
        b = _abi.make_batch(rb, paired=bool(self.params.paired), **kw)
        buf = buffers or Buffers(rb.n_pairs, self.T)
        rc = lib().hlm_multi_run(self.h, C.byref(b), C.byref(buf.scr), C.byref(buf.sc), C.byref(buf.asg))
        if rc != 0:
            raise HlmError(f"libhlamap rc={rc}: {lib().hlm_multi_last_error(self.h).decode()}")
        return buf

    def map_fastq(self, r1, r2, sample, out_dir, downsampling=30, batch_pairs=0, write_selected=1,
                  select_only=0, threads=0, star_cmd=None):
        """hlm_multi_map_fastq: the file pipeline fed to every GPU of the multi-context"""
        o = _abi.HlmFileOpts(C.sizeof(_abi.HlmFileOpts), downsampling, batch_pairs, write_selected,
                             select_only, threads, str(star_cmd).encode() if star_cmd else None)
        st = _abi.HlmFileStats()
        rc = lib().hlm_multi_map_fastq(self.h, str(r1).encode(), str(r2).encode() if r2 else None,
                                       sample.encode(), str(out_dir).encode(), C.byref(o), C.byref(st))
        if rc != 0:
            raise HlmError(f"libhlamap rc={rc}: {lib().hlm_multi_last_error(self.h).decode()}")
        return {k: getattr(st, k) for k, _ in _abi.HlmFileStats._fields_}

    def map_reads(self, reads, sample, out_dir, downsampling=30, batch_pairs=0, write_selected=1, select_only=0):
        """hlm_multi_map_reads: reads extracted from a BAM, every GPU of the multi-context"""
        o = _abi.HlmFileOpts(C.sizeof(_abi.HlmFileOpts), downsampling, batch_pairs, write_selected,
                             select_only, 0, None)
        st = _abi.HlmFileStats()
        rc = lib().hlm_multi_map_reads(self.h, reads.h, sample.encode(), str(out_dir).encode(), C.byref(o),
                                       C.byref(st))
        if rc != 0:
            raise HlmError(f"libhlamap rc={rc}: {lib().hlm_multi_last_error(self.h).decode()}")
        return {k: getattr(st, k) for k, _ in _abi.HlmFileStats._fields_}

    def close(self):
        if self.h:
            lib().hlm_multi_destroy(self.h)
            self.h = None


def dp_layout_ab(device=0, n_cands=1 << 18, read_len=150):
    """thread-per-candidate vs warp-per-candidate DP on the same synthetic candidates"""
    a, b = C.c_double(0), C.c_double(0)
    bad, cells = C.c_int64(0), C.c_int64(0)
    rc = lib().hlm_dp_layout_ab(int(device), int(n_cands), int(read_len), C.byref(a), C.byref(b),
                                C.byref(bad), C.byref(cells))
    if rc != 0:
        raise HlmError(f"hlm_dp_layout_ab rc={rc}")
    return {"ms_thread_per_candidate": a.value, "ms_warp_per_candidate": b.value,
            "mismatching_results": bad.value, "cells": cells.value,
            "gcups_thread_per_candidate": cells.value / a.value / 1e6,
            "gcups_warp_per_candidate": cells.value / b.value / 1e6}


def int32_peak(device=0):
    """measured giga lane-ops/s of IADD3, IMAD, IADD3+IMAD, VIADDMNMX, LOP3, SHF and 3 LOP3 : 1 SHF
    chains (hlm_int32_peaks)"""
    v = (C.c_double * 7)()
    rc = lib().hlm_int32_peaks(int(device), v, 7)
    if rc != 0:
        raise HlmError(f"hlm_int32_peaks rc={rc}")
    return dict(zip(("iadd3", "imad", "mix", "viaddmnmx", "lop3", "shf", "lop3_shf_mix"), list(v)))


def addressing_table(buf, ids, loci, rna=False, paired=True):
    """rows of <sample>_addressing_table.txt (map_dna.cpp:1036-1051 / map_rna.cpp:1028-1046)
    from result buffers: {read id: ([score column strings], Target string)}"""
    rows = {}
    T = len(loci)
    for i, rid in enumerate(ids):
        if not (buf.flags[i] & _abi.HLM_F_KEPT):
            continue
        cols = []
        vm = int(buf.visible_mask[i])
        for g in range(T):
            if not (vm >> g) & 1:
                cols.append("-")
            elif rna:
                cols.append(str(int(buf.s1[i, g])))
            else:
                s1 = int(buf.others_s1[i]) if g == T - 1 else int(buf.s1[i, g])
                s2 = int(buf.others_s2[i]) if g == T - 1 else int(buf.s2[i, g])
                cols.append(f"{s1},{s2}" if paired else f"{s1}")
        tm = int(buf.target_mask[i])
        tg = [loci[g] for g in range(T) if (tm >> g) & 1]
        rows[rid] = (cols, ",".join(tg) if tg else "-")
    return rows
