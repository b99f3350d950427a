"""ctypes mirror of include/hlamap.h (struct layouts only, no behaviour)."""
from __future__ import annotations

import ctypes as C

import numpy as np

HLM_F_SCREENED = 1
HLM_F_KEPT = 2
HLM_SCREEN_FASTQ = 0
HLM_SCREEN_BAM = 1
HLM_CLIP_LEAD_ONLY = 0
HLM_CLIP_BOTH = 1
HLM_MAX_READ = 184
HLM_E_TOOLONG = -6

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
i16p = C.POINTER(C.c_int16)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)


class HlmParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("rna", C.c_int32), ("paired", C.c_int32),
        ("screen_variant", C.c_int32), ("v_size", C.c_int32), ("minhit", C.c_int32),
        ("mm_max", C.c_int32), ("clip_mode", C.c_int32), ("tolerance", C.c_double),
        ("mtrim_error", C.c_double), ("chunk_pairs", C.c_int32), ("skip_screen", C.c_int32),
        ("cand_capacity", C.c_int64), ("single_slot", C.c_int32), ("score_cap", C.c_int32),
        ("no_dp_dedup", C.c_int32), ("reserved1", C.c_int32),
    ]


class HlmBatch(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int64), ("bases1", u8p), ("quals1", u8p), ("offs1", u64p),
        ("bases2", u8p), ("quals2", u8p), ("offs2", u64p), ("size_override1", i32p),
        ("rna_split1", u16p), ("rna_split2", u16p),
        ("rna_block_off", u64p), ("rna_blocks", C.c_void_p),
    ]


# hlm_rna_block as a numpy record
RNA_BLOCK = np.dtype([("gene", np.uint8), ("mate", np.uint8), ("start", np.uint16), ("len", np.uint16),
                      ("reserved", np.uint16)])


class HlmScreenOut(C.Structure):
    _fields_ = [("flags", u8p), ("trim_lo1", u16p), ("trim_len1", u16p), ("trim_lo2", u16p),
                ("trim_len2", u16p), ("locus_mask", u32p)]


class HlmStarRec(C.Structure):
    _fields_ = [("qname", C.c_void_p), ("qname_len", C.c_int32), ("flag", C.c_int32), ("seq_len", C.c_int32),
                ("piece", C.c_int32), ("gene", C.c_uint8), ("mate", C.c_uint8), ("start", C.c_uint16),
                ("len", C.c_uint16), ("reserved", C.c_uint16)]


class HlmAlleleOut(C.Structure):
    _fields_ = [("n_alleles", C.c_int32), ("reserved0", C.c_int32),
                ("nm1", C.POINTER(C.c_uint8)), ("clip1", C.POINTER(C.c_uint8)),
                ("nm2", C.POINTER(C.c_uint8)), ("clip2", C.POINTER(C.c_uint8))]


class HlmScoreOut(C.Structure):
    _fields_ = [("s1", u16p), ("s2", u16p), ("aln1", i16p), ("aln2", i16p), ("nm1", u8p),
                ("nm2", u8p), ("lead1", u8p), ("lead2", u8p), ("trail1", u8p), ("trail2", u8p)]


class HlmAssignOut(C.Structure):
    _fields_ = [("target_mask", u32p), ("visible_mask", u32p), ("min_sum", u16p),
                ("others_s1", u16p), ("others_s2", u16p)]


class HlmProfile(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ms_h2d", "ms_screen", "ms_seed", "ms_myers",
                                          "ms_expand", "ms_select", "ms_dp", "ms_assign",
                                          "ms_d2h", "ms_total")] + \
               [(k, C.c_int64) for k in ("n_pairs", "n_kept", "n_seed_candidates", "n_candidates", "n_dp",
                                         "n_dp_cells", "n_myers_cols", "launches", "h2d_bytes",
                                         "d2h_bytes", "bases_bytes", "n_overflow_splits")]


class HlmFileOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("downsampling", C.c_int32), ("batch_pairs", C.c_int32),
                ("write_selected", C.c_int32), ("select_only", C.c_int32), ("threads", C.c_int32),
                ("star_cmd", C.c_char_p)]


class HlmFileStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_pairs", "n_selected", "n_kept", "n_assigned",
                                         "read_mean_size")] + \
               [(k, C.c_double) for k in ("s_total", "s_gpu", "s_wait_input", "s_collect", "s_write", "s_sort")]


class HlmBamStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_records", "n_region", "n_unmapped", "n_names", "n_r0", "n_r1",
                                         "n_r2", "n_out", "bytes_compressed", "bytes_inflated")] + \
               [("paired", C.c_int32), ("cached", C.c_int32)] + \
               [(k, C.c_double) for k in ("s_pass1", "s_pass2", "s_total")]


def default_params(rna=0, **kw):
    """main.cpp:73-86 defaults (the float literal 0.08f widened to double included)"""
    p = HlmParams()
    p.struct_size = C.sizeof(HlmParams)
    p.rna = int(rna)
    p.paired = 1
    p.screen_variant = HLM_SCREEN_FASTQ
    p.v_size = 0
    p.minhit = 3
    p.mm_max = 20
    p.clip_mode = HLM_CLIP_LEAD_ONLY
    p.tolerance = 0.08 if rna else 0.05
    p.mtrim_error = float(np.float32(0.08))
    p.chunk_pairs = 0
    p.skip_screen = 0
    p.cand_capacity = 0
    p.single_slot = 0
    p.score_cap = 0
    p.no_dp_dedup = 0
    p.reserved1 = 0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


class Buffers:
    """numpy-backed output arrays + the ctypes structs that point into them"""

    def __init__(self, n_pairs, n_targets, diag=True):
        """diag=False: no per-alignment diagnostics (aln / nm / lead / trail): the score struct
        carries NULL for them and the library neither copies nor returns those arrays"""
        n, T = int(n_pairs), int(n_targets)
        self.n, self.T = n, T
        self.flags = np.zeros(n, np.uint8)
        self.trim_lo1 = np.zeros(n, np.uint16); self.trim_len1 = np.zeros(n, np.uint16)
        self.trim_lo2 = np.zeros(n, np.uint16); self.trim_len2 = np.zeros(n, np.uint16)
        self.locus_mask = np.zeros(n, np.uint32)
        self.s1 = np.zeros((n, T), np.uint16); self.s2 = np.zeros((n, T), np.uint16)
        self.aln1 = np.zeros((n, T), np.int16); self.aln2 = np.zeros((n, T), np.int16)
        self.nm1 = np.zeros((n, T), np.uint8); self.nm2 = np.zeros((n, T), np.uint8)
        self.lead1 = np.zeros((n, T), np.uint8); self.lead2 = np.zeros((n, T), np.uint8)
        self.trail1 = np.zeros((n, T), np.uint8); self.trail2 = np.zeros((n, T), np.uint8)
        self.target_mask = np.zeros(n, np.uint32); self.visible_mask = np.zeros(n, np.uint32)
        self.min_sum = np.zeros(n, np.uint16)
        self.others_s1 = np.zeros(n, np.uint16); self.others_s2 = np.zeros(n, np.uint16)
        self.scr = HlmScreenOut(_ptr(self.flags, u8p), _ptr(self.trim_lo1, u16p),
                                _ptr(self.trim_len1, u16p), _ptr(self.trim_lo2, u16p),
                                _ptr(self.trim_len2, u16p), _ptr(self.locus_mask, u32p))
        if diag:
            self.sc = HlmScoreOut(_ptr(self.s1, u16p), _ptr(self.s2, u16p), _ptr(self.aln1, i16p),
                                  _ptr(self.aln2, i16p), _ptr(self.nm1, u8p), _ptr(self.nm2, u8p),
                                  _ptr(self.lead1, u8p), _ptr(self.lead2, u8p),
                                  _ptr(self.trail1, u8p), _ptr(self.trail2, u8p))
        else:
            self.sc = HlmScoreOut(_ptr(self.s1, u16p), _ptr(self.s2, u16p))
        self.asg = HlmAssignOut(_ptr(self.target_mask, u32p), _ptr(self.visible_mask, u32p),
                                _ptr(self.min_sum, u16p), _ptr(self.others_s1, u16p),
                                _ptr(self.others_s2, u16p))

    FIELDS = ("flags", "trim_lo1", "trim_len1", "trim_lo2", "trim_len2", "locus_mask", "s1", "s2",
              "aln1", "aln2", "nm1", "nm2", "lead1", "lead2", "trail1", "trail2", "target_mask",
              "visible_mask", "min_sum", "others_s1", "others_s2")

    def as_dict(self):
        return {k: getattr(self, k) for k in self.FIELDS}


def make_batch(rb, size_override1=None, rna_split1=None, rna_split2=None, paired=True, rna_blocks=None):
    """hlm_batch over a synth.ReadBatch (keeps the arrays alive on the returned object);
    rna_blocks = (offsets uint64[n_pairs + 1], blocks RNA_BLOCK[])"""
    keep = []

    def c(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a

    b = HlmBatch()
    b.n_pairs = int(rb.n_pairs)
    b.bases1 = _ptr(c(rb.bases1, np.uint8), u8p)
    b.quals1 = _ptr(c(rb.quals1, np.uint8), u8p)
    b.offs1 = _ptr(c(rb.offs1, np.uint64), u64p)
    if paired:
        b.bases2 = _ptr(c(rb.bases2, np.uint8), u8p)
        b.quals2 = _ptr(c(rb.quals2, np.uint8), u8p)
        b.offs2 = _ptr(c(rb.offs2, np.uint64), u64p)
    if size_override1 is not None:
        b.size_override1 = _ptr(c(size_override1, np.int32), i32p)
    if rna_split1 is not None:
        b.rna_split1 = _ptr(c(rna_split1, np.uint16), u16p)
    if rna_split2 is not None:
        b.rna_split2 = _ptr(c(rna_split2, np.uint16), u16p)
    if rna_blocks is not None:
        off, blk = rna_blocks
        b.rna_block_off = _ptr(c(off, np.uint64), u64p)
        blk = c(blk, RNA_BLOCK)
        b.rna_blocks = blk.ctypes.data if len(blk) else None
    b._keep = keep
    return b
