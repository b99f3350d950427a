"""Seeded golden cases shared by make_golden.py (which runs the UNMODIFIED reference on them,
in the build container) and the tests (which regenerate the same inputs from the seeds)."""
import numpy as np

from hla_mapper_b200 import _abi, synth

DB_SMALL = dict(seed=7, n_loci=6, total_alleles=120, len_range=(600, 800), n_others=20)
DB_PARALOG = dict(seed=20240813, n_loci=9, total_alleles=300, len_range=(700, 900), n_others=40)

DB_TIES = dict(seed=99, n_loci=6, total_alleles=90, len_range=(600, 800), n_others=12,
               paralog_div=(0.002, 0.01), others_div=(0.002, 0.02))

DB_RNA = dict(seed=17, n_loci=5, total_alleles=80, len_range=(500, 700), n_others=10, kind="rna")
DB_RNA_PARALOG = dict(seed=18, n_loci=6, total_alleles=90, len_range=(500, 700), n_others=10, kind="rna",
                      paralog_div=(0.002, 0.02))

CASES = {
    # name: (db kwargs, reads kwargs, mode)
    "dna_paired": (DB_SMALL, dict(n_pairs=1500, read_len=150, seed=11, on_target=0.7, frag_mu=300,
                                  frag_sd=40), "paired"),
    "dna_paired_100": (DB_PARALOG, dict(n_pairs=1500, read_len=100, seed=1002, on_target=0.5,
                                        frag_mu=250, frag_sd=30, lowq_tail=0.3), "paired"),
    "dna_ties": (DB_TIES, dict(n_pairs=1200, read_len=150, seed=77, on_target=0.9, frag_mu=300,
                               frag_sd=40, lowq_tail=0.2), "paired"),
    "dna_ragged": (DB_SMALL, dict(n_pairs=800, read_len=120, seed=12, on_target=0.8, frag_mu=260,
                                  frag_sd=30, ragged=True), "paired"),
    "rna_paired": (DB_RNA, dict(n_pairs=800, read_len=100, seed=1004, on_target=0.8, frag_mu=220,
                                 frag_sd=25), "rna"),
    # the STAR stand-in in its "full" mode: per-gene splits, three-block and secondary alignments,
    # reverse-strand and unmapped records -> hlm_batch.rna_blocks
    "rna_blocks": (DB_RNA_PARALOG, dict(n_pairs=900, read_len=100, seed=1005, on_target=0.85, frag_mu=220,
                                        frag_sd=25, lowq_tail=0.2), "rna_full"),
    # bam= input: the BAM-derived screen loop (preselect.cpp:484-506), reads served by the samtools
    # stand-in (oracle/standin_samtools.py)
    "dna_bam": (DB_SMALL, dict(n_pairs=900, read_len=120, seed=14, on_target=0.8, frag_mu=260,
                               frag_sd=30, ragged=True), "bam"),
    "dna_single": (DB_SMALL, dict(n_pairs=800, read_len=120, seed=13, on_target=0.8, frag_mu=260,
                                  frag_sd=30, ragged=True), "single"),
}

# bam= input from a real BAM file (tests/bam_cases.py writes it): the whole front end of
# preselect.cpp:315-631 - name, (db kwargs, reads kwargs, make_bam kwargs)
BAM_CASES = {
    "dna_bamfile": (DB_SMALL, dict(n_pairs=1400, read_len=120, seed=21, on_target=0.8, frag_mu=260,
                                   frag_sd=30, ragged=True), dict(seed=31, single=False)),
    "dna_bamfile_single": (DB_SMALL, dict(n_pairs=1200, read_len=120, seed=22, on_target=0.8, frag_mu=260,
                                          frag_sd=30, ragged=True), dict(seed=32, single=True)),
}


def bam_inputs(name, root):
    """(SynthDB written under root, path of the BAM, params)"""
    import os

    import bam_cases

    dbk, rk, bk = BAM_CASES[name]
    db = synth.make_db(os.path.join(root, f"db_{dbk['seed']}"), **dbk)
    rk = dict(rk)
    rb = synth.make_reads(db, rk.pop("n_pairs"), **rk)
    bam = os.path.join(root, name + ".bam")
    bam_cases.make_bam(rb, bam, **bk)
    p = _abi.default_params(rna=0, paired=0 if bk["single"] else 1, screen_variant=1)
    return db, bam, p


def inputs(name, root):
    """(SynthDB written under root, ReadBatch, params, header lines R1)"""
    import os

    dbk, rk, mode = CASES[name]
    db = synth.make_db(os.path.join(root, f"db_{dbk['seed']}"), **dbk)
    rk = dict(rk)
    n = rk.pop("n_pairs")
    rb = synth.make_reads(db, n, **rk)
    p = _abi.default_params(rna=1 if mode in ("rna", "rna_full") else 0, paired=0 if mode == "single" else 1,
                            screen_variant=1 if mode == "bam" else 0)
    return db, rb, p, mode


def rna_splits(rb, buf):
    """block split of every kept trimmed mate, by the rule the STAR stand-in applies
    (oracle/standin_star.py): what map_rna.cpp:748-787 would cut the mate into"""
    import standin_star

    out = []
    for bases, offs, lo, ln in ((rb.bases1, rb.offs1, buf.trim_lo1, buf.trim_len1),
                                (rb.bases2, rb.offs2, buf.trim_lo2, buf.trim_len2)):
        sp = np.zeros(rb.n_pairs, np.uint16)
        for i in range(rb.n_pairs):
            if buf.flags[i] & 2:
                a = int(offs[i]) + int(lo[i])
                sp[i] = standin_star.split_point(bases[a:a + int(ln[i])].tobytes())
        out.append(sp)
    return out


def rna_blocks(rb, buf, n_loci, loci):
    """hlm_batch.rna_blocks of every kept pair, by the rule the STAR stand-in applies in its full
    mode (oracle/standin_star.py): one run per gene of the pair's locus mask, both mates"""
    import standin_star

    off, out = [0], []
    for i in range(rb.n_pairs):
        if buf.flags[i] & 2:
            for g in range(n_loci):
                if not (int(buf.locus_mask[i]) >> g) & 1:
                    continue
                for m, (bases, offs, lo, ln) in enumerate(((rb.bases1, rb.offs1, buf.trim_lo1, buf.trim_len1),
                                                           (rb.bases2, rb.offs2, buf.trim_lo2, buf.trim_len2))):
                    a = int(offs[i]) + int(lo[i])
                    seq = bases[a:a + int(ln[i])].tobytes()
                    out += [(g, m, st, bl, 0) for st, bl in standin_star.blocks(seq, loci[g])]
        off.append(len(out))
    return np.array(off, np.uint64), np.array(out, dtype=_abi.RNA_BLOCK)


# how the FASTQ files of a case are written (synth.write_fastq keywords)
FASTQ_KW = {
    "dna_paired_100": dict(suffix=False, comment=True),
    "dna_ragged": dict(suffix=True, comment=True),
    "dna_ties": dict(suffix=True, gz=True),
    "dna_single": dict(suffix=False),
    "dna_bam": dict(suffix=False),
}


def write_inputs(name, rb, mode, root):
    """FASTQ files of the case -> (r1, r2 or None)"""
    import os

    kw = dict(FASTQ_KW.get(name, {}))
    ext = ".fastq.gz" if kw.get("gz") else ".fastq"
    if mode == "single":
        r0 = os.path.join(root, "r0" + ext)
        synth.write_fastq(rb, r0, None, **kw)
        return r0, None
    r1, r2 = os.path.join(root, "r1" + ext), os.path.join(root, "r2" + ext)
    synth.write_fastq(rb, r1, r2, **kw)
    return r1, r2


def header_lines(rb, mode, name=None):
    """the FASTQ header line of R1/R0 as synth.write_fastq writes it"""
    kw = FASTQ_KW.get(name, {})
    return ["@" + i + ("/1" if kw.get("suffix", True) else "") + (" 1:N:0:ACGT" if kw.get("comment") else "")
            for i in rb.ids()]


def size_override(rb, mode, name=None):
    """single-end: sequence_size_r1 = length of the id line (preselect.cpp:1199)"""
    if mode != "single":
        return None
    return np.array([len(h) for h in header_lines(rb, mode, name)], dtype=np.int32)


def raw_digest(path):
    import hashlib

    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def canonical_digest(path):
    """sha256 of a file the reference wrote, after removing the orders it leaves to thread timing:
    FASTQ -> 4-line records sorted; addressing table -> header + sorted rows"""
    import gzip
    import hashlib
    import os

    if not os.path.exists(path) and os.path.exists(path + ".gz"):  # `rna` gzips some files at the end
        data = gzip.open(path + ".gz", "rb").read()
    else:
        data = open(path, "rb").read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if path.endswith("addressing_table.txt"):
        body = [lines[0]] + sorted(lines[1:])
    else:
        recs = [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines), 4)]
        body = sorted(recs)
    return hashlib.sha256(b"\n".join(body)).hexdigest() + f":{len(lines)}"
