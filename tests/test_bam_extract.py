"""BAM front end (hlm_bam_extract, host code only): against the restatement of the reference's
samtools calls + header edits + id join in tests/bam_cases.py, on synthetic BAMs and on the
reference's own test BAM; the `samtools fastq` restatement itself is pinned to the FASTQ files the
reference ships next to that BAM (they are its `samtools fastq` output)."""
import os

import numpy as np
import pytest

import bam_cases
import bamlite
from hla_mapper_b200 import api, synth

# the reference's own test BAM and the `samtools fastq` files shipped next to it, committed as fixtures
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL_BAM = os.path.join(GOLDEN, "HG01890_HLA-A.bam")
DBK = dict(seed=7, n_loci=4, total_alleles=40, len_range=(500, 600), n_others=6)


@pytest.fixture(scope="module")
def dbs(tmp_path_factory):
    root = tmp_path_factory.mktemp("bamdb")
    sdb = synth.make_db(str(root / "db"), **DBK)
    return sdb, api.Database(sdb.path)


def _dump(reads, tmp):
    r1, r2 = str(tmp / "x1.fq"), str(tmp / "x2.fq")
    reads.write_fastq(r1, r2 if reads.paired else None)

    def rd(p):
        lines = open(p).read().split("\n")[:-1]
        assert all(l == "+" for l in lines[2::4])
        return list(zip(lines[0::4], lines[1::4], lines[3::4]))

    a = rd(r1)
    b = rd(r2) if reads.paired else [(None, None, None)] * len(a)
    return [x + y for x, y in zip(a, b)]


@pytest.mark.parametrize("single,block,threads,cache_gb", [(False, 0xff00, 0, None), (False, 900, 3, "0"),
                                                          (True, 4000, 1, None), (False, 0xff00, 2, "0.2")])
def test_extract_matches_the_restated_samtools_calls(dbs, tmp_path, monkeypatch, single, block, threads, cache_gb):
    sdb, db = dbs
    if cache_gb is not None:  # "0": scan twice; "0.2": one 134 MB chunk fits, a second would not
        monkeypatch.setenv("HLM_BAM_CACHE_GB", cache_gb)
    rb = synth.make_reads(sdb, 700, read_len=120, seed=5, on_target=0.8, frag_mu=260, frag_sd=30, ragged=True)
    bam = str(tmp_path / "s.bam")
    n_rec = bam_cases.make_bam(rb, bam, seed=3, single=single)
    if block != 0xff00:  # small BGZF blocks: records and the header straddle block boundaries
        text, refs, recs = bamlite.read_bam(bam)
        bamlite.write_bam(bam, text, refs, recs, block=block)
    bed = os.path.join(sdb.path, "bed", "select_dna.bed")
    for skip in (0, 1):
        paired, want = bam_cases.expected_reads(bam, bed, skip_unmapped=bool(skip))
        reads = api.BamReads(db, bam, skip_unmapped=skip, threads=threads)
        assert reads.paired == paired == (not single)
        assert reads.stats["n_records"] == n_rec
        assert reads.stats["cached"] == (0 if cache_gb == "0" else 1)
        got = _dump(reads, tmp_path)
        assert len(got) == reads.n == len(want) > 100
        assert got == want
        reads.close()


def test_unmapped_reads_are_part_of_the_selection(dbs, tmp_path):
    sdb, db = dbs
    rb = synth.make_reads(sdb, 400, read_len=100, seed=6, on_target=1.0, frag_mu=260, frag_sd=30)
    bam = str(tmp_path / "s.bam")
    bam_cases.make_bam(rb, bam, seed=4)
    with_u = api.BamReads(db, bam, skip_unmapped=0)
    without = api.BamReads(db, bam, skip_unmapped=1)
    assert with_u.stats["n_unmapped"] > 0 and without.stats["n_unmapped"] == 0
    assert with_u.n > without.n > 0


def test_reference_error_messages(dbs, tmp_path):
    sdb, db = dbs
    rb = synth.make_reads(sdb, 50, read_len=100, seed=6, on_target=1.0, frag_mu=260, frag_sd=30)
    bam = str(tmp_path / "s.bam")
    bam_cases.make_bam(rb, bam, seed=4)
    bed = str(tmp_path / "none.bed")
    open(bed, "w").write("chrZ\t1\t1000\n")
    with pytest.raises(api.HlmError, match="Extraction failed!"):  # preselect.cpp:360-363
        api.BamReads(db, bam, bed=bed)
    with pytest.raises(api.HlmError, match="cannot read"):
        api.BamReads(db, bam, bed=str(tmp_path / "missing.bed"))
    with pytest.raises(api.HlmError, match="cannot open"):
        api.BamReads(db, str(tmp_path / "missing.bam"))
    junk = str(tmp_path / "junk.bam")
    open(junk, "wb").write(b"@SYN:1\nACGT\n+\n????\n" * 10)
    with pytest.raises(api.HlmError, match="not a BGZF block"):
        api.BamReads(db, junk)
    data = open(bam, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(data[:len(data) // 2])
    with pytest.raises(api.HlmError, match="truncated|corrupt"):
        api.BamReads(db, cut)
    # only secondary records of the selected names -> nothing for `samtools fastq` to print
    text, refs, recs = bamlite.read_bam(bam)
    for r in recs:
        r["flag"] |= 0x100
    sec = str(tmp_path / "sec.bam")
    bamlite.write_bam(sec, text, refs, recs)
    with pytest.raises(api.HlmError, match="There is not enough data to be processed"):  # :431-440
        api.BamReads(db, sec)


def test_reference_test_bam(dbs, tmp_path):
    sdb, db = dbs
    text, refs, recs = bamlite.read_bam(REAL_BAM)
    assert len(recs) == 3880 and len(refs) == 3366
    # (1) the conversion: the FASTQ files next to the BAM are its `samtools fastq` output
    sets = bamlite.fastq_sets((r["name"], r["flag"], r["seq"], "".join(chr(33 + v) for v in r["qual"]))
                              for r in recs)
    assert len(sets[0]) == 0
    for which, tag in ((1, "R1"), (2, "R2")):
        import gzip

        plain = str(tmp_path / f"{tag}.fastq")
        with gzip.open(REAL_BAM.replace(".bam", f"_{tag}.fastq.gz"), "rb") as f, open(plain, "wb") as o:
            o.write(f.read())
        rb, headers = synth.read_fastq_pair(plain)
        ship = {h.split(" ")[0]: (rb.bases1[int(rb.offs1[i]):int(rb.offs1[i + 1])].tobytes().decode(),
                                  rb.quals1[int(rb.offs1[i]):int(rb.offs1[i + 1])].tobytes().decode())
                for i, h in enumerate(headers)}
        mine = {h[:-2]: (s, q) for h, s, q in sets[which]}
        assert len(mine) == 1462 and mine == ship
    # (2) the extraction, regions on chr6 around HLA-A plus one alt contig
    bed = str(tmp_path / "a.bed")
    open(bed, "w").write("chr6\t29900000\t29950000\nchr6_GL000250v2_alt\t0\t5000000\n")
    paired, want = bam_cases.expected_reads(REAL_BAM, bed)
    reads = api.BamReads(db, REAL_BAM, bed=bed)
    assert reads.paired and paired
    got = _dump(reads, tmp_path)
    assert len(want) > 1000 and got == want
