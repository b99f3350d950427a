"""hlm_score_alleles: the per-(read, allele) NM matrix the reference's
pseudo-typing step reads from `bwa mem -a` against every allele of a gene
(src/functions.cpp:347-520).  GPU against the oracle's per-allele exhaustive definition
(oracle/hlm_oracle.c orc_score_alleles), bit for bit."""
import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth


def test_library_lists_the_alleles_in_fasta_order(small_db):
    from hla_mapper_b200 import api

    gdb = api.Database(small_db.path)
    for g, name in enumerate(small_db.loci):
        want = [l[1:].split()[0] for l in open(f"{small_db.path}/mapper/dna/{name}/{name}.fas") if l.startswith(">")]
        assert gdb.allele_names(g) == want
    others = [l[1:].split()[0] for l in open(f"{small_db.path}/others/dna/hla_gen_others.fasta") if l.startswith(">")]
    assert gdb.allele_names(gdb.n_loci) == others
    gdb.close()


def test_per_allele_minimum_equals_the_locus_score(orc, small_db, small_reads):
    """oracle self-check: the best allele's alignment is the locus' reported alignment"""
    rb = small_reads.slice(0, 120)
    p = _abi.default_params(clip_mode=1)
    odb = orc.OracleDB(small_db.path)
    buf, _ = orc.run(odb, p, rb, mode=0)
    for g in (0, 3):
        nm1, clip1, nm2, clip2 = orc.score_alleles(odb, p, rb, buf, g)
        for i in range(rb.n_pairs):
            if not buf.flags[i] & 2 or not (int(buf.locus_mask[i]) >> g) & 1:
                continue
            for nm, clip, s in ((nm1, clip1, buf.s1), (nm2, clip2, buf.s2)):
                ok = nm[i] != 255
                cost = (nm[i].astype(int) + clip[i])[ok]
                want = int(s[i, g])
                assert (cost.min() + 1 if ok.any() else 0) == want


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["dna", "dna_dense", "single", "rna"])
def test_gpu_allele_matrix_equals_oracle(orc, small_db, small_reads, tmp_path, case):
    from hla_mapper_b200 import api

    rna = int(case == "rna")
    if case == "dna":
        db, rb = small_db, small_reads.slice(0, 400)
    elif case == "dna_dense":  # many near-identical alleles: rep / child tiles, inherited results
        db = synth.make_db(str(tmp_path / "db"), seed=5, n_loci=3, total_alleles=900, len_range=(500, 600),
                           n_others=5, n_lineages=4)
        rb = synth.make_reads(db, 300, read_len=150, seed=3, on_target=1.0, frag_mu=300, frag_sd=20, ragged=True)
    elif case == "single":
        db, rb = small_db, small_reads.slice(400, 700)
    else:
        db = synth.make_db(str(tmp_path / "db"), seed=9, n_loci=4, total_alleles=150, len_range=(700, 900),
                           n_others=10, kind="rna")
        rb = synth.make_reads(db, 300, read_len=100, seed=4, on_target=1.0, frag_mu=250, frag_sd=30)
    p = _abi.default_params(rna=rna, paired=int(case != "single"))
    odb = orc.OracleDB(db.path, rna=rna)
    buf, _ = orc.run(odb, p, rb, stages=("screen",))
    gdb = api.Database(db.path, rna=rna)
    m = api.Mapper(gdb, 0, p)
    checked = 0
    for g in (0, gdb.n_loci - 1, gdb.n_loci):
        want = orc.score_alleles(odb, p, rb, buf, g)
        got = m.score_alleles(rb, buf, g)
        for k, (w, x) in enumerate(zip(want, got)):
            if case == "single" and k >= 2:
                continue
            assert w.shape == x.shape
            assert np.array_equal(w, x), (case, g, k, int((w != x).sum()))
        checked += int((want[0] != 255).sum())
    m.close()
    gdb.close()
    assert checked > 500
