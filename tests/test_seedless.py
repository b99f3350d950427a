"""Completeness of the seed-candidate aligner definition.

The product (and oracle modes 0 / 1) only align a mate on diagonals where one of its 12-mer seeds
matches an allele exactly.  Oracle mode 2 has no seeds: every diagonal of every allele whose band
can hold an alignment of cost <= mm_max gets the affine DP (oracle/hlm_oracle.c score_seedless).

Guarantee checked here: an alignment whose cost e (mismatches + gap bases + clipped bases) is
below G(n) leaves at least one seed untouched, so the seed candidates contain a diagonal whose
band holds it:
    seeds at stride 12 (dna; rna units of 96+ bases):  G = floor(n / 12)   (one cost unit kills <= 1 seed)
    seeds at stride 6 (rna units shorter than 96):     G = max(floor(n / 12), ceil(Q / 2)), Q = (n-12)//6+1
                                                       (the stride-12 seeds are a subset; one cost unit
                                                       kills <= 2 of the overlapping seeds)
Below G the seeded result must EQUAL the seedless one; at and above G the miss rate is printed."""
import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth


def guarantee(n, rna=False):
    if n < 12:
        return 0
    if n >= 96 or not rna:
        return n // 12
    q = (n - 12) // 6 + 1
    return max(n // 12, (q + 1) // 2)


def noisy_reads(db, n_pairs, read_len, seed, rates=(0.0, 0.03, 0.07, 0.12)):
    """on-target pairs with extra substitutions at 0 / 3 / 7 / 12 % (per pair) and a few indels"""
    rb = synth.make_reads(db, n_pairs, read_len=read_len, seed=seed, on_target=1.0, frag_mu=2 * read_len,
                          frag_sd=20, lowq_tail=0.0, ragged=True, indel_rate=4e-4)
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for bases, offs in ((rb.bases1, rb.offs1), (rb.bases2, rb.offs2)):
        for i in range(n_pairs):
            rate = rates[i % 4]
            a, b = int(offs[i]), int(offs[i + 1])
            hit = np.nonzero(rng.random(b - a) < rate)[0]
            bases[a + hit] = acgt[rng.integers(0, 4, size=len(hit))]
    return rb


def compare(want, got, T, label, rna=False):
    """want = seedless, got = seed candidates; scores are cost + 1 (clip_mode = both)"""
    below = above = miss_above = 0
    for s_w, s_g, lens in ((want.s1, got.s1, want.trim_len1), (want.s2, got.s2, want.trim_len2)):
        for i in range(s_w.shape[0]):
            g_n = guarantee(int(lens[i]), rna)
            for t in range(T):
                w, g = int(s_w[i, t]), int(s_g[i, t])
                if w == 0:
                    assert g == 0, "seed candidates found an alignment the seedless scan did not"
                    continue
                assert g == 0 or g >= w, "seed candidates beat the seedless scan"
                if w - 1 < g_n:
                    below += 1
                    assert g == w, (label, i, t, int(lens[i]), w, g)
                else:
                    above += 1
                    miss_above += int(g != w)
    print(f"[seedless {label}] {below} alignments below the guarantee: all equal; "
          f"{above} at or above it: {miss_above} differ ({100.0 * miss_above / max(above, 1):.1f} %)")
    assert below > 150 and above > 20
    return below, above, miss_above


@pytest.fixture(scope="module")
def tiny_db(tmp_path_factory):
    d = tmp_path_factory.mktemp("db_seedless")
    return synth.make_db(str(d / "db"), seed=21, n_loci=4, total_alleles=60, len_range=(500, 650), n_others=8)


@pytest.mark.parametrize("read_len", [150, 64])
def test_seed_candidates_equal_seedless_scan_below_the_guarantee(orc, tiny_db, read_len):
    rb = noisy_reads(tiny_db, 320, read_len, seed=77 + read_len)
    p = _abi.default_params(skip_screen=1, clip_mode=1, v_size=30)
    odb = orc.OracleDB(tiny_db.path)
    want, _ = orc.run(odb, p, rb, mode=2)
    got, _ = orc.run(odb, p, rb, mode=0)
    compare(want, got, odb.n_loci + 1, f"oracle n={read_len}")


@pytest.mark.gpu
@pytest.mark.parametrize("read_len", [150, 64])
def test_gpu_equals_seedless_scan_below_the_guarantee(orc, tiny_db, read_len):
    from hla_mapper_b200 import api

    rb = noisy_reads(tiny_db, 320, read_len, seed=77 + read_len)
    p = _abi.default_params(skip_screen=1, clip_mode=1, v_size=30)
    odb = orc.OracleDB(tiny_db.path)
    want, _ = orc.run(odb, p, rb, mode=2)
    gdb = api.Database(tiny_db.path)
    m = api.Mapper(gdb, 0, p)
    got = m.run(rb)
    m.close()
    gdb.close()
    compare(want, got, odb.n_loci + 1, f"gpu n={read_len}")


@pytest.mark.parametrize("split", [30, 45])
def test_rna_short_blocks_below_the_guarantee(orc, tmp_path, split):
    """rna: a 100-base mate cut into blocks of `split` and 100 - split bases (the STAR pre-split).
    Every block alignment below its guarantee must be found, so the summed score of the seed
    candidates equals the seedless one whenever every block is below G"""
    db = synth.make_db(str(tmp_path / "db"), seed=23, n_loci=3, total_alleles=40, len_range=(1000, 1200),
                       n_others=6, kind="rna", flank=0)
    rb = noisy_reads(db, 240, 100, seed=5 + split, rates=(0.0, 0.0, 0.005, 0.02))
    n = rb.n_pairs
    p = _abi.default_params(rna=1, skip_screen=1, clip_mode=1, v_size=30)
    s1 = np.full(n, split, dtype=np.uint16)
    s2 = np.full(n, split, dtype=np.uint16)
    odb = orc.OracleDB(db.path, rna=1)
    want, _ = orc.run(odb, p, rb, mode=2, rna_split1=s1, rna_split2=s2)
    got, _ = orc.run(odb, p, rb, mode=0, rna_split1=s1, rna_split2=s2)
    # the rna score is a SUM over blocks and mates (unmapped blocks add their length): a block
    # contributes its cost when found.  Equal sums whenever the seedless sum is small enough that
    # every block must be below its guarantee: sum < min over blocks of G
    T = odb.n_loci + 1
    below = differ_above = above = 0
    for i in range(n):
        if not want.flags[i] & 2:
            continue
        l1, l2 = int(want.trim_len1[i]), int(want.trim_len2[i])
        if l1 <= split or l2 <= split:
            continue
        # (a block shorter than 12 bases has no seed and no guarantee: G = 0)
        g_min = min(guarantee(split, True), guarantee(l1 - split, True), guarantee(l2 - split, True))
        for t in range(T):
            w, g = int(want.s1[i, t]), int(got.s1[i, t])
            assert g >= w
            if w < g_min:
                below += 1
                assert g == w, (i, t, w, g)
            else:
                above += 1
                differ_above += int(g != w)
    print(f"[seedless rna split={split}] {below} sums below the smallest block guarantee: all equal; {above} at or above: "
          f"{differ_above} differ")
    assert below > 5 and above > 20
