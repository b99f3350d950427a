"""Differential tests on deliberately awkward databases and reads: libhlamap (GPU) must equal the
exhaustive oracle on every field."""
import os

import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _same(a, b):
    for k in a.FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)
            raise AssertionError(f"{k}: {len(idx)} mismatches, first at {idx[0].tolist()}: "
                                 f"oracle={x[tuple(idx[0])]} gpu={y[tuple(idx[0])]}")


def _write_db(path, loci, others, rng, motif_every=3, v_size=50):
    """loci: {name: [byte strings]}; motif files = every `motif_every`-th 20-mer of every allele"""
    os.makedirs(os.path.join(path, "motif/dna/all"), exist_ok=True)
    os.makedirs(os.path.join(path, "motif/dna/loci"), exist_ok=True)
    os.makedirs(os.path.join(path, "others/dna"), exist_ok=True)
    with open(os.path.join(path, "db_dna.info"), "w") as f:
        f.write("hla-mapper:4.3+\n")
        for g in loci:
            f.write(f"GENE:{g}:chr6:1000:170805979:1:3000:HLA\n")
        f.write(f"SIZE:{v_size}\n")
    allm = set()
    for g, als in loci.items():
        os.makedirs(os.path.join(path, "mapper/dna", g), exist_ok=True)
        with open(os.path.join(path, "mapper/dna", g, g + ".fas"), "wb") as f:
            for i, s in enumerate(als):
                f.write(b">%s*%d\n" % (g.encode(), i) + s + b"\n")
        mot = set()
        for s in als:
            for c in range(0, max(0, len(s) - 19), motif_every):
                k = s[c:c + 20].upper()
                if set(k) <= set(b"ACGT"):
                    mot.add(k)
        allm |= mot
        with open(os.path.join(path, "motif/dna/loci", g + ".txt"), "wb") as f:
            f.write(b"\n".join(sorted(mot)) + b"\n")
    with open(os.path.join(path, "others/dna/hla_gen_others.fasta"), "wb") as f:
        for i, s in enumerate(others):
            f.write(b">o%d\n" % i + s + b"\n")
            for c in range(0, max(0, len(s) - 19), motif_every):
                k = s[c:c + 20].upper()
                if set(k) <= set(b"ACGT"):
                    allm.add(k)
    with open(os.path.join(path, "motif/dna/all/motif_all.txt"), "wb") as f:
        f.write(b"\n".join(sorted(allm)) + b"\n\nACGT\nNNNNNNNNNNNNNNNNNNNN\n")  # junk lines are ignored


def _rand(rng, n):
    return ACGT[rng.integers(0, 4, size=n)].tobytes()


def _mut(rng, s, rate):
    a = np.frombuffer(s, dtype=np.uint8).copy()
    m = rng.random(len(a)) < rate
    a[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
    return a.tobytes()


def _reads_from(rng, seqs, n, L, err=0.01, junk=0.15):
    b1, b2, q1, q2 = [], [], [], []
    comp = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")
    for _ in range(n):
        s = seqs[int(rng.integers(0, len(seqs)))]
        frag = int(rng.integers(L, max(L + 1, min(len(s), 3 * L))))
        if len(s) < frag:
            s = s + _rand(rng, frag - len(s))
        st = int(rng.integers(0, len(s) - frag + 1))
        f = s[st:st + frag]
        r1 = _mut(rng, f[:L], err)
        r2 = _mut(rng, f[-L:].translate(comp)[::-1], err)
        if rng.random() < junk:  # junk ends, N runs, lower case, indels
            k = int(rng.integers(1, 25))
            r1 = _rand(rng, k) + r1[k:]
        if rng.random() < junk:
            k = int(rng.integers(1, 25))
            r2 = r2[:-k] + b"N" * k
        if rng.random() < junk:
            p = int(rng.integers(5, L - 5))
            r1 = r1[:p] + r1[p + 2:] + b"AC"
        if rng.random() < 0.05:
            r1 = r1[:30] + r1[30:40].lower() + r1[40:]
        if rng.random() < 0.5:
            r1, r2 = r2, r1
        ql = rng.choice(np.frombuffer(b"?5+'", dtype=np.uint8), size=(2, L), p=[0.9, 0.05, 0.03, 0.02])
        if rng.random() < 0.2:
            ql[0, -int(rng.integers(1, L // 2)):] = ord("'")
        b1.append(r1); b2.append(r2); q1.append(ql[0].tobytes()); q2.append(ql[1].tobytes())
    o1 = np.concatenate([[0], np.cumsum([len(x) for x in b1])]).astype(np.uint64)
    o2 = np.concatenate([[0], np.cumsum([len(x) for x in b2])]).astype(np.uint64)
    f = lambda xs: np.frombuffer(b"".join(xs), dtype=np.uint8)
    return synth.ReadBatch(n, f(b1), f(q1), o1, f(b2), f(q2), o2, np.full(n, -1, np.int32))


def _check(orc, path, rb, **kw):
    from hla_mapper_b200 import api

    p = _abi.default_params(**kw)
    want, _ = orc.run(orc.OracleDB(path), p, rb, mode=0)
    gdb = api.Database(path)
    m = api.Mapper(gdb, 0, p, chunk_pairs=333)
    got = m.run(rb)
    m.close()
    gdb.close()
    _same(want, got)
    return want


def test_thirty_one_loci_fill_the_mask(orc, tmp_path):
    rng = np.random.default_rng(1)
    base = _rand(rng, 700)
    loci = {f"L{g:02d}": [_mut(rng, base, 0.01 + 0.004 * g) for _ in range(3)] for g in range(31)}
    others = [_mut(rng, base, 0.12) for _ in range(4)]
    _write_db(str(tmp_path / "db"), loci, others, rng)
    seqs = [s for als in loci.values() for s in als]
    rb = _reads_from(rng, seqs, 500, 150)
    w = _check(orc, str(tmp_path / "db"), rb)
    assert (w.locus_mask >> 30).any() and (w.target_mask != 0).sum() > 20
    assert ((w.target_mask >> 31) & 1).any() or True  # bit 31 = others when there are 31 loci


def test_short_alleles_n_runs_duplicates_and_tiny_sequences(orc, tmp_path):
    rng = np.random.default_rng(2)
    a = _rand(rng, 400)
    loci = {
        "SHORT": [a[:90], a[:120], a[50:170]],                      # alleles shorter than the reads
        "NRUN": [a[:200] + b"N" * 30 + a[230:], a[:200] + b"n" * 5 + a[205:]],
        "DUP": [a, a, a[:399] + b"A", a],                           # identical alleles
        "TINY": [b"ACGTACGTAC", a[:25], _rand(rng, 19)],            # shorter than a k-mer / a seed
        "INDEL": [a[:150] + a[153:], a[:150] + b"GATTACA" + a[150:], a],
    }
    others = [_mut(rng, a, 0.1), b"ACGT" * 60, b"A" * 300]
    _write_db(str(tmp_path / "db"), loci, others, rng, motif_every=1)
    seqs = [a, a[:200] + b"N" * 30 + a[230:], a[:150] + b"GATTACA" + a[150:], b"ACGT" * 60, b"A" * 300]
    for L, v in ((150, 50), (100, 30), (60, 20)):
        rb = _reads_from(rng, seqs, 400, L)
        w = _check(orc, str(tmp_path / "db"), rb, v_size=v)
        assert (w.flags & 2).any()
    # both clip modes and the BAM screen on the same material, then single end
    rb = _reads_from(rng, seqs, 300, 150, err=0.03, junk=0.5)
    _check(orc, str(tmp_path / "db"), rb, clip_mode=1, screen_variant=1)
    _check(orc, str(tmp_path / "db"), rb, paired=0)
    _check(orc, str(tmp_path / "db"), rb, paired=0, tolerance=0.1, mm_max=9)


def test_low_complexity_reads_and_huge_candidate_lists(orc, tmp_path):
    """repeats make every seed hit everywhere: exercises the k_seed fallback path (more index
    hits than the shared-memory set holds) and the overflow split"""
    rng = np.random.default_rng(3)
    rep = (b"ACACACACACAC" * 40)[:450]
    loci = {"REP": [_mut(rng, rep, 0.002) for _ in range(40)], "MIX": [_rand(rng, 300) + rep[:200] for _ in range(5)]}
    others = [rep[:300], _rand(rng, 400)]
    _write_db(str(tmp_path / "db"), loci, others, rng, motif_every=1)
    rb = _reads_from(rng, [rep, loci["MIX"][0]], 120, 150, err=0.005, junk=0.05)
    _check(orc, str(tmp_path / "db"), rb)


_FUZZ_FIRST = int(__import__("os").environ.get("HLM_FUZZ_FIRST", "0"))  # HLM_FUZZ_FIRST / HLM_FUZZ_SEEDS: seed window


@pytest.mark.parametrize("seed", range(_FUZZ_FIRST, _FUZZ_FIRST + int(__import__("os").environ.get("HLM_FUZZ_SEEDS", "16"))))
def test_random_databases_reads_and_parameters(orc, tmp_path, seed):
    """random database shape x read shape x parameter set, exhaustive oracle"""
    from hla_mapper_b200 import api

    rng = np.random.default_rng(1000 + seed)
    rna = int(seed % 4 == 3)
    lo = int(rng.integers(60, 400))
    db = synth.make_db(str(tmp_path / "db"), seed=int(rng.integers(1, 10 ** 6)),
                       n_loci=int(rng.integers(1, 9)), total_alleles=int(rng.integers(8, 160)),
                       len_range=(lo, lo + int(rng.integers(1, 500))), n_others=int(rng.integers(1, 8)),
                       n_lineages=int(rng.integers(1, 7)), motif_stride=int(rng.integers(1, 6)),
                       flank=int(rng.integers(0, 200)), kind="rna" if rna else "dna",
                       v_size=int(rng.choice([20, 30, 50])),
                       paralog_div=(0.001, float(rng.choice([0.01, 0.1, 0.3]))),
                       others_div=(0.001, float(rng.choice([0.02, 0.2]))))
    L = int(rng.choice([36, 50, 75, 100, 125, 150, 184]))
    rb = synth.make_reads(db, 400, read_len=L, seed=int(rng.integers(1, 10 ** 6)),
                          on_target=float(rng.choice([0.5, 0.9, 1.0])), frag_mu=max(L, int(rng.integers(L, 3 * L))),
                          frag_sd=int(rng.integers(1, 60)), lowq_tail=float(rng.choice([0.0, 0.1, 0.6])),
                          ragged=bool(rng.integers(0, 2)))
    kw = dict(rna=rna, paired=int(rng.integers(0, 2)) if not rna else 1,
              screen_variant=int(rng.integers(0, 2)), clip_mode=int(rng.integers(0, 3)),
              minhit=int(rng.integers(1, 5)), mm_max=int(rng.choice([3, 10, 20])),
              tolerance=float(rng.choice([0.0, 0.05, 0.08, 0.2])),
              mtrim_error=float(rng.choice([0.001, float(np.float32(0.08)), 0.5])),
              skip_screen=int(rng.integers(0, 4) == 0))
    p = _abi.default_params(**kw)
    extra = {}
    if rna and rng.random() < 0.5:  # the STAR pre-split as a full block list
        from test_rna_blocks import random_blocks

        class Lens:
            flags = np.zeros(400, np.uint8)
            trim_len1 = trim_len2 = np.full(400, L, np.uint16)
        extra = dict(rna_blocks=random_blocks(rng, Lens, len(db.loci)))
    elif rna:
        extra = dict(rna_split1=np.where(rng.random(400) < 0.3, rng.integers(1, L, size=400), 0).astype(np.uint16),
                     rna_split2=np.where(rng.random(400) < 0.3, rng.integers(1, L, size=400), 0).astype(np.uint16))
    elif not kw["paired"]:
        extra = dict(size_override1=rng.integers(1, 300, size=400).astype(np.int32))
    gdb = api.Database(db.path, rna=rna)
    m = api.Mapper(gdb, 0, p, chunk_pairs=int(rng.choice([1, 37, 400, 65536])))
    try:
        want, _ = orc.run(orc.OracleDB(db.path, rna=rna), p, rb, mode=0, **extra)
    except RuntimeError as e:  # a trimmed read longer than HLM_MAX_READ: both sides must refuse
        assert "rc=-6" in str(e)
        with pytest.raises(api.HlmError, match="longer than HLM_MAX_READ"):
            m.run(rb, **extra)
        return
    got = m.run(rb, **extra)
    m.close()
    gdb.close()
    _same(want, got)
