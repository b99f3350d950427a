"""BASELINE.json configs 2, 3 and 5 at their full sizes on the GPU: parity with the fast oracle on a
random sample of the batch, plus size-independent properties over the whole batch (chunk
invariance, idempotence, assignment consistency, decoys rejected, reads going home)."""
import os

import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth

pytestmark = pytest.mark.gpu

DB_FULL = dict(seed=42, n_loci=24, total_alleles=40000, n_others=2000)


def _same(a, b, n=None):
    for k in a.FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        if n is not None:
            y = y[:n]
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)
            raise AssertionError(f"{k}: {len(idx)} mismatches, first at {idx[0].tolist()}")


def _sample(rb, k, seed):
    """k pairs drawn at random (without replacement, ascending) from a fixed-length batch"""
    idx = np.sort(np.random.Generator(np.random.Philox(key=[seed, 9])).choice(rb.n_pairs, size=k, replace=False))
    L = int(rb.offs1[1] - rb.offs1[0])
    pick = (idx[:, None] * L + np.arange(L)[None, :]).reshape(-1)
    offs = np.arange(k + 1, dtype=np.uint64) * np.uint64(L)
    sub = synth.ReadBatch(k, rb.bases1[pick], rb.quals1[pick], offs, rb.bases2[pick], rb.quals2[pick],
                          offs.copy(), rb.truth_locus[idx], rb.seed, 0)
    return idx, sub


def _same_at(want, got, idx):
    for k in want.FIELDS:
        x, y = getattr(want, k), getattr(got, k)[idx]
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            raise AssertionError(f"{k}: {len(bad)} mismatches, first at sample row {bad[0].tolist()}")


def _assignment_consistent(got, T, rna=False):
    """targets are exactly the visible columns holding the minimum (map_dna.cpp:990-1001)"""
    if rna:
        return
    s1 = got.s1.astype(np.int64).copy()
    s2 = got.s2.astype(np.int64).copy()
    s1[:, T - 1] = got.others_s1
    s2[:, T - 1] = got.others_s2
    vis = ((got.visible_mask[:, None] >> np.arange(T)[None, :]) & 1).astype(bool)
    tgt = ((got.target_mask[:, None] >> np.arange(T)[None, :]) & 1).astype(bool)
    sums = np.where(vis, s1 + s2, 10 ** 6)
    mn = sums.min(axis=1)
    has = vis.any(axis=1)
    assert np.array_equal(tgt, (sums == mn[:, None]) & has[:, None])
    assert np.array_equal(got.min_sum[has], mn[has].astype(np.uint16))


@pytest.fixture(scope="module")
def full_db(tmp_path_factory):
    path = os.environ.get("HLM_BENCH_DB", "/tmp/hlm_bench/db_seed42_40k")
    if not os.path.exists(os.path.join(path, ".complete")):
        synth.make_db(path, **DB_FULL)
        open(os.path.join(path, ".complete"), "w").write("ok")
    return synth.make_db(path, write=False, **DB_FULL), path


def test_config2_wgs_2x150_full_size(orc, full_db):
    """configs[1]: 2 M pairs 2x150 (substitution + indel errors, the benchmark's own batch) vs 24
    loci + others, 40k-allele database, one GPU; oracle parity on a RANDOM sample of 100 000 pairs
    of the batch (HLM_SAMPLE; about 80 s of the 16-core CPU port)"""
    from hla_mapper_b200 import api

    db, path = full_db
    n = 2_000_000
    rb = synth.make_reads_fast(db, n, read_len=150, seed=1001, on_target=1.0, indel_rate=1e-4)
    gdb = api.Database(path)
    p = _abi.default_params()
    m = api.Mapper(gdb, 0, p)
    T = gdb.n_loci + 1
    got = m.run(rb, buffers=_abi.Buffers(n, T))
    idx, sub = _sample(rb, int(os.environ.get("HLM_SAMPLE", "100000")), 1)
    want, _ = orc.run(orc.OracleDB(path), p, sub, mode=1)
    _same_at(want, got, idx)
    # idempotence + chunk invariance over the whole batch (device-resident second pass)
    m2 = api.Mapper(gdb, 0, p, chunk_pairs=50021)
    h = m2.upload(rb)
    m2.run_device(h)
    again = m2.fetch(h, n)
    m2.free(h)
    m2.close()
    _same(got, again)
    _assignment_consistent(got, T)
    kept = (got.flags & 2) != 0
    tl = rb.truth_locus
    home = ((got.target_mask >> np.maximum(tl, 0).astype(np.uint32)) & 1).astype(bool)
    assert kept.sum() > 0.9 * n and home[kept].mean() > 0.85
    m.close()
    gdb.close()


def test_config3_wes_2x100_decoys(orc, full_db):
    """configs[2] at its BASELINE size: 5 M pairs 2x100 bp, 80 % off-target decoys (exercises the
    k-mer screen and its prefilter); oracle parity on a random sample of 40 000 pairs"""
    from hla_mapper_b200 import api

    db, path = full_db
    n = 5_000_000
    rb = synth.make_reads_fast(db, n, read_len=100, seed=1002, on_target=0.2, indel_rate=1e-4, frag_mu=300,
                               frag_sd=50)
    gdb = api.Database(path)
    p = _abi.default_params()
    m = api.Mapper(gdb, 0, p)
    T = gdb.n_loci + 1
    got = m.run(rb)
    idx, sub = _sample(rb, 40000, 3)
    odb = orc.OracleDB(path)
    want, _ = orc.run(odb, p, sub, mode=1)
    _same_at(want, got, idx)
    decoy = rb.truth_locus < 0
    assert (got.flags[decoy] & 1).mean() < 0.01          # decoys do not pass the screen
    assert (got.flags[~decoy] & 1).mean() > 0.9
    assert (got.target_mask[decoy] != 0).sum() == 0
    _assignment_consistent(got, T)
    # both screen loop variants agree with the oracle (the BAM loop has no prefilter)
    pb = _abi.default_params(screen_variant=1)
    mb = api.Mapper(gdb, 0, pb)
    wb, _ = orc.run(odb, pb, sub.slice(0, 10000), mode=1)
    _same(wb, mb.run(sub.slice(0, 10000)))
    mb.close()
    m.close()
    gdb.close()


def test_config5_rna_2x100(orc, tmp_path_factory):
    """configs[4] at its BASELINE size: `rna` mode, 10 M pairs 2x100 bp vs the cDNA allele set
    (40k alleles of 900-1300 bp), tolerance 0.08, with a 5 % subset of mates pre-split into two
    blocks (the sum rule of read_sam_rna); oracle parity on a random sample of 10 000 pairs"""
    from hla_mapper_b200 import api

    kw = dict(seed=42, n_loci=24, total_alleles=40000, n_others=2000, len_range=(900, 1300), kind="rna", flank=0)
    path = os.environ.get("HLM_BENCH_DB", "/tmp/hlm_bench/db_seed42_40k") + "_rna"
    if not os.path.exists(os.path.join(path, ".complete")):
        synth.make_db(path, **kw)
        open(os.path.join(path, ".complete"), "w").write("ok")
    db = synth.make_db(path, write=False, **kw)
    n = int(os.environ.get("HLM_RNA_PAIRS", "10000000"))
    rb = synth.make_reads_fast(db, n, read_len=100, seed=1004, on_target=1.0, indel_rate=1e-4, frag_mu=250,
                               frag_sd=40)
    rng = np.random.default_rng(5)
    sp1 = np.where(rng.random(n) < 0.05, rng.integers(25, 75, size=n), 0).astype(np.uint16)
    sp2 = np.where(rng.random(n) < 0.05, rng.integers(25, 75, size=n), 0).astype(np.uint16)
    gdb = api.Database(path, rna=1)
    p = _abi.default_params(rna=1)
    m = api.Mapper(gdb, 0, p)
    got = m.run(rb, rna_split1=sp1, rna_split2=sp2)
    idx, sub = _sample(rb, 10000, 5)
    want, _ = orc.run(orc.OracleDB(path, rna=1), p, sub, mode=1, rna_split1=sp1[idx], rna_split2=sp2[idx])
    _same_at(want, got, idx)
    # chunk invariance on the first million pairs
    k = min(n, 1_000_000)
    m2 = api.Mapper(gdb, 0, p, chunk_pairs=7919)
    _same(m2.run(rb.slice(0, k), rna_split1=sp1[:k], rna_split2=sp2[:k]), got, k)
    m2.close()
    kept = (got.flags & 2) != 0
    tl = rb.truth_locus
    home = ((got.target_mask >> np.maximum(tl, 0).astype(np.uint32)) & 1).astype(bool) & (tl >= 0)
    assert home[kept & (tl >= 0)].mean() > 0.7
    m.close()
    gdb.close()
