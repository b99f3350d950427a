"""hlm_batch.rna_blocks: the STAR pre-split in full (per gene, any number of blocks per mate, one
set per reported alignment; map_rna.cpp:740-790 -> read_sam_rna :39-101).

CPU: the oracle's block-list branch against its own split-point branch, which the reference goldens
pin (tests/golden/rna_paired): a block list that spells out the same records must give the same
sums.  GPU: libhlamap against the oracle on random block lists."""
import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth


def _db_reads(root, n=600):
    db = synth.make_db(str(root / "db"), seed=17, n_loci=5, total_alleles=80, len_range=(500, 700),
                       n_others=10, kind="rna")
    rb = synth.make_reads(db, n, read_len=100, seed=1004, on_target=0.8, frag_mu=220, frag_sd=25,
                          lowq_tail=0.2)
    return db, rb


def _blocks_of_splits(buf, n_loci, sp1, sp2):
    """the records the split-point model implies: per gene of the locus mask, per mate, the whole
    trimmed mate or its two halves"""
    off, out = [0], []
    for i in range(len(buf.flags)):
        if buf.flags[i] & 2:
            for g in range(n_loci):
                if not (int(buf.locus_mask[i]) >> g) & 1:
                    continue
                for m, sp, ln in ((0, sp1, buf.trim_len1), (1, sp2, buf.trim_len2)):
                    n, s = int(ln[i]), int(sp[i])
                    if 0 < s < n:
                        out += [(g, m, 0, s, 0), (g, m, s, n - s, 0)]
                    else:
                        out.append((g, m, 0, n, 0))
        off.append(len(out))
    return np.array(off, np.uint64), np.array(out, dtype=_abi.RNA_BLOCK)


def random_blocks(rng, buf, n_loci):
    """anything a STAR SAM could turn into: genes outside the locus mask, several alignments of a
    mate, blocks that overlap / run past the end / are empty / shorter than a seed, invalid mates"""
    off, out = [0], []
    for i in range(len(buf.flags)):
        for _ in range(int(rng.integers(0, 9))):
            m = int(rng.integers(0, 2)) if rng.random() < 0.97 else 2
            n = int((buf.trim_len1, buf.trim_len2)[min(m, 1)][i]) or 100
            g = int(rng.integers(0, n_loci)) if rng.random() < 0.97 else n_loci + int(rng.integers(0, 3))
            st = int(rng.integers(0, n + 5))
            ln = int(rng.choice([0, 5, 11, 12, 13, n, 400, int(rng.integers(1, n + 1))]))
            out.append((g, m, st, ln, 0))
        off.append(len(out))
    return np.array(off, np.uint64), np.array(out, dtype=_abi.RNA_BLOCK)


def _same(a, b):
    for k in ("flags", "trim_lo1", "trim_len1", "trim_lo2", "trim_len2", "locus_mask", "s1", "target_mask",
              "visible_mask", "min_sum"):
        x, y = getattr(a, k), getattr(b, k)
        assert np.array_equal(x, y), (k, np.flatnonzero(x.ravel() != y.ravel())[:5])


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_block_list_equals_split_points(orc, tmp_path, mode):
    db, rb = _db_reads(tmp_path)
    p = _abi.default_params(rna=1)
    od = orc.OracleDB(db.path, rna=1)
    rng = np.random.default_rng(3)
    sp1 = np.where(rng.random(rb.n_pairs) < 0.4, rng.integers(1, 99, size=rb.n_pairs), 0).astype(np.uint16)
    sp2 = np.where(rng.random(rb.n_pairs) < 0.4, rng.integers(1, 99, size=rb.n_pairs), 0).astype(np.uint16)
    want, _ = orc.run(od, p, rb, mode=mode, rna_split1=sp1, rna_split2=sp2)
    blocks = _blocks_of_splits(want, od.n_loci, sp1, sp2)
    assert len(blocks[1]) > 500
    got, _ = orc.run(od, p, rb, mode=mode, rna_blocks=blocks)
    _same(want, got)
    assert (got.target_mask != 0).sum() > 100
    # a gene without a record has no score
    none = (np.zeros(rb.n_pairs + 1, np.uint64), np.zeros(0, _abi.RNA_BLOCK))
    bare, _ = orc.run(od, p, rb, mode=mode, rna_blocks=none)
    kept = (bare.flags & 2) != 0
    assert (bare.s1[kept][:, :od.n_loci] == 0xFFFF).all() and (bare.s1[kept][:, od.n_loci] != 0xFFFF).all()


@pytest.mark.gpu
def test_gpu_block_lists(orc, tmp_path):
    from hla_mapper_b200 import api

    db, rb = _db_reads(tmp_path, 900)
    p = _abi.default_params(rna=1)
    od = orc.OracleDB(db.path, rna=1)
    gd = api.Database(db.path, rna=1)
    scr, _ = orc.run(od, p, rb, stages=("screen",))
    rng = np.random.default_rng(11)
    sp1 = np.where(rng.random(rb.n_pairs) < 0.4, rng.integers(1, 99, size=rb.n_pairs), 0).astype(np.uint16)
    sp2 = np.where(rng.random(rb.n_pairs) < 0.4, rng.integers(1, 99, size=rb.n_pairs), 0).astype(np.uint16)
    for chunk in (65536, 101, 1):
        m = api.Mapper(gd, 0, p, chunk_pairs=chunk)
        sub = rb if chunk > 1 else rb.slice(0, 60)
        n = sub.n_pairs
        # (a) the records of the split-point model
        want = m.run(sub, rna_split1=sp1[:n], rna_split2=sp2[:n])
        _same(want, m.run(sub, rna_blocks=_blocks_of_splits(want, od.n_loci, sp1[:n], sp2[:n])))
        # (b) random lists against the oracle (exhaustive definition), host and device-resident paths
        for rep in range(2):
            blocks = random_blocks(rng, scr_slice(scr, n), od.n_loci)
            ref, _ = orc.run(od, p, sub, mode=0, rna_blocks=blocks)
            got = m.run(sub, rna_blocks=blocks)
            _same(ref, got)
            h = m.upload(sub, rna_blocks=blocks)
            m.run_device(h)
            _same(ref, m.fetch(h, n))
            m.free(h)
        m.close()
    # all GPUs of the box
    mm = api.MultiMapper(gd, None, p)
    blocks = random_blocks(rng, scr, od.n_loci)
    ref, _ = orc.run(od, p, rb, mode=1, rna_blocks=blocks)
    _same(ref, mm.run(rb, rna_blocks=blocks))
    mm.close()
    # malformed lists are refused
    m = api.Mapper(gd, 0, p)
    bad = (np.array([0] + [5] * (rb.n_pairs - 1) + [2], np.uint64), blocks[1])
    with pytest.raises(api.HlmError, match="not ascending"):
        m.run(rb, rna_blocks=bad)
    m.close()
    gd.close()


def scr_slice(buf, n):
    class S:
        pass
    s = S()
    s.flags, s.trim_len1, s.trim_len2 = buf.flags[:n], buf.trim_len1[:n], buf.trim_len2[:n]
    return s
