"""Parity tests proper: libhlamap (CUDA, through the C-ABI) vs the CPU oracle on the same seeded
inputs - every output field bit-exact - plus size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

from hla_mapper_b200 import _abi, synth

pytestmark = pytest.mark.gpu


def _same(a, b, fields=None):
    for k in fields or a.FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)
            raise AssertionError(f"{k}: {len(idx)} mismatches, first at {idx[0].tolist()}: "
                                 f"oracle={x[tuple(idx[0])]} gpu={y[tuple(idx[0])]}")


@pytest.fixture(scope="module")
def gdb(small_db):
    from hla_mapper_b200 import api

    db = api.Database(small_db.path)
    yield db
    db.close()


@pytest.fixture(scope="module")
def odb(orc, small_db):
    return orc.OracleDB(small_db.path)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("clip", [0, 1, 2])
def test_fused_run_matches_exhaustive_oracle(orc, odb, gdb, small_reads, variant, clip):
    from hla_mapper_b200 import api

    p = _abi.default_params(screen_variant=variant, clip_mode=clip)
    want, _ = orc.run(odb, p, small_reads, mode=0)
    m = api.Mapper(gdb, 0, p)
    got = m.run(small_reads)
    m.close()
    _same(want, got)
    assert ((got.flags & 2) != 0).sum() > 300


def test_stagewise_and_device_resident_entry_points(orc, odb, gdb, small_reads):
    from hla_mapper_b200 import api

    p = _abi.default_params()
    want, _ = orc.run(odb, p, small_reads, mode=1)
    m = api.Mapper(gdb, 0, p)
    b = m.screen_trim(small_reads)
    m.score(small_reads, b)
    m.assign(small_reads, b)
    _same(want, b)
    h = m.upload(small_reads)
    m.run_device(h)
    _same(want, m.fetch(h, small_reads.n_pairs))
    m.run_device(h)  # idempotent
    _same(want, m.fetch(h, small_reads.n_pairs))
    m.free(h)
    assert m.profile()["launches"] > 0
    m.close()


@pytest.mark.parametrize("chunk", [1, 7, 512, 100000])
def test_chunking_does_not_change_results(orc, odb, gdb, small_db, chunk):
    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 700, read_len=120, seed=12, on_target=0.8, frag_mu=260,
                          frag_sd=30, ragged=True)
    p = _abi.default_params()
    want, _ = orc.run(odb, p, rb, mode=1)
    m = api.Mapper(gdb, 0, p, chunk_pairs=chunk)
    _same(want, m.run(rb))
    m.close()


@pytest.mark.parametrize("kw", [
    dict(tolerance=0.0), dict(tolerance=0.1), dict(mm_max=5), dict(mm_max=12, clip_mode=1),
    dict(v_size=30), dict(v_size=70), dict(minhit=1), dict(minhit=5, screen_variant=1),
    dict(mtrim_error=0.01), dict(mtrim_error=0.2), dict(skip_screen=1),
])
def test_parameter_variants(orc, odb, gdb, small_db, kw):
    """size=, tolerance=, error= and the bwa -n limit (main.cpp:73-86, :304-671) reach the kernels"""
    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 800, read_len=150, seed=31, on_target=0.8, frag_mu=300,
                          frag_sd=40, lowq_tail=0.4)
    p = _abi.default_params(**kw)
    want, _ = orc.run(odb, p, rb, mode=1)
    m = api.Mapper(gdb, 0, p)
    _same(want, m.run(rb))
    m.close()


def test_candidate_overflow_is_handled_by_splitting_chunks(orc, odb, gdb, small_reads):
    """a chunk whose candidate list does not fit the device buffer is re-run in halves"""
    from hla_mapper_b200 import api

    p = _abi.default_params()
    want, _ = orc.run(odb, p, small_reads, mode=1)
    m = api.Mapper(gdb, 0, p, chunk_pairs=600, cand_capacity=6000)
    _same(want, m.run(small_reads))
    assert m.profile()["n_overflow_splits"] > 0
    h = m.upload(small_reads)
    m.run_device(h)
    _same(want, m.fetch(h, small_reads.n_pairs))
    assert m.profile()["n_overflow_splits"] > 0
    m.free(h)
    m.close()
    # a capacity that a single pair exceeds is an error, not a wrong answer
    m = api.Mapper(gdb, 0, p, chunk_pairs=64, cand_capacity=3)
    with pytest.raises(api.HlmError, match="capacity"):
        m.run(small_reads)
    m.close()


def test_multi_context_sharding_gathers_in_order(orc, odb, gdb, small_db):
    """hlm_multi_run: contiguous pair ranges on one context per listed device (here the same
    device three times, or every GPU of the box when there are several), results gathered by
    pointer arithmetic into the caller's arrays"""
    import ctypes as C

    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 1001, read_len=120, seed=12, on_target=0.8, frag_mu=260,
                          frag_sd=30, ragged=True)
    p = _abi.default_params()
    want, _ = orc.run(odb, p, rb, mode=1)
    mm = api.MultiMapper(gdb, devices=[0, 0, 0], params=p, chunk_pairs=100)
    assert mm.n_devices == 3
    _same(want, mm.run(rb))
    mm.close()
    mm = api.MultiMapper(gdb, params=p)  # all GPUs of the box
    assert mm.n_devices >= 1
    _same(want, mm.run(rb))
    mm.close()


def test_dp_layouts_agree_bit_for_bit():
    """thread-per-candidate (product) and warp-per-candidate (shuffle wavefront) DP kernels"""
    from hla_mapper_b200 import api

    for n in (150, 100, 57, 184):
        r = api.dp_layout_ab(0, 40000, n)
        assert r["mismatching_results"] == 0, r
    r = api.dp_layout_ab(0, 1 << 18, 150)
    print(r)
    assert r["gcups_thread_per_candidate"] > r["gcups_warp_per_candidate"]


def test_async_run_overlaps_host_work(orc, odb, gdb, small_reads):
    from hla_mapper_b200 import api

    p = _abi.default_params()
    want, _ = orc.run(odb, p, small_reads, mode=1)
    m = api.Mapper(gdb, 0, p)
    buf = m.run_async(small_reads)
    with pytest.raises(api.HlmError, match="already in flight"):
        m.run_async(small_reads)
    m.wait()
    _same(want, buf)
    m.close()


def test_single_end_with_header_length_quirk(orc, odb, gdb, small_db):
    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 900, read_len=150, seed=21, on_target=0.8, frag_mu=300, frag_sd=30)
    so = np.random.default_rng(4).integers(8, 60, size=rb.n_pairs).astype(np.int32)  # preselect.cpp:1199
    p = _abi.default_params(paired=0)
    want, _ = orc.run(odb, p, rb, mode=1, size_override1=so)
    m = api.Mapper(gdb, 0, p)
    _same(want, m.run(rb, size_override1=so))
    m.close()
    assert len(set(want.target_mask.tolist())) > 2


def test_edge_cases_empty_short_long_and_n(orc, odb, gdb, small_db):
    from hla_mapper_b200 import api

    p = _abi.default_params()
    m = api.Mapper(gdb, 0, p)
    # empty batch
    z8, z64 = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    e = synth.ReadBatch(0, z8, z8, z64, z8, z8, z64, np.zeros(0, np.int32))
    assert m.run(e).flags.shape == (0,)
    # hand-made ragged batch: empty reads, < 20, < v_size, all-N, low quality, 256-base reads
    src = synth.make_reads(small_db, 64, read_len=150, seed=5, on_target=1.0, frag_mu=300, frag_sd=20)
    rows1, rows2, q1, q2 = [], [], [], []
    rng = np.random.default_rng(8)
    for i in range(64):
        b1 = src.bases1[150 * i:150 * i + 150].copy(); b2 = src.bases2[150 * i:150 * i + 150].copy()
        a1 = src.quals1[150 * i:150 * i + 150].copy(); a2 = src.quals2[150 * i:150 * i + 150].copy()
        kind = i % 8
        if kind == 0: b1, a1 = b1[:0], a1[:0]
        elif kind == 1: b1, a1 = b1[:19], a1[:19]
        elif kind == 2: b2, a2 = b2[:49], a2[:49]
        elif kind == 3: b1[:] = ord("N")
        elif kind == 4: a1[:] = ord("'")
        elif kind == 5:
            b1 = np.concatenate([b1, b1[:106]]); a1 = np.concatenate([a1, a1[:106]])  # 256 bases
        elif kind == 6: b1[rng.integers(0, 150, size=6)] = ord("n")  # lower case is not a base
        rows1.append(b1); rows2.append(b2); q1.append(a1); q2.append(a2)
    o1 = np.concatenate([[0], np.cumsum([len(x) for x in rows1])]).astype(np.uint64)
    o2 = np.concatenate([[0], np.cumsum([len(x) for x in rows2])]).astype(np.uint64)
    rb = synth.ReadBatch(64, np.concatenate(rows1), np.concatenate(q1), o1, np.concatenate(rows2),
                         np.concatenate(q2), o2, np.full(64, -1, np.int32))
    # the 256-base reads trim to > HLM_MAX_READ: both sides must refuse them the same way
    from hla_mapper_b200 import api as _api
    with pytest.raises(_api.HlmError, match="longer than HLM_MAX_READ"):
        m.run(rb)
    keep = [i for i in range(64) if i % 8 != 5]
    rows = lambda R: [R[i] for i in keep]
    o1 = np.concatenate([[0], np.cumsum([len(x) for x in rows(rows1)])]).astype(np.uint64)
    o2 = np.concatenate([[0], np.cumsum([len(x) for x in rows(rows2)])]).astype(np.uint64)
    rb = synth.ReadBatch(len(keep), np.concatenate(rows(rows1)), np.concatenate(rows(q1)), o1,
                         np.concatenate(rows(rows2)), np.concatenate(rows(q2)), o2,
                         np.full(len(keep), -1, np.int32))
    want, _ = orc.run(odb, p, rb, mode=0)
    _same(want, m.run(rb))
    m.close()


def test_rna_mode_with_split_blocks(orc, tmp_path_factory):
    from hla_mapper_b200 import api

    d = tmp_path_factory.mktemp("db_rna")
    db = synth.make_db(str(d / "db"), seed=17, n_loci=5, total_alleles=80, len_range=(500, 700),
                       n_others=10, kind="rna")
    rb = synth.make_reads(db, 800, read_len=100, seed=1004, on_target=0.8, frag_mu=220, frag_sd=25)
    rng = np.random.default_rng(2)
    sp1 = np.where(rng.random(800) < 0.3, rng.integers(20, 80, size=800), 0).astype(np.uint16)
    sp2 = np.where(rng.random(800) < 0.3, rng.integers(20, 80, size=800), 0).astype(np.uint16)
    p = _abi.default_params(rna=1)
    od = orc.OracleDB(db.path, rna=1)
    want, _ = orc.run(od, p, rb, mode=0, rna_split1=sp1, rna_split2=sp2)
    gd = api.Database(db.path, rna=1)
    m = api.Mapper(gd, 0, p)
    got = m.run(rb, rna_split1=sp1, rna_split2=sp2)
    m.close(); gd.close()
    _same(want, got)
    assert (got.target_mask != 0).sum() > 100


def test_medium_db_fast_oracle_prefix_and_properties(orc, tmp_path_factory):
    """4 000-allele database: fast oracle on a prefix; at full size the size-independent
    properties: chunk invariance, idempotence, pair-order invariance, assignment consistency"""
    from hla_mapper_b200 import api

    d = tmp_path_factory.mktemp("db_med")
    db = synth.make_db(str(d / "db"), seed=42, n_loci=24, total_alleles=4000, n_others=200)
    rb = synth.make_reads(db, 60000, read_len=150, seed=1001, on_target=0.9)
    p = _abi.default_params()
    gd = api.Database(db.path)
    m = api.Mapper(gd, 0, p)
    got = m.run(rb)
    od = orc.OracleDB(db.path)
    sub = rb.slice(0, 400)
    want, _ = orc.run(od, p, sub, mode=1)
    pre = _abi.Buffers(400, gd.n_loci + 1)
    for k in pre.FIELDS:
        getattr(pre, k)[...] = getattr(got, k)[:400]
    _same(want, pre)
    # chunk invariance + idempotence
    m2 = api.Mapper(gd, 0, p, chunk_pairs=4099)
    _same(got, m2.run(rb))
    _same(got, m2.run(rb))
    m2.close()
    # pair-order invariance: a permuted batch gives permuted results
    perm = np.random.default_rng(0).permutation(rb.n_pairs)
    L = 150
    b1 = rb.bases1.reshape(-1, L)[perm].reshape(-1); q1 = rb.quals1.reshape(-1, L)[perm].reshape(-1)
    b2 = rb.bases2.reshape(-1, L)[perm].reshape(-1); q2 = rb.quals2.reshape(-1, L)[perm].reshape(-1)
    rp = synth.ReadBatch(rb.n_pairs, b1, q1, rb.offs1, b2, q2, rb.offs2, rb.truth_locus[perm])
    gp = m.run(rp)
    for k in got.FIELDS:
        assert np.array_equal(getattr(got, k)[perm], getattr(gp, k)), k
    # assignment consistency (map_dna.cpp:932-1001): every target holds the minimum pair sum
    T = gd.n_loci + 1
    s1 = got.s1.astype(np.int64).copy(); s2 = got.s2.astype(np.int64).copy()
    s1[:, T - 1] = got.others_s1; s2[:, T - 1] = got.others_s2
    vis = ((got.visible_mask[:, None] >> np.arange(T)[None, :]) & 1).astype(bool)
    tgt = ((got.target_mask[:, None] >> np.arange(T)[None, :]) & 1).astype(bool)
    sums = np.where(vis, s1 + s2, 10 ** 6)
    mn = sums.min(axis=1)
    has = vis.any(axis=1)
    assert np.array_equal(tgt, (sums == mn[:, None]) & has[:, None])
    assert np.array_equal(got.min_sum[has], mn[has].astype(np.uint16))
    # sanity: reads go home
    tl = rb.truth_locus
    kept = (got.flags & 2) != 0
    home = ((got.target_mask >> np.maximum(tl, 0).astype(np.uint32)) & 1).astype(bool) & (tl >= 0)
    assert home.sum() > 0.8 * (kept & (tl >= 0)).sum()
    m.close(); gd.close()


def test_dp_footprint_deduplication_is_exact_and_saves_work(orc, tmp_path_factory):
    """the DP list is de-duplicated by exact DP footprint (k_dp_dedup): same results as without
    it and as the oracle, on a database with many near-identical alleles; and fewer DP alignments -
    the oracle's fast port counts algorithmic cells per distinct footprint"""
    from hla_mapper_b200 import api

    d = tmp_path_factory.mktemp("db_dedup")
    db = synth.make_db(str(d / "db"), seed=5, n_loci=3, total_alleles=900, len_range=(500, 600), n_others=5,
                       n_lineages=4)
    rb = synth.make_reads(db, 1200, read_len=150, seed=3, on_target=1.0, frag_mu=300, frag_sd=20, lowq_tail=0.3,
                          ragged=True)
    p = _abi.default_params()
    want, cells = orc.run(orc.OracleDB(db.path), p, rb, mode=1)
    gdb = api.Database(db.path)
    out = {}
    for off in (0, 1):
        m = api.Mapper(gdb, 0, p, no_dp_dedup=off)
        got = m.run(rb)
        out[off] = (m.profile()["n_dp"], m.profile()["n_dp_cells"])
        m.close()
        for k in want.FIELDS:
            assert np.array_equal(getattr(want, k), getattr(got, k)), (off, k)
    gdb.close()
    assert out[0][0] < 0.97 * out[1][0], out
    # within a few percent of the oracle's own count of cells (per distinct footprint)
    print("dp cells: oracle", cells, "gpu dedup", out[0][1], "gpu no dedup", out[1][1])
    assert out[0][1] < 1.15 * cells


def test_file_pipeline_names_a_read_it_cannot_hold(small_db, tmp_path):
    """a read longer than HLM_RAW_MAX ends hlm_map_fastq with the read named and no half-written
    selected*.fastq left behind (the reference takes any length; this build says where it stops)"""
    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 300, read_len=150, seed=5, on_target=0.9, frag_mu=300, frag_sd=30)
    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    synth.write_fastq(rb, r1, r2)
    lines = open(r1).read().split("\n")
    lines[4 * 200] = "@LONGREAD/1"
    lines[4 * 200 + 1] = "ACGT" * 75
    lines[4 * 200 + 3] = "?" * 300
    open(r1, "w").write("\n".join(lines))
    gdb = api.Database(small_db.path)
    m = api.Mapper(gdb, 0, _abi.default_params())
    out = tmp_path / "out"
    out.mkdir()
    with pytest.raises(api.HlmError, match="LONGREAD.* is 300 bases long"):
        m.map_fastq(r1, r2, "S", out, batch_pairs=64)
    assert not list(out.iterdir())
    m.close()
    gdb.close()


def test_plain_fastq_parallel_parser_edge_cases(small_db, tmp_path):
    """plain FASTQ files are mapped and parsed by several threads (batch boundaries from a newline
    count): a last line without newline completes its record, an incomplete record at the end is
    dropped, r1 / r2 of different lengths are refused before anything is written - and the files
    equal those of the single-reader path (the same input gzip-compressed)"""
    import gzip
    import hashlib

    from hla_mapper_b200 import api

    rb = synth.make_reads(small_db, 700, read_len=150, seed=8, on_target=0.8, frag_mu=300, frag_sd=30, ragged=True)
    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    synth.write_fastq(rb, r1, r2)
    gdb = api.Database(small_db.path)
    m = api.Mapper(gdb, 0, _abi.default_params())

    def digest(d):
        return {f: hashlib.sha256(open(d / f, "rb").read()).hexdigest() for f in sorted(os.listdir(d))}

    def run(a, b, name, **kw):
        out = tmp_path / name
        out.mkdir()
        st = m.map_fastq(a, b, "S", out, **kw)
        return st, digest(out)

    st0, ref = run(r1, r2, "plain", batch_pairs=64)
    assert st0["n_pairs"] == 700 and st0["n_kept"] > 300
    # the single-reader path on the same records
    z1, z2 = r1 + ".gz", r2 + ".gz"
    for src, dst in ((r1, z1), (r2, z2)):
        open(dst, "wb").write(gzip.compress(open(src, "rb").read(), 1))
    assert run(z1, z2, "gz", batch_pairs=64)[1] == ref
    assert run(r1, r2, "one_batch", batch_pairs=100000)[1] == ref
    # no newline at the end of the files; then two stray lines after the last record
    for suffix, name in ((b"", "nonl"), (b"\n@partial\nACGT", "partial")):
        a, b = str(tmp_path / (name + "1.fastq")), str(tmp_path / (name + "2.fastq"))
        for src, dst in ((r1, a), (r2, b)):
            open(dst, "wb").write(open(src, "rb").read().rstrip(b"\n") + suffix)
        st, d = run(a, b, name, batch_pairs=97)
        assert st["n_pairs"] == 700 and d == ref, name
    # one record fewer in r2
    short = str(tmp_path / "short2.fastq")
    open(short, "w").write("\n".join(open(r2).read().split("\n")[:-5]) + "\n")
    bad = tmp_path / "bad"
    bad.mkdir()
    with pytest.raises(api.HlmError, match="different numbers"):
        m.map_fastq(r1, short, "S", bad)
    assert not list(bad.iterdir())
    m.close()
    gdb.close()
