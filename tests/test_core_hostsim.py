"""The product's kernel arithmetic (hla_mapper_b200/csrc/hlm_core.cuh) and host index builder
(hlm_db.cpp), compiled as plain C++ (tests/hostsim), against the oracle - no GPU needed.
This is not a CPU path of the product: libhlamap only ever runs these on the device."""
import ctypes as C

import numpy as np
import pytest

from hla_mapper_b200 import _abi

u8p = C.POINTER(C.c_uint8)
BAND = 20


def _p(a):
    return a.ctypes.data_as(u8p)


def _tile_from(t, d):
    """256 tile columns such that read base 0 sits at column off on diagonal d"""
    off = 20 + (d % 32)
    cols = np.full(256, 4, dtype=np.uint8)
    for c in range(256):
        x = d + (c - off)
        if 0 <= x < len(t):
            cols[c] = t[x]
    return cols, off


def _case(rng, nmax=184):
    tlen = int(rng.integers(220, 500))
    t = rng.integers(0, 4, size=tlen, dtype=np.uint8)
    if rng.random() < 0.2:
        t[rng.integers(0, tlen, size=3)] = 4
    n = int(rng.integers(20, nmax + 1))
    p = int(rng.integers(-10, tlen - n + 10))
    r = np.array([t[x] if 0 <= x < tlen else rng.integers(0, 4) for x in range(p, p + n)], dtype=np.uint8)
    ne = int(rng.integers(0, 25)) if rng.random() < 0.5 else int(rng.integers(0, 4))
    for _ in range(ne):
        r[int(rng.integers(0, n))] = rng.integers(0, 5)
    if rng.random() < 0.4 and n > 30:
        q = int(rng.integers(5, n - 5))
        k = int(rng.integers(1, 4))
        if rng.random() < 0.5:
            r = np.delete(r, np.arange(q, q + k))
        else:
            r = np.insert(r, q, rng.integers(0, 4, size=k))
        r = np.ascontiguousarray(r[:nmax])
    d = p + int(rng.integers(-6, 7))
    return np.ascontiguousarray(r), t, d


@pytest.mark.parametrize("seed", range(5))
def test_register_dp_equals_oracle_align(orc, hostsim, seed):
    rng = np.random.default_rng(seed)
    out = (C.c_int * 4)()
    for _ in range(200):
        r, t, d = _case(rng)
        cols, off = _tile_from(t, d)
        hostsim.hs_dp(_p(r), len(r), _p(cols), off, out)
        assert tuple(out) == orc.align(r, t, d)


@pytest.mark.parametrize("seed", range(5))
def test_myers_prefilter_is_a_lower_bound(orc, hostsim, seed):
    rng = np.random.default_rng(100 + seed)
    exact = 0
    for _ in range(300):
        r, t, d = _case(rng)
        cols, off = _tile_from(t, d)
        lb = hostsim.hs_myers(_p(r), len(r), _p(cols), off)
        S, cost, lead, trail = orc.align(r, t, d)
        assert 0 <= lb <= cost
        assert lb <= orc.banded_ed(r, t, d)
        exact += lb == cost
    assert exact > 30  # and it is tight often enough to be useful


def test_myers_zero_iff_exact_match(hostsim):
    rng = np.random.default_rng(9)
    for _ in range(100):
        t = rng.integers(0, 4, size=300, dtype=np.uint8)
        n = int(rng.integers(20, 185))
        p = int(rng.integers(0, 300 - n))
        r = np.ascontiguousarray(t[p:p + n])
        cols, off = _tile_from(t, p)
        assert hostsim.hs_myers(_p(r), n, _p(cols), off) == 0
        r2 = r.copy()
        r2[n // 2] = 4  # N never matches
        assert hostsim.hs_myers(_p(r2), n, _p(cols), off) == 1


def test_seed_code_and_match(hostsim):
    rng = np.random.default_rng(3)
    code = C.c_uint32(0)
    for _ in range(200):
        t = rng.integers(0, 4, size=300, dtype=np.uint8)
        n = int(rng.integers(24, 185))
        p = int(rng.integers(0, 300 - n))
        r = np.ascontiguousarray(t[p:p + n])
        cols, off = _tile_from(t, p)
        q = int(rng.integers(0, n // 12))
        rc = hostsim.hs_seed(_p(r), n, _p(cols), off, q, C.byref(code))
        assert rc == 3
        want = 0
        for x in range(12):
            want |= int(r[12 * q + x]) << (2 * x)
        assert code.value == want
        r2 = r.copy()
        r2[12 * q + 5] = (r2[12 * q + 5] + 1) & 3
        assert hostsim.hs_seed(_p(r2), n, _p(cols), off, q, C.byref(code)) == 1  # code ok, no match
        r2[12 * q + 5] = 4
        assert hostsim.hs_seed(_p(r2), n, _p(cols), off, q, C.byref(code)) == 0  # N in the seed


def test_index_builder_kmers_match_motif_files(hostsim, small_db):
    import os

    err = C.create_string_buffer(256)
    h = hostsim.hs_db_open(small_db.path.encode(), 0, err, 256)
    assert h, err.value
    assert hostsim.hs_db_stat(h, 0) == len(small_db.loci)
    assert hostsim.hs_db_stat(h, 6) == 50
    allm = set(open(os.path.join(small_db.path, "motif/dna/all/motif_all.txt")).read().split())
    locm = [set(open(os.path.join(small_db.path, f"motif/dna/loci/{g}.txt")).read().split())
            for g in small_db.loci]
    union = set(allm)
    for m in locm:
        union |= m
    assert hostsim.hs_db_stat(h, 1) == len(union)
    rng = np.random.default_rng(1)
    probe = list(union)[:3000] + ["".join("ACGT"[x] for x in rng.integers(0, 4, size=20)) for _ in range(500)]
    for k in probe:
        want = (1 << 31 if k in allm else 0)
        for g, m in enumerate(locm):
            if k in m:
                want |= 1 << g
        assert hostsim.hs_db_kmer(h, k.encode()) == want
    assert hostsim.hs_db_kmer(h, b"ACGTNACGTACGTACGTACG") == 0
    hostsim.hs_db_close(h)


def test_tile_candidates_reproduce_exhaustive_oracle(orc, hostsim, small_db, small_reads):
    """seed index -> (tile, off) candidates -> Myers bound -> DP, emulated on the host with the
    product's own arithmetic, equals the oracle's exhaustive per-allele definition"""
    err = C.create_string_buffer(256)
    h = hostsim.hs_db_open(small_db.path.encode(), 0, err, 256)
    odb = orc.OracleDB(small_db.path)
    p = _abi.default_params(skip_screen=1)
    rb = small_reads.slice(0, 60)
    want, _ = orc.run(odb, p, rb, mode=0)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    cand = (C.c_uint32 * 65536)()
    out5 = (C.c_int * 5)()
    T = odb.n_loci + 1
    checked = 0
    for i in range(rb.n_pairs):
        if not want.flags[i] & 2:
            continue
        a = int(rb.offs1[i]) + int(want.trim_lo1[i])
        raw = rb.bases1[a:a + int(want.trim_len1[i])]
        fwd = np.array([code.get(int(c), 4) for c in raw], dtype=np.uint8)
        rev = np.array([4 if c > 3 else 3 - c for c in fwd[::-1]], dtype=np.uint8)
        for g in range(T):
            if g < odb.n_loci and not (int(want.locus_mask[i]) >> g) & 1:
                continue
            best = None
            for strand in (fwd, rev):
                strand = np.ascontiguousarray(strand)
                n = hostsim.hs_db_candidates(h, _p(strand), len(strand), g, cand, 65536)
                assert n <= 65536
                for c in range(n):
                    hostsim.hs_db_eval(h, _p(strand), len(strand), cand[c], out5)
                    lb, S, cost, lead, trail = out5
                    assert lb <= cost
                    key = (cost, -S, lead, trail)
                    if best is None or key < best:
                        best = key
            got = 0
            if best is not None and best[0] <= 20:
                cost, mS, lead, trail = best
                got = (cost - lead - trail) + lead + 1
            assert got == int(want.s1[i, g]), (i, g)
            checked += 1
    assert checked > 50
    hostsim.hs_db_close(h)


def test_rep_child_hierarchy_covers_every_seed_candidate(orc, hostsim, tmp_path_factory):
    """a database with many near-identical alleles: the reduced seed index + child expansion must
    produce exactly the brute-force candidate set, minus children that are identical to their
    (present) rep inside the DP footprint"""
    from hla_mapper_b200 import synth

    d = tmp_path_factory.mktemp("db_dense")
    db = synth.make_db(str(d / "db"), seed=5, n_loci=3, total_alleles=900, len_range=(500, 600),
                       n_others=5, n_lineages=4)
    err = C.create_string_buffer(256)
    h = hostsim.hs_db_open(db.path.encode(), 0, err, 256)
    assert h, err.value
    n_tiles, n_reps, n_child = (hostsim.hs_db_stat(h, 2), hostsim.hs_db_stat(h, 8), hostsim.hs_db_stat(h, 9))
    assert n_reps + n_child == n_tiles and n_child > 2 * n_reps  # the hierarchy is real here
    rb = synth.make_reads(db, 40, read_len=150, seed=3, on_target=1.0, frag_mu=300, frag_sd=20)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    a = (C.c_uint32 * 200000)()
    b = (C.c_uint32 * 200000)()
    skipped = expanded = 0
    for i in range(rb.n_pairs):
        raw = rb.bases1[150 * i:150 * i + int(np.random.default_rng(i).integers(60, 151))]
        fwd = np.ascontiguousarray(np.array([code.get(int(c), 4) for c in raw], dtype=np.uint8))
        n = len(fwd)
        for g in range(3):
            na = hostsim.hs_db_candidates(h, _p(fwd), n, g, a, 200000)
            nb = hostsim.hs_db_candidates_brute(h, _p(fwd), n, g, b, 200000)
            A, B = list(a[:na]), set(b[:nb])
            assert len(set(A)) == len(A)          # every candidate exactly once
            assert set(A) <= B
            for c in B - set(A):
                tile, off = c >> 8, c & 0xFF
                parent = hostsim.hs_db_tile_info(h, tile, 0)
                diff = hostsim.hs_db_tile_info(h, tile, 1)
                assert parent != tile and ((parent << 8) | off) in set(A)
                cols = [(diff >> (8 * k)) & 0xFF for k in range(diff >> 24)]
                assert all(not (off - BAND <= x < off + n + BAND) for x in cols)
                skipped += 1
            expanded += sum(1 for c in A if hostsim.hs_db_tile_info(h, c >> 8, 0) != (c >> 8))
    assert skipped > 0 and expanded > 0
    hostsim.hs_db_close(h)
