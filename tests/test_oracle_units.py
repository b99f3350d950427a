"""Unit-level checks of the oracle's pieces against literal Python transcriptions of the
reference lines they restate, and of the aligner definition against a brute-force DP."""
import math
import random

import numpy as np
import pytest

K = 20
BAND = 20


# ---------------------------------------------------------------- mtrim (functions.cpp:89-145)
def ref_mtrim(seq, qual, err, v_size):
    start = end = -1
    low = high = 0
    for a in range(len(qual)):
        P = math.pow(10.0, (float(ord(qual[a]) - 33) / -10.0))
        if P <= err and start == -1:
            start = end = a
            continue
        if P > err and start >= 0:
            end = a - 1
            if (end - start) > (high - low):
                low, high = start, end
            start = end = -1
            continue
    if start >= 0:
        end = len(qual)
        if (end - start) > (high - low):
            low, high = start, end
    if low >= high:
        seq = ""
    else:
        seq = seq[low:low + (high - low + 1)]
    return (low, seq) if len(seq) >= v_size else (low, "")


@pytest.mark.parametrize("seed", range(6))
def test_mtrim_matches_literal_transcription(orc, seed):
    rng = random.Random(seed)
    err = float(np.float32(0.08))
    for _ in range(400):
        n = rng.randint(0, 160)
        mode = rng.random()
        if mode < 0.3:
            qual = "".join(rng.choice("?5+'") for _ in range(n))
        elif mode < 0.6:  # long good runs with bad islands
            qual = "".join("'" if rng.random() < 0.03 else "?" for _ in range(n))
        else:
            qual = "".join(chr(rng.randint(33, 74)) for _ in range(n))
        seq = "".join(rng.choice("ACGT") for _ in range(n))
        v_size = rng.choice([1, 20, 50])
        low, want = ref_mtrim(seq, qual, err, v_size)
        ok, lo, ln = orc.mtrim(qual.encode(), err, v_size)
        assert bool(ok) == (want != "")
        if want:
            assert (lo, ln) == (low, len(want))
            assert seq[lo:lo + ln] == want


def test_mtrim_known_answers(orc):
    err = float(np.float32(0.08))
    # Q10 ('+') is bad (P = 0.1), Q11 (',') is good (P = 0.0794 <= 0.08f)
    assert orc.mtrim(b"+" * 60, err, 50)[0] == 0
    assert orc.mtrim(b"," * 60, err, 50) == (1, 0, 60)
    # tail run gets end = len (one past): a 50-base run at the end is kept at v_size 50,
    # the same run followed by one bad base scores 49 and is cut to 50 bases anyway
    assert orc.mtrim(b"+" * 10 + b"?" * 50, err, 50) == (1, 10, 50)
    assert orc.mtrim(b"+" * 10 + b"?" * 50 + b"+", err, 50) == (1, 10, 50)
    # interior single good base never wins (end - start = 0 is not > 0)
    assert orc.mtrim(b"+?+", err, 1)[0] == 0
    # first maximal run wins on ties
    ok, lo, ln = orc.mtrim(b"?" * 30 + b"+" + b"?" * 29 + b"+", err, 1)
    assert (ok, lo, ln) == (1, 0, 30)
    # empty quality string
    assert orc.mtrim(b"", err, 50)[0] == 0


# ---------------------------------------------------------------- screens (preselect.cpp)
def ref_screen_fastq(seq, motifs, minhit=3):  # preselect.cpp:712-731
    hit = 0
    c = 0
    while c < len(seq) - K:
        if seq[c:c + K] in motifs:
            hit += 1
            c += 1
        if hit == minhit:
            return True
        c += 1
        c += 1
    return False


def ref_screen_bam(seq, motifs, minhit=3):  # preselect.cpp:487-506
    hit = 0
    c = 0
    while c < len(seq) - K:
        if seq[c:c + K] in motifs:
            hit += 1
            c += 1
        if hit == minhit:
            return True
        c += 1
    return False


def ref_locus_mask(seq, loci_motifs):  # map_dna.cpp:611-627
    m = 0
    for g, mot in enumerate(loci_motifs):
        for c in range(0, max(0, len(seq) - K)):
            if seq[c:c + K] in mot:
                m |= 1 << g
                break
    return m


def test_screen_and_locus_mask_match_literal_loops(orc, small_db, small_reads):
    import ctypes as C
    import os

    from hla_mapper_b200 import _abi

    odb = orc.OracleDB(small_db.path)
    base = small_db.path
    allm = set(open(os.path.join(base, "motif/dna/all/motif_all.txt")).read().split())
    locm = [set(open(os.path.join(base, f"motif/dna/loci/{g}.txt")).read().split())
            for g in small_db.loci]
    rb = small_reads
    L = orc.lib()
    nsel = 0
    for i in range(0, 600):
        s = rb.bases1[int(rb.offs1[i]):int(rb.offs1[i + 1])]
        seq = s.tobytes().decode()
        for cut in (len(seq), 20, 21, 22, 23, 77):
            sub = np.ascontiguousarray(s[:cut])
            p = sub.ctypes.data_as(_abi.u8p)
            assert bool(L.orc_screen_read(odb.h, p, cut, 0, 3)) == ref_screen_fastq(seq[:cut], allm)
            assert bool(L.orc_screen_read(odb.h, p, cut, 1, 3)) == ref_screen_bam(seq[:cut], allm)
            assert L.orc_locus_mask(odb.h, p, cut) == ref_locus_mask(seq[:cut], locm)
        nsel += ref_screen_fastq(seq, allm)
    assert 50 < nsel < 600


# ---------------------------------------------------------------- aligner definition
def brute_align(r, t, d, noclip=False):
    """lexicographic (S, -cost, -lead) DP over the band, independent of the packed-integer
    formulation: match +1, mismatch -4, gap k -(6+k), end clip -5; cost = mismatches + gap
    bases + clipped bases; ties: max S, min cost, min lead clip, then shortest trailing clip"""
    n = len(r)
    NEG = (-10 ** 6, 0, 0)

    def add(v, ds, dc, dl=0):
        return NEG if v[0] <= -10 ** 5 else (v[0] + ds, v[1] - dc, v[2] - dl)

    def tb(j):  # reference base consumed on entering column j (1-based), N outside
        return t[j - 1] if 1 <= j <= len(t) else 4

    H = {}
    E = {}
    for k in range(2 * BAND + 1):
        H[(0, k)] = (0, 0, 0)
        E[(0, k)] = NEG
    best, best_i = NEG, 0
    for i in range(1, n + 1):
        start = NEG if noclip else (-5, -i, -i)
        f = NEG
        rowmax = NEG
        for k in range(2 * BAND + 1):
            j = i + d - BAND + k
            sub = (1, 0) if (r[i - 1] < 4 and r[i - 1] == tb(j)) else (-4, 1)
            h = add(H[(i - 1, k)], sub[0], sub[1])
            if k + 1 <= 2 * BAND:
                e = max(add(E[(i - 1, k + 1)], -1, 1), add(H[(i - 1, k + 1)], -7, 1))
            else:
                e = NEG
            if k >= 1:
                f = max(add(f, -1, 1), add(H[(i, k - 1)], -7, 1))
            else:
                f = NEG
            h = max(h, start, e, f)
            H[(i, k)] = h
            E[(i, k)] = e
            rowmax = max(rowmax, h)
        fin = rowmax if i == n else add(rowmax, -5, n - i)
        if noclip and i < n:
            continue
        if fin >= best:
            best, best_i = fin, i
    return best[0], -best[1], -best[2], n - best_i


@pytest.mark.parametrize("seed", range(8))
def test_align_matches_bruteforce_definition(orc, seed):
    rng = np.random.default_rng(seed)
    for _ in range(25):
        tlen = int(rng.integers(40, 120))
        t = rng.integers(0, 4, size=tlen, dtype=np.uint8)
        n = int(rng.integers(12, 48))
        p = int(rng.integers(-5, tlen - n + 5))
        r = np.array([t[x] if 0 <= x < tlen else rng.integers(0, 4) for x in range(p, p + n)],
                     dtype=np.uint8)
        # errors: substitutions, an indel, N, a junk end
        for _e in range(int(rng.integers(0, 4))):
            r[int(rng.integers(0, n))] = rng.integers(0, 5)
        if rng.random() < 0.4:
            q = int(rng.integers(2, n - 2))
            r = np.delete(r, q) if rng.random() < 0.5 else np.insert(r, q, rng.integers(0, 4))
        if rng.random() < 0.3:
            k = int(rng.integers(1, 8))
            r[:k] = rng.integers(0, 4, size=k)
        if rng.random() < 0.3:
            k = int(rng.integers(1, 8))
            r[-k:] = rng.integers(0, 4, size=k)
        d = p + int(rng.integers(-3, 4))
        assert orc.align(r, t, d) == brute_align(list(map(int, r)), list(map(int, t)), d)
        # HLM_CLIP_NONE: the same recurrence without the two clipping moves
        orc.lib().orc_set_noclip(1)
        try:
            got = orc.align(r, t, d)
        finally:
            orc.lib().orc_set_noclip(0)
        assert got == brute_align(list(map(int, r)), list(map(int, t)), d, noclip=True)
        assert got[2] == 0 and got[3] == 0


def test_align_known_answers(orc):
    t = np.random.default_rng(123).integers(0, 4, size=80, dtype=np.uint8)
    r = t[10:50].copy()
    assert orc.align(r, t, 10) == (40, 0, 0, 0)          # perfect
    r2 = r.copy()
    r2[20] ^= 1
    assert orc.align(r2, t, 10) == (39 - 4, 1, 0, 0)     # one mismatch
    r3 = r.copy()
    r3[0] ^= 1                                          # mismatch at the first base: -4 vs clip -5
    assert orc.align(r3, t, 10) == (39 - 4, 1, 0, 0)
    r4 = r.copy()
    r4[:3] = 3 - r4[:3]                                  # three leading mismatches: clip wins
    assert orc.align(r4, t, 10) == (37 - 5, 3, 3, 0)
    r5 = np.delete(r, 20)                                # 1-base deletion from the read
    S, cost, lead, trail = orc.align(r5, t, 10)
    assert (cost, lead, trail) == (1, 0, 0) and S == 39 - 7
    assert orc.align(np.full(30, 4, np.uint8), t, 0)[1] == 30  # all N: everything clipped/mismatched


def test_banded_distance_is_a_lower_bound(orc):
    rng = np.random.default_rng(5)
    for _ in range(300):
        t = rng.integers(0, 4, size=200, dtype=np.uint8)
        n = int(rng.integers(30, 150))
        p = int(rng.integers(0, 200 - n))
        r = t[p:p + n].copy()
        for _e in range(int(rng.integers(0, 12))):
            r[int(rng.integers(0, n))] = rng.integers(0, 4)
        d = p + int(rng.integers(-4, 5))
        assert orc.banded_ed(r, t, d) <= orc.align(r, t, d)[1]


# ---------------------------------------------------------------- SAM reducers
def test_sam_reducers_follow_read_sam_semantics(orc):
    def rec(q, rname, cigar, nm, seq="ACGT" * 25):
        return f"{q}\t0\t{rname}\t1\t37\t{cigar}\t*\t0\t0\t{seq}\t{'I' * len(seq)}\tXT:A:U\tNM:i:{nm}"

    # map_dna.cpp:44-118: NM + first clip + 1; the trailing clip is never added because
    # splitcigar() leaves an empty last token (CLIP_LEAD_ONLY = literal behaviour)
    assert orc.sam_score(rec("r/1", "A*01", "100M", 3)) == (4, "r")
    assert orc.sam_score(rec("r/2", "A*01", "5S95M", 3)) == (9, "r")
    assert orc.sam_score(rec("r", "A*01", "95M5S", 3)) == (4, "r")
    assert orc.sam_score(rec("r", "A*01", "95M5S", 3), clip_mode=1) == (9, "r")
    assert orc.sam_score(rec("r", "A*01", "4H90M6S", 0), clip_mode=1) == (11, "r")
    assert orc.sam_score(rec("r", "*", "*", 0))[0] == -1          # unmapped: skipped
    assert orc.sam_score("@SQ\tSN:x\tLN:1")[0] == -1
    # map_rna.cpp:39-101: id up to '$'; unmapped adds len(SEQ); mapped adds NM + clips, no +1
    assert orc.sam_score(rec("r$0$1", "A*01", "100M", 2), rna=True) == (2, "r")
    assert orc.sam_score(rec("r$16$2", "*", "*", 0, seq="A" * 37), rna=True) == (37, "r")
    assert orc.sam_score(rec("r$0$1", "A*01", "7S93M", 1), rna=True) == (8, "r")
