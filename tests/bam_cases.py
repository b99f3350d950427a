"""Synthetic BAM files for the bam= front end, and the restatement (on oracle/bamlite.py) of what
the reference's four samtools calls + header edits + id join (preselect.cpp:315-447, :469-472,
:549-603) hand to the screen loop.  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bamlite  # noqa: E402

from hla_mapper_b200 import synth  # noqa: E402

REFS = [("chr1", 248956422), ("chr6", 170805979), ("chr6_GL000250v2_alt", 4672374)]
HEADER = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in REFS)


def _mate(rb, i, m):
    b, q, o = (rb.bases1, rb.quals1, rb.offs1) if m == 1 else (rb.bases2, rb.quals2, rb.offs2)
    a, e = int(o[i]), int(o[i + 1])
    return b[a:e].tobytes().decode(), [v - 33 for v in q[a:e].tobytes()]


def _cigar(rng, n):
    r = rng.random()
    if n < 40 or r < 0.7:
        return [("M", n)]
    if r < 0.8:
        return [("S", 7), ("M", n - 7)]
    if r < 0.9:
        return [("M", 20), ("D", 5), ("M", n - 20)]
    if r < 0.95:
        return [("M", 15), ("I", 3), ("M", n - 18)]
    return [("H", 4), ("M", n - 6), ("S", 6)]


def make_bam(rb, path, seed, single=False):
    """every pair of `rb` as BAM records in one of the placements the extraction has to tell apart
    (in / outside the BED regions, mate elsewhere, unmapped, interval boundaries, secondary and
    supplementary copies, reverse strand, missing qualities, READ1/READ2-less records)"""
    rng = np.random.default_rng(seed)
    ids = rb.ids()
    (b0, e0), _, (b2, e2) = synth.SELECT_BED
    recs = []

    def add(name, flag, ref, pos, seq, qual, cigar=None, nref=-1, npos=-1):
        if flag & 0x10:  # stored on the forward strand of the reference
            seq, qual = bamlite.revcomp(seq), (None if qual is None else qual[::-1])
        if cigar is None:
            cigar = _cigar(rng, len(seq)) if ref >= 0 and not flag & 4 else []
        recs.append(dict(name=name, flag=flag, ref=ref, pos=pos, mapq=60 if ref >= 0 else 0, cigar=cigar,
                         next_ref=nref, next_pos=npos, tlen=0, seq=seq, qual=qual))

    for i in range(rb.n_pairs):
        s1, q1 = _mate(rb, i, 1)
        s2, q2 = _mate(rb, i, 2)
        name = ids[i]
        if i % 97 == 5:
            name += "/1x"  # the header edits remove the FIRST "/1": such a pair can never be joined
        if rng.random() < 0.01:
            q1 = None
        if single:
            c = rng.random()
            fl = 16 if rng.random() < 0.5 else 0
            if c < 0.6:
                add(name, fl, 1, int(rng.integers(b0, e0 - 500)), s1, q1)
            elif c < 0.8:
                add(name, fl, 0, int(rng.integers(1000, 10 ** 8)), s1, q1)
            else:
                add(name, 4, -1, -1, s1, q1)
            continue
        f1, f2 = (99, 147) if rng.random() < 0.5 else (83, 163)
        c = rng.random()
        p_in = int(rng.integers(b0, e0 - 1000))
        p_out = int(rng.integers(1000, 10 ** 8))
        if c < 0.50:
            add(name, f1, 1, p_in, s1, q1, nref=1, npos=p_in + 200)
            add(name, f2, 1, p_in + 200, s2, q2, nref=1, npos=p_in)
            if rng.random() < 0.06:  # supplementary copy of R1, hard-clipped
                add(name, (f1 | 0x800) & ~0x2, 1, p_in + 5000, s1[:60], (q1 or [30] * len(s1))[:60],
                    cigar=[("M", 60), ("H", len(s1) - 60)])
        elif c < 0.60:
            add(name, f1, 1, p_in, s1, q1, nref=0, npos=p_out)
            add(name, f2, 0, p_out, s2, q2, nref=1, npos=p_in)
        elif c < 0.70:
            add(name, f1, 0, p_out, s1, q1, nref=1, npos=p_in)
            add(name, f2, 1, p_in, s2, q2, nref=0, npos=p_out)
        elif c < 0.78:  # both outside; one in ten has a secondary alignment without SEQ inside
            add(name, f1, 0, p_out, s1, q1, nref=0, npos=p_out + 200)
            add(name, f2, 0, p_out + 200, s2, q2, nref=0, npos=p_out)
            if rng.random() < 0.1:
                add(name, f1 | 0x100, 1, p_in, "", [], cigar=[("M", len(s1))])
        elif c < 0.86:  # unmapped pair, at the end of the file
            add(name, 77, -1, -1, s1, q1)
            add(name, 141, -1, -1, s2, q2)
        elif c < 0.91:  # R2 unmapped, placed at its mate's position (inside or outside the regions)
            ref, pos = (1, p_in) if rng.random() < 0.5 else (0, p_out)
            add(name, 73, ref, pos, s1, q1, nref=ref, npos=pos)
            add(name, 133, ref, pos, s2, q2, cigar=[], nref=ref, npos=pos)
        elif c < 0.96:  # the edges of the third interval: [pos, end) against [b2, e2)
            n = len(s1)
            pos = [b2 - n, b2 - n + 1, e2 - 1, e2][int(rng.integers(0, 4))]
            add(name, f1, 1, pos, s1, q1, cigar=[("M", n)], nref=0, npos=p_out)
            add(name, f2, 0, p_out, s2, q2, nref=1, npos=pos)
        else:
            add(name, f1, 2, 5000 + i, s1, q1, nref=2, npos=5200 + i)
            add(name, f2, 2, 5200 + i, s2, q2, nref=2, npos=5000 + i)
    if not single:
        for k in range(max(2, rb.n_pairs // 100)):  # stragglers without READ1 / READ2 -> the -0 file
            s, q = _mate(rb, k, 1)
            add(f"SE:{k}", 16 * (k & 1), 1, b0 + 777 * k, s, q)
    key = [(r["ref"] if r["ref"] >= 0 else 1 << 30, r["pos"]) for r in recs]
    order = sorted(range(len(recs)), key=lambda j: key[j])
    bamlite.write_bam(path, HEADER, REFS, [recs[j] for j in order])
    return len(recs)


def _erase_first(s, pat):
    i = s.find(pat)
    return s if i < 0 else s[:i] + s[i + len(pat):]


def expected_reads(bam, bed, skip_unmapped=False):
    """-> (paired, [(h1, s1, q1, h2, s2, q2)]) in the order selected.R*.fastq would hold them if
    every read passed the screen; single end: h2 = s2 = q2 = None"""
    text, refs, recs = bamlite.read_bam(bam)
    regions = bamlite.read_bed(bed)
    names = {r["name"] for r in recs if bamlite.in_regions(r, refs, regions)}
    if not names:
        raise RuntimeError("Extraction failed!")
    if not skip_unmapped:
        names |= {r["name"] for r in recs if r["flag"] & 4}
    sets = bamlite.fastq_sets(
        (r["name"], r["flag"], r["seq"], None if r["qual"] is None else "".join(chr(33 + v) for v in r["qual"]))
        for r in recs if r["name"] in names)
    size = {k: sum(len(h) + 2 * len(s) + 6 for h, s, q in v) for k, v in sets.items()}
    if size[0] == 0 and size[1] == 0:
        raise RuntimeError("There is not enough data to be processed. No mapped reads retrieved from BAM.")
    paired = not size[0] > size[1]

    def keyed(recs_, pat):
        d = {}
        for h, s, q in recs_:
            h = _erase_first(h, pat)
            d[h.split(" ")[0][1:]] = (h, s, q)  # later records replace earlier ones
        return d

    if not paired:
        d = keyed(sets[0], "/1")
        return False, [d[k] + (None, None, None) for k in sorted(d, key=lambda x: x.encode())]
    d1, d2 = keyed(sets[1], "/1"), keyed(sets[2], "/2")
    return True, [d1[k] + d2[k] for k in sorted(d1, key=lambda x: x.encode()) if k in d2]
