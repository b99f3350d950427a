#!/usr/bin/env python3
"""Stand-in for the external `STAR` binary, for driving the UNMODIFIED reference through
`hla-mapper rna` in a sandbox without STAR.  TEST INFRASTRUCTURE ONLY.

Honours the one invocation on the hot path (/root/reference/src/map_rna.cpp:728): reads
--readFilesIn R1 R2, writes <--outFileNamePrefix>Aligned.out.sam with one record per mate.  A mate
is reported as spliced (CIGAR aM100NbM, which the reference then cuts into two blocks,
map_rna.cpp:748-787) by the deterministic rule `split_point()` below, shared with the tests, so
that the block structure handed to libhlamap (hlm_batch.rna_split{1,2}) is the same.

With HLM_STANDIN_STAR=full in the environment the stand-in uses everything the reference's SAM
loop reacts to instead (`alignments()` below): splits that depend on the gene the run is for
(--genomeDir .../loci/<gene>/), three-block alignments (whose third block the reference cuts at
start = size of the second, map_rna.cpp:784), a second (secondary) alignment of the same mate,
reverse-strand records (SEQ reverse-complemented) and unmapped records (CIGAR "*").  The tests
turn the same rule into hlm_batch.rna_blocks."""
import os
import sys
import zlib

_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def alignments(seq: bytes, gene: str):
    """[(secondary, reverse, unmapped, [block sizes])] reported for one mate in the run for `gene`"""
    n = len(seq)
    h = zlib.crc32(seq + b"|" + gene.encode())
    kind = h % 16
    if n < 40 or kind < 7:
        return [(False, False, False, [n])]
    a = 8 + (h >> 8) % (n - 16)
    if kind < 10:
        return [(False, False, False, [a, n - a])]
    if kind < 12:
        b = 4 + (h >> 16) % (n - a - 7)
        return [(False, False, False, [a, b, n - a - b])]
    if kind == 12:
        return [(False, False, False, [n]), (True, False, False, [a, n - a])]
    if kind == 13:
        return [(False, True, False, [a, n - a])]
    if kind == 14:
        return [(False, True, False, [n]), (True, True, False, [n])]
    return [(False, False, True, [n])]


def blocks(seq: bytes, gene: str):
    """the (start, length) ranges of the mate - in its own orientation - that the reference writes
    to sorted_spliced_<gene>.fastq for `alignments(seq, gene)`: block k starts where map_rna.cpp:784
    puts it (0, then the SIZE of the previous block), inside SEQ as printed"""
    n = len(seq)
    out = []
    for _sec, rev, _unm, sizes in alignments(seq, gene):
        start = 0
        for sz in sizes:
            lo, hi = min(start, n), min(start + sz, n)
            out.append((n - hi, hi - lo) if rev else (lo, hi - lo))
            start = sz
    return out


def split_point(seq: bytes) -> int:
    """0 = unspliced; otherwise the length of the first block"""
    n = len(seq)
    h = sum(seq) * 2654435761 & 0xFFFFFFFF
    if n < 60 or (h >> 8) % 4 != 0:
        return 0
    return 24 + (h >> 16) % (n - 48)


def records(path):
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                return
            s = f.readline().rstrip(b"\n")
            f.readline()
            q = f.readline().rstrip(b"\n")
            name = h.rstrip(b"\n").split(b" ")[0][1:].replace(b"/1", b"").replace(b"/2", b"")
            yield name, s, q


def main(argv):
    if "--version" in argv:
        print("2.7.10b-standin")
        return 0
    files, prefix = [], None
    i = 0
    while i < len(argv):
        if argv[i] == "--readFilesIn":
            i += 1
            while i < len(argv) and not argv[i].startswith("--"):
                files.append(argv[i])
                i += 1
            continue
        if argv[i] == "--outFileNamePrefix":
            prefix = argv[i + 1]
            i += 2
            continue
        i += 1
    if prefix is None or not files:
        return 0
    full = os.environ.get("HLM_STANDIN_STAR") == "full"
    gene = ""
    if "--genomeDir" in argv:
        gene = os.path.basename(argv[argv.index("--genomeDir") + 1].rstrip("/"))
    with open(prefix + "Aligned.out.sam", "wb") as out:
        out.write(b"@HD\tVN:1.4\n")
        mates = [list(records(p)) for p in files]
        if full:
            for k in range(len(mates[0])):
                for m, flag in zip(mates, (99, 147)):
                    name, s, q = m[k]
                    for sec, rev, unm, sizes in alignments(s, gene):
                        f = (flag ^ 0x10 if rev else flag) | (256 if sec else 0)
                        cigar = b"100N".join(b"%dM" % z for z in sizes)
                        ss, qq = (s.translate(_COMP)[::-1], q[::-1]) if rev else (s, q)
                        if unm:
                            f, cigar = (77 if flag == 99 else 141), b"*"
                        out.write(b"\t".join([name, b"%d" % f, b"*" if unm else b"standin", b"1", b"255", cigar,
                                               b"=", b"1", b"0", ss, qq]) + b"\n")
            return 0
        for k in range(len(mates[0])):
            for m, flag in zip(mates, (b"99", b"147")):
                name, s, q = m[k]
                a = split_point(s)
                cigar = b"%dM" % len(s) if a == 0 else b"%dM100N%dM" % (a, len(s) - a)
                out.write(b"\t".join([name, flag, b"standin", b"1", b"255", cigar, b"=", b"1", b"0", s, q]) + b"\n")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
