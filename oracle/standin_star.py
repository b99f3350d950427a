#!/usr/bin/env python3
"""Stand-in for the external `STAR` binary, for driving the UNMODIFIED reference through
`hla-mapper rna` in a sandbox without STAR.  TEST INFRASTRUCTURE ONLY.

Honours the one invocation on the hot path (/root/reference/src/map_rna.cpp:728): reads
--readFilesIn R1 R2, writes <--outFileNamePrefix>Aligned.out.sam with one record per mate.  A mate
is reported as spliced (CIGAR aM100NbM, which the reference then cuts into two blocks,
map_rna.cpp:748-787) by the deterministic rule `split_point()` below, shared with the tests, so
that the block structure handed to libhlamap (hlm_batch.rna_split{1,2}) is the same."""
import sys


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
    with open(prefix + "Aligned.out.sam", "wb") as out:
        out.write(b"@HD\tVN:1.4\n")
        mates = [list(records(p)) for p in files]
        for k in range(len(mates[0])):
            for m, flag in zip(mates, (b"99", b"147")):
                name, s, q = m[k]
                a = split_point(s)
                cigar = b"%dM" % len(s) if a == 0 else b"%dM100N%dM" % (a, len(s) - a)
                out.write(b"\t".join([name, flag, b"standin", b"1", b"255", cigar, b"=", b"1", b"0", s, q]) + b"\n")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
