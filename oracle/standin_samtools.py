#!/usr/bin/env python3
"""Stand-in for the external `samtools` binary, for driving the UNMODIFIED reference through its
`bam=` input path (/root/reference/src/preselect.cpp:315-447) in a sandbox without samtools.
TEST INFRASTRUCTURE ONLY.

The "BAM" handed to the reference is either a real BAM file (BGZF magic; parsed by
oracle/bamlite.py, the selections below restate the documented behaviour of samtools:
`view -L` = records overlapping a BED interval, `view -f4` = unmapped flag, `view -N` = records
with one of the names, `fastq` = primary records, reverse-flagged ones reverse-complemented, /1 /2
appended on READ1 / READ2, split over -0 / -1 / -2) or a two-line text manifest: the paths of an
R1 and an R2 FASTQ.  Honoured invocations for a manifest (everything else exits 0 without
output, like /bin/true):
    samtools view -H <bam>                         -> an @SQ line for chr6
    samtools view -@ T -ML <bed> <bam>             -> one SAM-shaped line per read (QNAME used only)
    samtools view -@ T -f4 <bam>                   -> nothing (no unmapped reads)
    samtools view -@ T -h -N <names> <bam>         -> a header
    samtools fastq -0 r0 -1 r1 -2 r2 <sam>         -> R1 / R2 with /1 /2 appended to the read
                                                      names (what `samtools fastq` does), empty r0
The manifest path reaches `fastq` through the environment (HLM_STANDIN_BAM), set by the harness."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def names(path):
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                return
            f.readline(); f.readline(); f.readline()
            yield h.rstrip(b"\n").split(b" ")[0][1:]


def is_bgzf(path):
    try:
        with open(path, "rb") as f:
            return f.read(2) == b"\x1f\x8b"
    except OSError:
        return False


def real_view(a, bam):
    import bamlite

    text, refs, recs = bamlite.read_bam(bam)
    out = sys.stdout
    if "-H" in a:
        out.write(text)
        return 0
    if "-ML" in a:
        bed = bamlite.read_bed(a[a.index("-ML") + 1])
        keep = [r for r in recs if bamlite.in_regions(r, refs, bed)]
    elif "-f4" in a:
        keep = [r for r in recs if r["flag"] & 4]
    elif "-N" in a:
        names = set(open(a[a.index("-N") + 1]).read().split("\n"))
        keep = [r for r in recs if r["name"] in names]
    else:
        keep = recs
    if "-h" in a:
        out.write(text)
    for r in keep:
        out.write(bamlite.sam_line(r, refs) + "\n")
    return 0


def real_fastq(a, dst):
    import bamlite

    def records():
        for line in open(a[-1]):
            if line.startswith("@") or not line.strip():
                continue
            t = line.rstrip("\n").split("\t")
            yield t[0], int(t[1]), "" if t[9] == "*" else t[9], None if t[10] == "*" else t[10]

    sets = bamlite.fastq_sets(records())
    for key, which in (("-0", 0), ("-1", 1), ("-2", 2)):
        with open(dst[key], "w") as o:
            for h, s, q in sets[which]:
                o.write(f"{h}\n{s}\n+\n{q}\n")
    return 0


def main(a):
    if not a:
        return 0
    if a[0] == "view" and is_bgzf(a[-1]):
        return real_view(a, a[-1])
    if a[0] == "view":
        bam = a[-1]
        if "-H" in a:
            sys.stdout.write("@SQ\tSN:chr6\tLN:170805979\n")
            return 0
        if "-f4" in a:
            return 0
        if "-N" in a:
            sys.stdout.write("@SQ\tSN:chr6\tLN:170805979\n")
            return 0
        if "-ML" in a:
            r1 = open(bam).read().split("\n")[0]
            out = sys.stdout.buffer
            for n in names(r1):
                out.write(n + b"\t99\tchr6\t29942470\t60\t150M\t=\t29942600\t280\t*\t*\n")
            return 0
        return 0
    if a[0] == "fastq":
        bam = os.environ.get("HLM_STANDIN_BAM")
        dst = {a[i]: a[i + 1] for i in range(1, len(a) - 1) if a[i] in ("-0", "-1", "-2")}
        if not bam:
            return real_fastq(a, dst)
        r1, r2 = open(bam).read().split("\n")[:2]
        open(dst["-0"], "wb").close()
        for src, key, tag in ((r1, "-1", b"/1"), (r2, "-2", b"/2")):
            with open(src, "rb") as f, open(dst[key], "wb") as o:
                k = 0
                for line in f:
                    if k % 4 == 0:
                        t = line.rstrip(b"\n").split(b" ")
                        t[0] += tag
                        line = b" ".join(t) + b"\n"
                    o.write(line)
                    k += 1
        return 0
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
