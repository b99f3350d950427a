#!/usr/bin/env python3
"""Stand-in for the external `samtools` binary, for driving the UNMODIFIED reference through its
`bam=` input path (/root/reference/src/preselect.cpp:315-447) in a sandbox without samtools.
TEST INFRASTRUCTURE ONLY.

The "BAM" handed to the reference is a two-line text manifest: the paths of an R1 and an R2 FASTQ.
Honoured invocations (everything else exits 0 without output, like /bin/true):
    samtools view -H <bam>                         -> an @SQ line for chr6
    samtools view -@ T -ML <bed> <bam>             -> one SAM-shaped line per read (QNAME used only)
    samtools view -@ T -f4 <bam>                   -> nothing (no unmapped reads)
    samtools view -@ T -h -N <names> <bam>         -> a header
    samtools fastq -0 r0 -1 r1 -2 r2 <sam>         -> R1 / R2 with /1 /2 appended to the read
                                                      names (what `samtools fastq` does), empty r0
The manifest path reaches `fastq` through the environment (HLM_STANDIN_BAM), set by the harness."""
import os
import sys


def names(path):
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                return
            f.readline(); f.readline(); f.readline()
            yield h.rstrip(b"\n").split(b" ")[0][1:]


def main(a):
    if not a:
        return 0
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
        if not bam:
            return 0
        r1, r2 = open(bam).read().split("\n")[:2]
        dst = {a[i]: a[i + 1] for i in range(1, len(a) - 1) if a[i] in ("-0", "-1", "-2")}
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
