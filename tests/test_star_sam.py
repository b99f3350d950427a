"""hlm_star_sam_read: STAR's SAM -> the block records of
src/map_rna.cpp:740-790, host-only.  The SAM comes from the STAR stand-in in its full mode
(per-gene splits, three-block alignments - whose third block starts at the SIZE of the second,
:784 -, secondary alignments, reverse-strand and unmapped records); the expected ranges are the
stand-in's own statement of what the reference cuts (`standin_star.blocks`), which the golden
`rna_blocks` case pins to the files the unmodified reference wrote."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import standin_star  # noqa: E402


def test_star_sam_reader_cuts_the_reference_blocks(tmp_path, monkeypatch):
    from hla_mapper_b200 import api

    rng = np.random.default_rng(4)
    n = 400
    mates = []
    for m in range(2):
        path = tmp_path / f"sorted_G_R{m + 1}.fastq"
        seqs = []
        with open(path, "w") as f:
            for i in range(n):
                L = int(rng.integers(30, 151))
                s = "".join("ACGT"[x] for x in rng.integers(0, 4, size=L))
                seqs.append(s)
                f.write(f"@read{i}/{m + 1} c\n{s}\n+\n{'A' * L}\n")
        mates.append(seqs)
    monkeypatch.setenv("HLM_STANDIN_STAR", "full")
    gene = "HLA-B"
    prefix = str(tmp_path / "S_HLA-B_splice_")
    assert standin_star.main(["--genomeDir", str(tmp_path / "loci" / gene) + "/", "--readFilesIn",
                              str(tmp_path / "sorted_G_R1.fastq"), str(tmp_path / "sorted_G_R2.fastq"),
                              "--outFileNamePrefix", prefix]) == 0
    recs = api.star_sam_blocks(prefix + "Aligned.out.sam", 3)
    got = {}
    for qname, flag, seq_len, piece, mate, start, ln in recs:
        got.setdefault((qname, mate), []).append((start, ln))
        assert seq_len == len(mates[mate][int(qname[4:])])
    kinds = set()
    for m in range(2):
        for i in range(n):
            want = standin_star.blocks(mates[m][i].encode(), gene)
            if m == 1:
                # the stand-in gives second mates FLAG 147 (0x10 set) like a proper pair but prints
                # SEQ as read (and clears 0x10 where it does reverse-complement): against its own
                # FLAG bits every range is mirrored.  (hlm_map_fastq decides by content.)
                L = len(mates[m][i])
                want = [(L - st - ln, ln) for st, ln in want]
            assert got[(f"read{i}", m)] == want, (i, m)
            kinds.add(len(want))
    assert {1, 2, 3} <= kinds and len(recs) > 2 * n
    # a missing file is "no record", as for the reference's reader
    assert api.star_sam_blocks(str(tmp_path / "nope.sam"), 0) == []
