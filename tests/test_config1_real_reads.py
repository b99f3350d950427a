"""BASELINE.json configs[0]: the reference's own test reads (test_sequences/HG01890_HLA-A_R{1,2}
.fastq, 1 462 pairs of 2x150 bp Illumina data) through `hla-mapper dna` on the CPU.

The real hla-mapper database is not mounted, so a database is derived from the reads themselves
(seed 20240813): HLA-A "alleles" are fragments spanned by read pairs, a second locus holds 4 %
diverged copies, `others` 12 % diverged ones.  The UNMODIFIED reference binary is run on the real
FASTQ files with that database and must agree with the oracle on every addressing-table row.
CPU-only and only where /root/reference is mounted: the reads are reference data and are not
copied into this repository."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import run_reference as rr  # noqa: E402
from hla_mapper_b200 import _abi, synth  # noqa: E402

R1 = "/root/reference/test_sequences/HG01890_HLA-A_R1.fastq"
R2 = "/root/reference/test_sequences/HG01890_HLA-A_R2.fastq"
pytestmark = pytest.mark.skipif(not (os.path.exists(R1) and rr.available()),
                                reason="reference test_sequences / oracle/_ref not available")
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _derive_db(path, rb, rng):
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

    def mut(s, rate):
        a = np.frombuffer(s, dtype=np.uint8).copy()
        m = rng.random(len(a)) < rate
        a[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        return a.tobytes()

    frags = []
    for i in rng.choice(rb.n_pairs, size=60, replace=False):
        a = rb.bases1[int(rb.offs1[i]):int(rb.offs1[i + 1])].tobytes()
        b = rb.bases2[int(rb.offs2[i]):int(rb.offs2[i + 1])].tobytes()
        if b"N" in a or b"N" in b:
            continue
        frags.append(a + acgt[rng.integers(0, 4, size=80)].tobytes() + b.translate(COMP)[::-1])
    loci = {"HLA-A": frags, "HLA-H": [mut(f, 0.04) for f in frags[:30]]}
    others = [mut(f, 0.12) for f in frags[:20]]
    os.makedirs(os.path.join(path, "motif/dna/all"), exist_ok=True)
    os.makedirs(os.path.join(path, "motif/dna/loci"), exist_ok=True)
    os.makedirs(os.path.join(path, "others/dna"), exist_ok=True)
    with open(os.path.join(path, "db_dna.info"), "w") as f:
        f.write("hla-mapper:4.3+\n")
        for k, g in enumerate(loci):
            f.write(f"GENE:{g}:chr6:{29900000 + 1000 * k}:170805979:1:3500:HLA\n")
        f.write("SIZE:50\n")
    allm = set()
    for g, als in loci.items():
        os.makedirs(os.path.join(path, "mapper/dna", g), exist_ok=True)
        with open(os.path.join(path, "mapper/dna", g, g + ".fas"), "wb") as f:
            for i, s in enumerate(als):
                f.write(b">%s*%d\n%s\n" % (g.encode(), i, s))
        mot = {s[c:c + 20] for s in als for c in range(0, len(s) - 19, 2)}
        mot |= {k.translate(COMP)[::-1] for k in mot}
        allm |= mot
        with open(os.path.join(path, "motif/dna/loci", g + ".txt"), "wb") as f:
            f.write(b"\n".join(sorted(mot)) + b"\n")
    with open(os.path.join(path, "others/dna/hla_gen_others.fasta"), "wb") as f:
        for i, s in enumerate(others):
            f.write(b">o%d\n%s\n" % (i, s))
    with open(os.path.join(path, "motif/dna/all/motif_all.txt"), "wb") as f:
        f.write(b"\n".join(sorted(allm)) + b"\n")


def test_reference_test_sequences_through_hla_mapper_dna(orc, tmp_path):
    rb, headers = synth.read_fastq_pair(R1, R2)
    assert rb.n_pairs == 1462
    rng = np.random.default_rng(20240813)
    dbp = str(tmp_path / "db")
    _derive_db(dbp, rb, rng)
    out = rr.run_reference(dbp, str(tmp_path / "out"), "HG01890", r1=R1, r2=R2, threads=4)
    names, rows = rr.parse_addressing_table(out + "HG01890_addressing_table.txt")
    odb = orc.OracleDB(dbp)
    buf, _ = orc.run(odb, _abi.default_params(), rb, mode=1)
    ids = [h.split(" ")[0][1:] for h in headers]
    want = rr.expected_table(buf, ids, odb.loci)
    assert names == odb.loci and set(rows) == set(want) and len(rows) > 500
    for rid in want:
        assert [rows[rid][0], rows[rid][1]] == [want[rid][0], want[rid][1]], rid
    assert sum(1 for v in rows.values() if v[1] == "HLA-A") > 200
