"""BASELINE.json configs[0]: the reference's own test reads (test_sequences/HG01890_HLA-A_R{1,2}
.fastq, 1 462 pairs of 2x150 bp Illumina data) through `hla-mapper dna`.

The real hla-mapper database is not mounted, so a database is derived from the reads themselves
(seed 20240813): HLA-A "alleles" are fragments spanned by read pairs, a second locus holds 4 %
diverged copies, `others` 12 % diverged ones.  The reads are committed as fixtures
(tests/golden/HG01890_HLA-A_R*.fastq.gz, gzip of the reference's files) together with the
addressing table the UNMODIFIED reference binary wrote for them on that database
(tests/golden/config1_hg01890.json.gz; `python tests/test_config1_real_reads.py` regenerates it
where /root/reference is mounted).  CPU: the reference run live (where available), the golden
table and the oracle agree on every row.  GPU: libhlamap's file pipeline on the gzip files writes
that table, and its buffers equal the oracle's."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import run_reference as rr  # noqa: E402
from hla_mapper_b200 import _abi, synth  # noqa: E402

GZ1 = os.path.join(HERE, "golden", "HG01890_HLA-A_R1.fastq.gz")
GZ2 = os.path.join(HERE, "golden", "HG01890_HLA-A_R2.fastq.gz")
GOLD = os.path.join(HERE, "golden", "config1_hg01890.json.gz")
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _plain(tmp):
    """the fixtures as plain FASTQ files (what the reference reads)"""
    out = []
    for k, z in enumerate((GZ1, GZ2)):
        p = os.path.join(str(tmp), f"HG01890_HLA-A_R{k + 1}.fastq")
        with gzip.open(z, "rb") as f, open(p, "wb") as o:
            o.write(f.read())
        out.append(p)
    return out


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


def _setup(tmp_path):
    r1, r2 = _plain(tmp_path)
    rb, headers = synth.read_fastq_pair(r1, r2)
    assert rb.n_pairs == 1462
    dbp = str(tmp_path / "db")
    _derive_db(dbp, rb, np.random.default_rng(20240813))
    ids = [h.split(" ")[0][1:] for h in headers]
    return r1, r2, rb, ids, dbp


def _golden():
    g = json.load(gzip.open(GOLD))
    return g["targets"], {k: (v[0], v[1]) for k, v in g["table"].items()}


def test_reference_test_sequences_through_hla_mapper_dna(orc, tmp_path):
    r1, r2, rb, ids, dbp = _setup(tmp_path)
    names, rows = _golden()
    if rr.available():  # the unmodified reference, live: it wrote the golden table
        out = rr.run_reference(dbp, str(tmp_path / "out"), "HG01890", r1=r1, r2=r2, threads=4)
        live_names, live = rr.parse_addressing_table(out + "HG01890_addressing_table.txt")
        assert live_names == names and {k: (v[0], v[1]) for k, v in live.items()} == rows
    odb = orc.OracleDB(dbp)
    buf, _ = orc.run(odb, _abi.default_params(), rb, mode=1)
    want = rr.expected_table(buf, ids, odb.loci)
    assert names == odb.loci and set(rows) == set(want) and len(rows) > 500
    for rid in want:
        assert [rows[rid][0], rows[rid][1]] == [want[rid][0], want[rid][1]], rid
    assert sum(1 for v in rows.values() if v[1] == "HLA-A") > 200


@pytest.mark.gpu
def test_gpu_config1_real_reads(orc, tmp_path):
    """the reference's own reads through libhlamap: buffers equal the oracle's, and the file
    pipeline (gzip FASTQ in) writes the addressing table the unmodified reference wrote"""
    from hla_mapper_b200 import api

    r1, r2, rb, ids, dbp = _setup(tmp_path)
    names, rows = _golden()
    p = _abi.default_params()
    want, _ = orc.run(orc.OracleDB(dbp), p, rb, mode=0)
    gdb = api.Database(dbp)
    m = api.Mapper(gdb, 0, p)
    got = m.run(rb)
    for k in want.FIELDS:
        assert np.array_equal(getattr(want, k), getattr(got, k)), k
    out = tmp_path / "gpu_out"
    out.mkdir()
    st = m.map_fastq(GZ1, GZ2, "HG01890", out)
    assert st["n_pairs"] == 1462
    gnames, grows = rr.parse_addressing_table(str(out / "HG01890_addressing_table.txt"))
    assert gnames == names and {k: (v[0], v[1]) for k, v in grows.items()} == rows
    m.close()
    gdb.close()


if __name__ == "__main__":  # regenerate the golden table (build container only)
    import tempfile

    assert rr.available(), "build oracle/_ref first: make -C oracle ref"
    with tempfile.TemporaryDirectory() as tmp:
        import pathlib

        r1, r2, rb, ids, dbp = _setup(pathlib.Path(tmp))
        out = rr.run_reference(dbp, os.path.join(tmp, "out"), "HG01890", r1=r1, r2=r2, threads=4)
        names, rows = rr.parse_addressing_table(out + "HG01890_addressing_table.txt")
        with gzip.GzipFile(GOLD, "wb", mtime=0) as f:
            f.write(json.dumps({"targets": names, "table": {k: [v[0], v[1]] for k, v in rows.items()}},
                               sort_keys=True).encode())
        print(len(rows), "rows ->", GOLD)
