"""Randomised pinning of the oracle against the UNMODIFIED reference binary (oracle/_ref, built
from /root/reference/src by oracle/Makefile), run live through `hla-mapper dna|rna` with the
stand-in bwa / samtools / STAR.  Only runs where the reference was built (the build container);
the committed golden files (tests/golden/) carry the same evidence to the GPU box."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))

import cases  # noqa: E402
import golden_check  # noqa: E402,F401
import run_reference as rr  # noqa: E402
from hla_mapper_b200 import _abi, synth  # noqa: E402

pytestmark = pytest.mark.skipif(not rr.available(), reason="oracle/_ref not built (no /root/reference)")


def _fastq_ids(path):
    import gzip

    if not os.path.exists(path) and os.path.exists(path + ".gz"):
        path += ".gz"
    out = {}
    if not os.path.exists(path):
        return out
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline(); f.readline()
            out[h.rstrip("\n").split(" ")[0][1:].replace("/1", "").replace("/2", "")] = s
    return out


@pytest.mark.parametrize("seed", range(int(os.environ.get("HLM_FUZZ_SEEDS", "10"))))
def test_reference_binary_agrees_with_oracle_on_random_cases(orc, tmp_path, seed):
    rng = np.random.default_rng(500 + seed)
    mode = ["paired", "single", "bam", "rna", "paired"][seed % 5]
    rna = mode == "rna"
    lo = int(rng.integers(120, 500))
    db = synth.make_db(str(tmp_path / "db"), seed=int(rng.integers(1, 10 ** 6)),
                       n_loci=int(rng.integers(2, 8)), total_alleles=int(rng.integers(10, 120)),
                       len_range=(lo, lo + int(rng.integers(50, 400))), n_others=int(rng.integers(1, 8)),
                       n_lineages=int(rng.integers(1, 6)), motif_stride=int(rng.integers(1, 5)),
                       kind="rna" if rna else "dna", v_size=int(rng.choice([30, 50])),
                       paralog_div=(0.002, float(rng.choice([0.01, 0.1]))),
                       others_div=(0.002, float(rng.choice([0.02, 0.2]))))
    L = int(rng.choice([60, 100, 150]))
    n = 300
    rb = synth.make_reads(db, n, read_len=L, seed=int(rng.integers(1, 10 ** 6)),
                          on_target=float(rng.choice([0.6, 1.0])), frag_mu=int(rng.integers(L, 3 * L)),
                          frag_sd=int(rng.integers(1, 50)), lowq_tail=float(rng.choice([0.0, 0.2, 0.6])),
                          ragged=bool(rng.integers(0, 2)))
    suffix = bool(rng.integers(0, 2)) and mode != "single"
    comment = bool(rng.integers(0, 2))
    if mode == "single":
        r0 = str(tmp_path / "r0.fastq")
        synth.write_fastq(rb, r0, None, suffix=False, comment=comment)
        out = rr.run_reference(db.path, str(tmp_path / "out"), "F", r0=r0)
    else:
        r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
        synth.write_fastq(rb, r1, r2, suffix=suffix and mode != "bam", comment=comment)
        out = rr.run_reference(db.path, str(tmp_path / "out"), "F", rna=rna,
                               **({"bam": (r1, r2)} if mode == "bam" else {"r1": r1, "r2": r2}))
    p = _abi.default_params(rna=int(rna), paired=int(mode != "single"), screen_variant=int(mode == "bam"))
    odb = orc.OracleDB(db.path, rna=int(rna))
    so = None
    if mode == "single":
        so = np.array([len("@" + i + (" 1:N:0:ACGT" if comment else "")) for i in rb.ids()], dtype=np.int32)
    if rna:
        buf, _ = orc.run(odb, p, rb, stages=("screen",))
        sp1, sp2 = cases.rna_splits(rb, buf)
        buf, _ = orc.run(odb, p, rb, mode=1, stages=("score", "assign"), buffers=buf, rna_split1=sp1, rna_split2=sp2)
    else:
        buf, _ = orc.run(odb, p, rb, mode=1, size_override1=so)
    ids = rb.ids()
    tag = "R0" if mode == "single" else "R1"
    assert sorted(_fastq_ids(out + f"F_selected.{tag}.fastq")) == sorted(i for i, f in zip(ids, buf.flags) if f & 1)
    trim = _fastq_ids(out + f"F_selected.trim.{tag}.fastq")
    kept = [k for k in range(n) if buf.flags[k] & 2]
    assert sorted(trim) == sorted(ids[k] for k in kept)
    for k in kept:
        a = int(rb.offs1[k]) + int(buf.trim_lo1[k])
        assert rb.bases1[a:a + int(buf.trim_len1[k])].tobytes().decode() == trim[ids[k]]
    table = out + "F_addressing_table.txt"
    if not kept:
        return
    names, rows = rr.parse_addressing_table(table)
    assert names == odb.loci
    want = rr.expected_table(buf, ids, odb.loci, rna=rna, paired=mode != "single")
    assert set(rows) == set(want)
    for rid in want:
        assert [rows[rid][0], rows[rid][1]] == [want[rid][0], want[rid][1]], (mode, rid)
