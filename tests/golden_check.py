"""Compares a result buffer (from the oracle or from libhlamap) with a golden file written by
tests/golden/make_golden.py, i.e. with what the unmodified reference binary produced."""
import gzip
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with gzip.open(os.path.join(HERE, "golden", name + ".json.gz"), "rb") as f:
        return json.loads(f.read().decode())


def check(buf, rb, loci, gold, paired, rna=False):
    """loci = DB loci + ['others'] in DB order"""
    import run_reference as rr  # table formatting helper (oracle/, test infrastructure)

    ids = rb.ids()
    assert gold["targets"] == loci
    flags = buf.flags
    # a-1: the global screen (selected.R*.fastq)
    sel = sorted(i for i, f in zip(ids, flags) if f & 1)
    assert sel == gold["selected"]
    # a-4 / a-5: mtrim + pair filter (selected.trim.R*.fastq holds the trimmed sequences)
    kept = [k for k in range(rb.n_pairs) if flags[k] & 2]
    assert sorted(ids[k] for k in kept) == sorted(gold["trim1"])
    for k in kept:
        a = int(rb.offs1[k]) + int(buf.trim_lo1[k])
        s1 = rb.bases1[a:a + int(buf.trim_len1[k])].tobytes().decode()
        assert s1 == gold["trim1"][ids[k]], ids[k]
        if paired:
            a = int(rb.offs2[k]) + int(buf.trim_lo2[k])
            s2 = rb.bases2[a:a + int(buf.trim_len2[k])].tobytes().decode()
            assert s2 == gold["trim2"][ids[k]], ids[k]
    # a-6: per-locus sort (sorted_<gene>_R1.fastq)
    for g, name in enumerate(loci[:-1]):
        mine = sorted(ids[k] for k in kept if (int(buf.locus_mask[k]) >> g) & 1)
        assert mine == gold["sorted"][name], name
    # a-8 / a-10 / a-11 / a-13: every row of the addressing table
    rows = rr.expected_table(buf, ids, loci, rna=rna, paired=paired)
    assert set(rows) == set(gold["table"])
    for rid, (cols, tgt) in rows.items():
        assert [cols, tgt] == gold["table"][rid], rid
    # a-14: per-gene FASTQ emit = sorted into the gene AND addressed to it
    for g, name in enumerate(loci[:-1]):
        mine = sorted(ids[k] for k in kept
                      if (int(buf.locus_mask[k]) >> g) & 1 and (int(buf.target_mask[k]) >> g) & 1)
        assert mine == gold["emitted"][name], name
    return len(rows)
