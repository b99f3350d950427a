"""Generates tests/golden/*.json.gz by running the UNMODIFIED reference binary
(oracle/_ref/hla-mapper-ref, built from /root/reference/src by oracle/Makefile) on seeded
synthetic inputs.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What is recorded per case is what the reference itself wrote: the ids in selected.R*.fastq, the
trimmed sequences of selected.trim.R*.fastq, the sorted_<gene> candidate lists, every row of
<sample>_addressing_table.txt and the per-gene <sample>_<gene>_R1.fastq membership.  The
aligner inside the stand-in `bwa` is the oracle's (the real bwa is not available: the scores
themselves are therefore "parity unpinned", everything around them is the reference's code)."""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import run_reference as rr  # noqa: E402
from hla_mapper_b200 import synth  # noqa: E402


def fastq_records(path):
    out = {}
    if not os.path.exists(path) and os.path.exists(path + ".gz"):  # `rna` gzips at the end
        path += ".gz"
    if not os.path.exists(path):
        return out
    with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline()
            f.readline()
            key = h.rstrip("\n").split(" ")[0][1:].replace("/1", "").replace("/2", "")
            out[key] = s
    return out


def bam_case(name):
    """bam=<real BAM>: the reference's whole bam= front end, samtools served by the stand-in"""
    with tempfile.TemporaryDirectory() as tmp:
        db, bam, p = cases.bam_inputs(name, tmp)
        out = rr.run_reference(db.path, os.path.join(tmp, "out"), "G", bam=bam)
        tag = "R1" if p.paired else "R0"
        names, rows = rr.parse_addressing_table(out + "G_addressing_table.txt")
        want = [f"G_selected.{tag}.fastq", f"G_selected.trim.{tag}.fastq", "G_addressing_table.txt"]
        for g in names[:-1]:
            want += [f"G_{g}_{tag}.fastq", f"G_{g}.trim.{tag}.fastq"]
        if p.paired:
            want += [w.replace(tag, "R2") for w in want if tag in w]
        files = {w: cases.canonical_digest(out + w) for w in want}
        # the bam= path writes selected.R*.fastq from a std::map: that order is defined
        exact = {w: cases.raw_digest(out + w) for w in want if w.startswith("G_selected.")}
        trim1 = fastq_records(out + f"G_selected.trim.{tag}.fastq")
        sel = fastq_records(out + f"G_selected.{tag}.fastq")
        gold = {"case": name, "mode": "bamfile", "targets": names, "files": files, "files_exact": exact,
                "n_selected": len(sel), "n_kept": len(trim1), "n_rows": len(rows),
                "paired": int(p.paired)}
        path = os.path.join(HERE, name + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(json.dumps(gold, sort_keys=True).encode())
        multi = sum(1 for v in rows.values() if "," in v[1])
        print(f"{name}: {len(sel)} selected, {len(trim1)} kept, {len(rows)} rows ({multi} multi-target)"
              f" -> {os.path.getsize(path)} bytes")


def main():
    assert rr.available(), "build oracle/_ref first: make -C oracle ref"
    only = set(sys.argv[1:])
    for name in cases.BAM_CASES:
        if not only or name in only:
            bam_case(name)
    for name in cases.CASES:
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            db, rb, p, mode = cases.inputs(name, tmp)
            r1, r2 = cases.write_inputs(name, rb, mode, tmp)
            if mode == "bam":
                out = rr.run_reference(db.path, os.path.join(tmp, "out"), "G", bam=(r1, r2))
                tag = "R1"
            elif mode != "single":
                out = rr.run_reference(db.path, os.path.join(tmp, "out"), "G", r1=r1, r2=r2,
                                       rna=mode in ("rna", "rna_full"), star_full=(mode == "rna_full"))
                tag = "R1"
            else:
                out = rr.run_reference(db.path, os.path.join(tmp, "out"), "G", r0=r1)
                tag = "R0"
            names, rows = rr.parse_addressing_table(out + "G_addressing_table.txt")
            sel = fastq_records(out + f"G_selected.{tag}.fastq")
            trim1 = fastq_records(out + f"G_selected.trim.{tag}.fastq")
            trim2 = fastq_records(out + "G_selected.trim.R2.fastq") if mode != "single" else {}
            sorted_ = {g: sorted(fastq_records(out + f"G_sorted_{g}_{tag}.fastq")) for g in names[:-1]}
            emitted = {g: sorted(fastq_records(out + f"G_{g}_{tag}.fastq")) for g in names[:-1]}
            files = {}
            if mode in ("paired", "single"):  # byte-level digests of what the reference wrote (file pipeline)
                want = [f"G_selected.{tag}.fastq", f"G_selected.trim.{tag}.fastq", "G_addressing_table.txt"]
                for g in names[:-1]:
                    want += [f"G_{g}_{tag}.fastq", f"G_{g}.trim.{tag}.fastq"]
                if mode == "paired":
                    want += [w.replace(tag, "R2") for w in want if tag in w]
                for w in want:
                    files[w] = cases.canonical_digest(out + w)
            if mode in ("rna", "rna_full"):
                # the rna file pipeline: the same files plus what goes into / comes out of STAR
                want = ["G_selected.R1.fastq", "G_selected.trim.R1.fastq", "G_addressing_table.txt"]
                for g in names[:-1]:
                    want += [f"G_{g}_R1.fastq", f"G_{g}.trim.R1.fastq", f"G_sorted_{g}_R1.fastq"]
                want += [w.replace("R1", "R2") for w in want if "R1" in w]
                want += [f"G_sorted_spliced_{g}.fastq" for g in names[:-1]]
                for w in want:
                    files[w] = cases.canonical_digest(out + w)
            gold = {
                "case": name, "mode": mode, "targets": names, "files": files,
                "selected": sorted(sel),
                "trim1": trim1, "trim2": trim2,
                "sorted": sorted_, "emitted": emitted,
                "table": {k: [v[0], v[1]] for k, v in rows.items()},
            }
            path = os.path.join(HERE, name + ".json.gz")
            with gzip.GzipFile(path, "wb", mtime=0) as f:
                f.write(json.dumps(gold, sort_keys=True).encode())
            multi = sum(1 for v in rows.values() if "," in v[1])
            print(f"{name}: {len(sel)} selected, {len(trim1)} kept, {len(rows)} rows "
                  f"({multi} multi-target) -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
