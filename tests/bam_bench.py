"""BAM ingest throughput of hlm_bam_extract + hlm_map_reads on a whole-genome-shaped BAM: a block of
records (2 % in the BED regions, 1 % unmapped, the rest elsewhere) written once with
oracle/bamlite.py and its BGZF blocks repeated K times behind one header.  Usage (GPU box):
    python tests/bam_bench.py [copies] [threads]
HLM_BAM_BENCH_EXTRACT_ONLY=1 times hlm_bam_extract alone (host code: runs without a GPU)."""
import json
import os
import struct
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bam_cases  # noqa: E402
import bamlite  # noqa: E402
from hla_mapper_b200 import _abi, api, synth  # noqa: E402


def main():
    copies = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    with tempfile.TemporaryDirectory(dir="/tmp") as tmp:
        db = synth.make_db(os.path.join(tmp, "db"), seed=42, n_loci=24, total_alleles=4000, n_others=200)
        n = 20000
        rb = synth.make_reads(db, n, read_len=150, seed=1003, on_target=1.0)
        rng = np.random.default_rng(5)
        ids = rb.ids()
        (b0, e0) = synth.SELECT_BED[0]
        recs = []
        for i in range(n):
            s1, q1 = bam_cases._mate(rb, i, 1)
            s2, q2 = bam_cases._mate(rb, i, 2)
            c = rng.random()
            if c < 0.02:
                ref, pos = 1, int(rng.integers(b0, e0 - 1000))
            elif c < 0.03:
                ref, pos = -1, -1
            else:
                ref, pos = 0, int(rng.integers(1000, 2 * 10 ** 8))
            un = ref < 0
            for fl, s, q in ((77 if un else 99, s1, q1), (141 if un else 147, s2, q2)):
                if fl & 16:
                    s, q = bamlite.revcomp(s), q[::-1]
                recs.append(dict(name=ids[i], flag=fl, ref=ref, pos=pos, mapq=60, cigar=[] if un else [("M", len(s))],
                                 next_ref=ref, next_pos=pos, tlen=0, seq=s, qual=q))
        recs.sort(key=lambda r: (r["ref"] if r["ref"] >= 0 else 1 << 30, r["pos"]))
        one = os.path.join(tmp, "one.bam")
        bamlite.write_bam(one, bam_cases.HEADER, bam_cases.REFS, recs)
        # header bytes | record bytes -> header blocks + K x record blocks + EOF block
        raw = bamlite.bgzf_inflate(open(one, "rb").read())
        l_text, = struct.unpack_from("<I", raw, 4)
        p = 8 + l_text
        n_ref, = struct.unpack_from("<I", raw, p)
        p += 4
        for _ in range(n_ref):
            l, = struct.unpack_from("<I", raw, p)
            p += 8 + l
        head = bamlite.bgzf_deflate(raw[:p])[:-28]
        body = bamlite.bgzf_deflate(raw[p:])[:-28]
        big = os.path.join(tmp, "big.bam")
        with open(big, "wb") as f:
            f.write(head)
            for _ in range(copies):
                f.write(body)
            f.write(bamlite._bgzf_block(b""))
        size = os.path.getsize(big)
        gdb = api.Database(db.path)
        extract_only = os.environ.get("HLM_BAM_BENCH_EXTRACT_ONLY") == "1"  # host only: no GPU needed
        m = None if extract_only else api.Mapper(gdb, 0, _abi.default_params(screen_variant=1))
        for rep in range(2):
            t = time.perf_counter()
            reads = api.BamReads(gdb, big, threads=threads)
            t1 = time.perf_counter() - t
            out = os.path.join(tmp, f"out{rep}")
            os.makedirs(out)
            st = {"n_kept": None} if extract_only else m.map_reads(reads, "W", out)
            t2 = time.perf_counter() - t
            s = reads.stats
            print(json.dumps({
                "rep": rep, "bam_gb": round(size / 1e9, 3), "inflated_gb": round(s["bytes_inflated"] / 1e9, 3),
                "records": s["n_records"], "threads": threads or os.cpu_count(),
                "extract_s": round(t1, 3), "pass1_s": round(s["s_pass1"], 3), "pass2_s": round(s["s_pass2"], 3),
                "extract_records_per_s": round(s["n_records"] / t1), "extract_bam_gb_per_s": round(size / t1 / 1e9, 3),
                "extract_inflated_gb_per_s": round(2 * s["bytes_inflated"] / t1 / 1e9, 3),
                "pairs_out": s["n_out"], "map_s": round(t2 - t1, 3), "kept": st["n_kept"]}))
            reads.close()
        if m:
            m.close()


if __name__ == "__main__":
    main()
