"""FASTQ ingest throughput of hlm_map_fastq on decoy-heavy input (the whole-genome case: almost
nothing passes the screen, so the run is bound by the text path).  Usage (GPU box):
    python tools/file_bench.py [n_pairs] [on_target] [plain|bgzf|gzip] [ENV=VALUE[,ENV=VALUE] ...]
every further argument is a set of environment variables (the library's measurement knobs) to run
the same input under, two passes each, after the default.
bgzf = the FASTQ files bgzip-compressed (inflated block-parallel by the reader), gzip = plain gzip
(zlib's single stream per file)."""
import os
import struct
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hla_mapper_b200 import _abi, api, synth


def bgzf_write(src, dst, block=0xff00):
    """bgzip: RFC 1952 members of <= 64 KiB with the BC extra subfield (SAM specification 4.1)"""
    with open(src, "rb") as f, open(dst, "wb") as o:
        while True:
            raw = f.read(block)
            c = zlib.compressobj(1, zlib.DEFLATED, -15)
            comp = c.compress(raw) + c.flush()
            o.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
                    + struct.pack("<H", len(comp) + 25) + comp
                    + struct.pack("<II", zlib.crc32(raw) & 0xffffffff, len(raw)))
            if not raw:
                return


def gzip_write(src, dst):
    import gzip
    import shutil
    import subprocess

    if shutil.which("gzip"):  # the usual writer of .fastq.gz files, at its default level
        with open(dst, "wb") as o:
            subprocess.check_call(["gzip", "-6", "-c", src], stdout=o)
        return
    with open(src, "rb") as f, gzip.open(dst, "wb", compresslevel=6) as o:
        shutil.copyfileobj(f, o, 1 << 24)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    on = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
    kind = sys.argv[3] if len(sys.argv) > 3 else "plain"
    with tempfile.TemporaryDirectory(dir="/tmp") as tmp:
        big = os.environ.get("HLM_FILE_BENCH_DB") == "40k"  # the benchmark's database instead of the small one
        db = synth.make_db(os.path.join(tmp, "db"), seed=42, n_loci=24, total_alleles=40000 if big else 4000,
                           n_others=2000 if big else 200)
        rb = synth.make_reads_fast(db, n, first_pair=0, read_len=150, seed=1003, on_target=on, indel_rate=1e-4)
        r1, r2 = os.path.join(tmp, "r1.fastq"), os.path.join(tmp, "r2.fastq")
        synth.write_fastq_fast(rb, r1, r2)
        size = os.path.getsize(r1) + os.path.getsize(r2)
        if kind != "plain":
            import threading

            ths = [threading.Thread(target=bgzf_write if kind == "bgzf" else gzip_write, args=(src, src + ".gz"))
                   for src in (r1, r2)]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()
            for src in (r1, r2):
                os.remove(src)
            r1, r2 = r1 + ".gz", r2 + ".gz"
            print(f"compressed: {(os.path.getsize(r1) + os.path.getsize(r2)) / 1e9:.2f} GB", flush=True)
        gdb = api.Database(db.path)
        m1 = api.Mapper(gdb, 0, _abi.default_params(score_cap=1))  # as the command line runs
        k = 0
        for setting in [""] + sys.argv[4:]:
            env = dict(kv.split("=", 1) for kv in setting.split(",") if kv)
            # CTX=k: k contexts on GPU 0 behind hlm_multi_map_fastq (batches of different contexts overlap)
            nctx = int(env.pop("CTX", "1"))
            m = m1 if nctx == 1 else api.MultiMapper(gdb, [0] * nctx, _abi.default_params(score_cap=1))
            os.environ.update(env)
            for rep in range(2):
                out = os.path.join(tmp, f"out{k}")
                k += 1
                os.makedirs(out)
                t = time.perf_counter()
                st = m.map_fastq(r1, r2, "W", out)
                dt = time.perf_counter() - t
                print(f"[{setting or 'default'}] rep {rep} [{kind}]: {n} pairs ({size / 1e9:.2f} GB of FASTQ) in {dt:.3f}s = "
                      f"{n / dt / 1e6:.2f} M pairs/s, {size / dt / 1e9:.2f} GB/s text; stats "
                      f"{ {a: (round(b, 3) if isinstance(b, float) else b) for a, b in st.items()} }", flush=True)
                import shutil
                shutil.rmtree(out)
            for a in env:
                del os.environ[a]
            if m is not m1:
                m.close()
        m1.close()


if __name__ == "__main__":
    main()
