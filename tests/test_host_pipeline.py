"""The host side of the file pipeline (csrc/hlm_files.cpp: readers, batch plan, driver threads,
collection, id sort, output preparation, writers) on a machine without a GPU: the file is compiled
unchanged and linked against a stub of the device (tests/hoststub/stub_device.cpp) whose "results"
are a function of the CRC-32 of a pair's first mate.  This module restates that function and the
reference's file formats (src/preselect.cpp:863-864, :1121-1149, :1182-1203; src/map_dna.cpp:1036-1051,
:1125-1133, :1168-1266) and predicts every output file byte for byte - for plain, gzip and bgzip
input, uniform and ramped batches, one and several contexts, paired and single end.  The GPU tests
(tests/test_golden.py) check the same files against the unmodified reference; this one covers the
host logic where no GPU is present, and - HLM_HOSTSTUB_ASAN=1 - under the sanitizers."""
import gzip
import json
import os
import random
import subprocess
import sys
import zlib

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
STUB = os.path.join(HERE, "hoststub")
ASAN = os.environ.get("HLM_HOSTSTUB_ASAN") == "1"
EXE = os.path.join(STUB, "hoststub_run_asan" if ASAN else "hoststub_run")


@pytest.fixture(scope="module", autouse=True)
def build():
    subprocess.check_call(["make", "-s", "-C", STUB] + (["asan"] if ASAN else []))


def make_reads(n, seed, paired, dup_ids=False, slash=False):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        rid = f"A00296:39:HFH3TDSXX:1:{1101 + rng.randrange(40)}:{rng.randrange(30000)}:{rng.randrange(30000)}"
        if dup_ids and i % 50 == 7 and recs:
            rid = recs[rng.randrange(len(recs))][0]
        l1, l2 = rng.randint(20, 151), rng.randint(20, 151)
        s1 = "".join(rng.choice("ACGTN") for _ in range(l1))
        s2 = "".join(rng.choice("ACGT") for _ in range(l2))
        q1 = "".join(rng.choice("?5+'") for _ in range(l1))
        q2 = "".join(rng.choice("?5+'") for _ in range(l2))
        recs.append((rid, s1, q1, s2, q2))
    def hdr(rid, mate):
        return f"@{rid}/{mate} {mate}:N:0:ACGT" if slash else f"@{rid} {mate}:N:0:ACGT"
    r1 = "".join(f"{hdr(r[0], 1)}\n{r[1]}\n+\n{r[2]}\n" for r in recs)
    r2 = "".join(f"{hdr(r[0], 2)}\n{r[3]}\n+\n{r[4]}\n" for r in recs) if paired else None
    return recs, r1, r2, hdr


def expected_files(recs, hdr, paired, G, down, sample, select_only=False):
    """what hlm_map_fastq must write for the stub's results"""
    T = G + 1
    names = [f"HLA-T{g}" for g in range(G)] + ["others"]
    sizes = [400 + 350 * g for g in range(G)]
    tag = "R1" if paired else "R0"
    pre = sample + "_"
    out = {}
    sel1, sel2, kept = [], [], []
    n_sel = sel_bases = 0

    def strip(h):
        return h.replace("/1", "").replace("/2", "") if paired else h

    for rid, s1, q1, s2, q2 in recs:
        h = zlib.crc32(s1.encode())
        L1, L2 = len(s1), (len(s2) if paired else 0)
        screened = h % 10 != 0
        keep = screened and (h >> 4) % 7 != 0
        if not screened:
            continue
        h1, h2 = strip(hdr(rid, 1)), strip(hdr(rid, 2))
        n_sel += 1
        sel_bases += L1
        sel1.append(f"{h1}\n{s1}\n+\n{q1}\n")
        sel2.append(f"{h2}\n{s2}\n+\n{q2}\n")
        if not keep:
            continue
        lo1, lo2 = (h % 5) % L1, (((h >> 3) % 4) % L2 if L2 else 0)
        n1 = max(1, L1 - lo1 - (h >> 8) % 3)
        n2 = max(1, L2 - lo2 - (h >> 10) % 2) if L2 else 0
        lmask = (h >> 12) & ((1 << G) - 1)
        sc1 = [1 + (h >> (g + 1)) % 9 if (g == T - 1 or (lmask >> g) & 1) else 0 for g in range(T)]
        sc2 = [1 + (h >> (g + 2)) % 11 if (g == T - 1 or (lmask >> g) & 1) else 0 for g in range(T)]
        visible = lmask | (((h >> 22) & 1) << G)
        target = sum(1 << g for g in range(T) if (visible >> g) & 1 and (h >> (g + 5)) & 1)
        os1, os2 = 1 + (h >> 24) % 7, 1 + (h >> 27) % 5
        tok = h1.split(" ")[0]
        idgen = tok[1:].replace("/1", "").replace("/2", "")
        kept.append(dict(id=idgen, tok=tok, h1=h1, h2=h2, s1=s1, q1=q1, s2=s2, q2=q2, lo1=lo1, n1=n1, lo2=lo2, n2=n2,
                         lmask=lmask, sc1=sc1, sc2=sc2, visible=visible, target=target, os1=os1, os2=os2))
    out[f"{pre}selected.{tag}.fastq"] = "".join(sel1)
    if paired:
        out[f"{pre}selected.R2.fastq"] = "".join(sel2)
    kept.sort(key=lambda k: k["id"].encode())  # std::map order: byte-wise; stable for equal ids
    out[f"{pre}selected.trim.{tag}.fastq"] = "".join(
        f"{k['h1']}\n{k['s1'][k['lo1']:k['lo1'] + k['n1']]}\n+\n{'A' * k['n1']}\n" for k in kept)
    if paired:
        out[f"{pre}selected.trim.R2.fastq"] = "".join(
            f"{k['h2']}\n{k['s2'][k['lo2']:k['lo2'] + k['n2']]}\n+\n{'A' * k['n2']}\n" for k in kept)
    if select_only:
        return out, n_sel, len(kept), 0
    rows = ["Read\t" + "\t".join(names) + "\tTarget\n"]
    n_assigned = 0
    for k in kept:
        cols = []
        for g in range(T):
            if not (k["visible"] >> g) & 1:
                cols.append("-")
                continue
            a, b = (k["os1"], k["os2"]) if g == T - 1 else (k["sc1"][g], k["sc2"][g])
            cols.append(f"{a},{b}" if paired else f"{a}")
        tg = [names[g] for g in range(T) if (k["target"] >> g) & 1]
        n_assigned += bool(tg)
        rows.append(k["id"] + "\t" + "\t".join(cols) + "\t" + (",".join(tg) if tg else "-") + "\n")
    out[f"{pre}addressing_table.txt"] = "".join(rows)
    mean = sel_bases // n_sel if n_sel else 0
    for g in range(G):
        mem = [k for k in kept if (k["lmask"] >> g) & 1 and (k["target"] >> g) & 1 and k["tok"][1:] == k["id"]]
        down_s = (max(down, 5) * sizes[g]) // mean if mean else 0
        out[f"{pre}{names[g]}_{tag}.fastq"] = "".join(f"{k['h1']}\n{k['s1']}\n+\n{k['q1']}\n" for k in mem)
        out[f"{pre}{names[g]}.trim.{tag}.fastq"] = "".join(
            f"{k['tok']} 1:trim\n{k['s1'][k['lo1']:k['lo1'] + k['n1']]}\n+\n{'A' * k['n1']}\n" for k in mem[:down_s])
        if paired:
            out[f"{pre}{names[g]}_R2.fastq"] = "".join(f"{k['h2']}\n{k['s2']}\n+\n{k['q2']}\n" for k in mem)
            out[f"{pre}{names[g]}.trim.R2.fastq"] = "".join(
                f"{k['tok']} 2:trim\n{k['s2'][k['lo2']:k['lo2'] + k['n2']]}\n+\n{'A' * k['n2']}\n" for k in mem[:down_s])
    return out, n_sel, len(kept), n_assigned


def run(tmp, r1, r2, batch, n_ctx, G, down, env=None, select_only=0, name="out"):
    out = os.path.join(tmp, name)
    os.makedirs(out)
    e = dict(os.environ, **(env or {}))
    if ASAN:
        e.setdefault("ASAN_OPTIONS", "detect_leaks=0")
    p = subprocess.run([EXE, r1, r2 or "-", "S", out, str(batch), str(n_ctx), str(G), str(down), str(select_only)],
                       env=e, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return out, json.loads(p.stdout)


def check(out_dir, want):
    have = sorted(os.listdir(out_dir))
    assert have == sorted(want), (set(have) ^ set(want))
    for name, text in want.items():
        got = open(os.path.join(out_dir, name), "rb").read()
        assert got == text.encode(), name


CASES = [
    # n, paired, batch, n_ctx, env, kind
    (3000, True, 257, 1, {}, "plain"),
    (3000, True, 512, 2, {"HLM_RAMP_MIN_PAIRS": "64"}, "plain"),
    (5000, True, 1000, 3, {"HLM_RAMP_MIN_PAIRS": "8"}, "gzip"),
    (4000, True, 640, 1, {"HLM_RAMP_MIN_PAIRS": "64", "HLM_PGZIP_CHUNK_KB": "8"}, "gzip"),
    (4000, True, 700, 2, {}, "bgzf"),
    (2500, False, 300, 1, {"HLM_RAMP_MIN_PAIRS": "16"}, "plain"),
    (2500, False, 1 << 19, 1, {}, "gzip"),
    (60000, True, 9000, 2, {"HLM_RAMP_MIN_PAIRS": "1000"}, "plain"),  # the parallel id sort (> 32 Ki kept reads)
]


@pytest.mark.parametrize("n,paired,batch,n_ctx,env,kind", CASES)
def test_files_are_the_predicted_ones(tmp_path, n, paired, batch, n_ctx, env, kind):
    G, down = 5, 30
    recs, t1, t2, hdr = make_reads(n, n + batch, paired, dup_ids=True, slash=(n % 2 == 0))
    r1, r2 = str(tmp_path / "r1.fastq"), (str(tmp_path / "r2.fastq") if paired else None)
    for path, text in ((r1, t1), (r2, t2)):
        if path is None:
            continue
        data = text.encode()
        if kind == "gzip":
            data = gzip.compress(data, 6)
        elif kind == "bgzf":
            sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
            import bamlite

            data = bamlite.bgzf_deflate(data, block=5000)
        open(path, "wb").write(data)
    out, st = run(str(tmp_path), r1, r2, batch, n_ctx, G, down, env)
    want, n_sel, n_kept, n_assigned = expected_files(recs, hdr, paired, G, down, "S")
    assert (st["n_pairs"], st["n_selected"], st["n_kept"], st["n_assigned"]) == (n, n_sel, n_kept, n_assigned)
    check(out, want)


def test_select_only_writes_the_selected_files(tmp_path):
    recs, t1, t2, hdr = make_reads(2000, 3, True)
    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    open(r1, "w").write(t1)
    open(r2, "w").write(t2)
    out, st = run(str(tmp_path), r1, r2, 300, 1, 4, 30, select_only=1)
    want, n_sel, n_kept, _ = expected_files(recs, hdr, True, 4, 30, "S", select_only=True)
    assert (st["n_selected"], st["n_kept"]) == (n_sel, n_kept)
    check(out, want)


def test_mismatched_mates_are_an_error(tmp_path):
    recs, t1, t2, hdr = make_reads(500, 4, True)
    r1, r2 = str(tmp_path / "r1.fastq"), str(tmp_path / "r2.fastq")
    open(r1, "w").write(t1)
    open(r2, "w").write("\n".join(t2.split("\n")[:-9]) + "\n")  # two records short
    out = str(tmp_path / "o")
    os.makedirs(out)
    p = subprocess.run([EXE, r1, r2, "S", out, "100", "1", "3", "30"], capture_output=True, text=True,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert p.returncode == 1 and "different numbers of reads" in p.stderr
