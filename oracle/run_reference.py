"""Drive the UNMODIFIED reference binary (oracle/_ref/hla-mapper-ref, built by oracle/Makefile
from /root/reference/src with a Boost header shim) through `hla-mapper dna|rna`, with
`bwa` replaced by oracle/_ref/bwa (stand-in that calls the oracle aligner) and `samtools` by
/bin/true.  Used to pin the oracle's screen / trim / locus sort / SAM reducer / compare
restatements against the reference's own code, and to generate tests/golden/*.

TEST INFRASTRUCTURE ONLY.  The reference reads its config from getpwuid()->pw_dir/.hla-mapper
(/root/reference/src/main.cpp:266-267).  Every run gets a private directory for it: the
binary is started with oracle/_ref/libpwshim.so preloaded (oracle/shim/pwshim.c), which answers
getpwuid() with pw_dir = $HLM_REF_HOME.  The real ~/.hla-mapper is never touched, so concurrent
harness processes (pytest-xdist, the bench arm beside a test run) cannot disturb each other.
"""
from __future__ import annotations

import contextlib
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "hla-mapper-ref")
BWA_BIN = os.path.join(HERE, "_ref", "bwa")
PWSHIM = os.path.join(HERE, "_ref", "libpwshim.so")
STAR_BIN = os.path.join(HERE, "standin_star.py")
SAMTOOLS_BIN = os.path.join(HERE, "standin_samtools.py")


def available():
    return os.path.exists(REF_BIN) and os.path.exists(BWA_BIN) and os.path.exists(PWSHIM)


@contextlib.contextmanager
def _config(db_path, env):
    """a private home directory holding .hla-mapper; env gets LD_PRELOAD + HLM_REF_HOME"""
    home = tempfile.mkdtemp(prefix="hlm_ref_home_")
    with open(os.path.join(home, ".hla-mapper"), "w") as f:
        f.write(f"db={db_path}\nsamtools={SAMTOOLS_BIN}\nbwa={BWA_BIN}\nwhatshap=DISABLED\n"
                f"freebayes=DISABLED\nblast=DISABLED\nstar={STAR_BIN}\n")
    env["HLM_REF_HOME"] = home
    env["LD_PRELOAD"] = PWSHIM + (":" + env["LD_PRELOAD"] if env.get("LD_PRELOAD") else "")
    try:
        yield
    finally:
        shutil.rmtree(home, ignore_errors=True)


def run_reference(db_path, out_dir, sample, r1=None, r2=None, r0=None, bam=None, rna=False, threads=4,
                  debug=True, exhaustive=False, extra=(), timeout=3600, star_full=False):
    """returns the output directory (with trailing slash, as the reference stores it)"""
    if os.path.exists(out_dir):
        shutil.rmtree(out_dir)
    cmd = [REF_BIN, "rna" if rna else "dna"]
    env = dict(os.environ)
    if isinstance(bam, str):  # a real BAM file: parsed by the stand-in samtools (oracle/bamlite.py)
        env.pop("HLM_STANDIN_BAM", None)
        cmd.append(f"bam={bam}")
    elif bam:  # bam = (r1, r2): the stand-in samtools serves them as the content of a "BAM"
        os.makedirs(os.path.dirname(out_dir.rstrip("/")) or ".", exist_ok=True)
        manifest = out_dir.rstrip("/") + ".standin.bam"
        with open(manifest, "w") as f:
            f.write(f"{bam[0]}\n{bam[1]}\n")
        env["HLM_STANDIN_BAM"] = manifest
        cmd.append(f"bam={manifest}")
    elif r0:
        cmd.append(f"r0={r0}")
    else:
        cmd += [f"r1={r1}", f"r2={r2}"]
    cmd += [f"sample={sample}", f"output={out_dir}", f"threads={threads}", "--skip-adjust",
            "--quiet"]
    if debug:
        cmd.append("--debug")
    cmd += list(extra)
    if exhaustive:
        env["HLM_ORACLE_EXHAUSTIVE"] = "1"
    if star_full:  # oracle/standin_star.py: gene-dependent splits, multi-mappers, reverse strand
        env["HLM_STANDIN_STAR"] = "full"
    with _config(db_path, env):
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env,
                       timeout=timeout, cwd=os.path.dirname(out_dir.rstrip("/")) or ".")
    return out_dir.rstrip("/") + "/"


def parse_addressing_table(path):
    """-> (target names, {read id: (score strings per target, Target string)})"""
    rows = {}
    with open(path) as f:
        hdr = f.readline().rstrip("\n").split("\t")
        names = hdr[1:-1]
        for line in f:
            t = line.rstrip("\n").split("\t")
            rows[t[0]] = (t[1:-1], t[-1])
    return names, rows


def parse_fastq_ids(path, with_len=False):
    out = {}
    with open(path) as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline()
            f.readline()
            key = h.rstrip("\n").split(" ")[0][1:].replace("/1", "").replace("/2", "")
            out[key] = s if not with_len else len(s)
    return out


def expected_table(buf, ids, loci, rna=False, paired=True):
    """what the reference's addressing table must contain, from oracle/GPU output buffers.
    loci = DB loci + ['others'].  Mirrors map_dna.cpp:932-1016 / map_rna.cpp:945-1013."""
    rows = {}
    T = len(loci)
    for i, rid in enumerate(ids):
        if not (buf.flags[i] & 2):
            continue
        cols = []
        for g in range(T):
            if not rna:
                if (int(buf.visible_mask[i]) >> g) & 1:
                    s1 = int(buf.others_s1[i]) if g == T - 1 else int(buf.s1[i, g])
                    s2 = int(buf.others_s2[i]) if g == T - 1 else int(buf.s2[i, g])
                    cols.append(f"{s1},{s2}" if paired else f"{s1}")
                else:
                    cols.append("-")
            else:
                cols.append(str(int(buf.s1[i, g])) if (int(buf.visible_mask[i]) >> g) & 1 else "-")
        tg = [loci[g] for g in range(T) if (int(buf.target_mask[i]) >> g) & 1]
        rows[rid] = (cols, ",".join(tg) if tg else "-")
    return rows
