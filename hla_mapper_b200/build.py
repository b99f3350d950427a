"""Builds libhlamap.so (the C-ABI library: CUDA kernels for sm_100a + C++ host runtime) in-tree
with nvcc.  No torch, no JIT cache: the .so sits next to this file so it travels with the repo
snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhlamap.so")
CLI = os.path.join(HERE, "hla-mapper-b200")
SOURCES = ["hlm_api.cu", "hlm_kernels.cu", "hlm_db.cpp", "hlm_files.cpp", "hlm_bam.cpp"]
HEADERS = ["hlm_cli.cpp", "hlm_common.h", "hlm_files.h", "hlm_bgzf.h", "hlm_pgzip.h", "hlm_crc32.h", "hlm_core.cuh", "hlm_db.h", "hlm_kernels.cuh", "../../include/hlamap.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        objs.append(obj)
        cmd = [nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    log = []
    for src, p in procs:
        out = p.communicate()[0].decode()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    subprocess.check_call([nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"])
    # the thin `hla-mapper dna|select` command line over the library
    subprocess.check_call([nvcc(), "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "hlm_cli.cpp"),
                           "-L" + HERE, "-lhlamap", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    # instruction mix of the inner loops of THIS object: bench.py's roofline denominators
    subprocess.check_call([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "sass_mix.py"), "--json",
                           os.path.join(HERE, "build", "sass_mix.json")])
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
