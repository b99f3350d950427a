"""Instruction mix of the hot inner loops, from the SASS of the built object (no GPU needed).
    python tools/sass_mix.py > profiles/r1_sass_mix.txt
ALU pipe = everything that is not IMAD*/FFMA (fma pipe), LD*/ST* (lsu) or control."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "hla_mapper_b200", "build", "hlm_kernels.cu.o")
FMA = ("IMAD", "FFMA", "FMUL", "HFMA2")
LSU = ("LDG", "LDS", "STS", "STG", "LDC", "LDCU", "ATOM", "RED", "ATOMG", "LDL", "STL")
CTL = ("BRA", "BSSY", "BSYNC", "EXIT", "NOP", "WARPSYNC", "CALL", "RET", "BAR")


def loops(fn_prefix):
    txt = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
    i = txt.index("Function : " + fn_prefix)
    j = txt.find("Function :", i + 10)
    L = [l for l in txt[i:j].splitlines() if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    out = []
    for l in L:
        m = re.search(r"/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?BRA\S*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", l)
        if m:
            src, dst = int(m.group(1), 16), int(m.group(3), 16)
            if dst < src:
                out.append((src - dst, dst, src))
    out.sort(reverse=True)
    res = []
    for ln, dst, src in out:
        c = collections.Counter()
        for l in L:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
            if dst <= int(m.group(1), 16) <= src:
                c[m.group(3).split(".")[0]] += 1
        res.append((dst, src, c))
    return res


def report(name, fn_prefix, marker, min_count, per, unit):
    # the innermost loop that holds the marker instruction at least min_count times
    cands = [l for l in loops(fn_prefix) if l[2][marker] >= min_count]
    dst, src, c = min(cands, key=lambda l: l[1] - l[0])
    tot = sum(c.values())
    fma = sum(v for k, v in c.items() if k in FMA)
    lsu = sum(v for k, v in c.items() if k in LSU)
    ctl = sum(v for k, v in c.items() if k in CTL)
    alu = tot - fma - lsu - ctl
    print(f"## {name}: loop {dst:#x}..{src:#x}, {tot} instructions per {per} {unit}")
    print(f"   alu-pipe {alu} ({alu / per:.2f}/{unit[:-1]})  fma-pipe {fma} ({fma / per:.2f})  lsu {lsu}  control {ctl}")
    print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common()))


if __name__ == "__main__":
    print("# SASS instruction mix of the inner loops (cuobjdump -sass hla_mapper_b200/build/hlm_kernels.cu.o)")
    report("k_dp row loop (one 41-cell band row)", "_Z4k_dp8", "VIADDMNMX", 70, 41, "cells")
    report("k_myers main loop (8 band rows)", "_Z7k_myers8", "LDS", 20, 8, "rows")
