"""Instruction mix of the hot inner loops, from the SASS of the built object (no GPU needed).
    python tools/sass_mix.py > profiles/r2_sass_mix.txt
    python tools/sass_mix.py --json hla_mapper_b200/build/sass_mix.json   (done by build())
ALU pipe = everything that is not IMAD*/FFMA (fma pipe), LD*/ST* (lsu) or control.
bench.py takes its instructions-per-cell / per-row figures from mix() - i.e. from the object
that is being timed - so the roofline denominators cannot go stale."""
import collections
import json
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


def analyse(fn_prefix, marker, min_count, per):
    # the innermost loop that holds the marker instruction at least min_count times
    cands = [l for l in loops(fn_prefix) if l[2][marker] >= min_count]
    dst, src, c = min(cands, key=lambda l: l[1] - l[0])
    tot = sum(c.values())
    fma = sum(v for k, v in c.items() if k in FMA)
    lsu = sum(v for k, v in c.items() if k in LSU)
    ctl = sum(v for k, v in c.items() if k in CTL)
    alu = tot - fma - lsu - ctl
    return {"loop": [dst, src], "instructions": tot, "per": per, "alu": alu, "fma": fma, "lsu": lsu,
            "control": ctl, "alu_per_unit": alu / per, "fma_per_unit": fma / per,
            "int_per_unit": (alu + fma) / per, "mix": dict(c.most_common())}


SPECS = {
    # kernel: (mangled prefix, marker instruction, min count in the loop, units per iteration, unit)
    "k_dp": ("_Z4k_dp8", "VIADDMNMX", 70, 41, "cells", "k_dp row loop (one 41-cell band row)"),
    "k_myers": ("_Z7k_myers8", "LDS", 20, 8, "rows", "k_myers main loop (8 band rows)"),
}


def mix():
    """{kernel: instruction mix of its inner loop} from the built object"""
    return {k: analyse(v[0], v[1], v[2], v[3]) for k, v in SPECS.items()}


def report(name, a, unit):
    per = a["per"]
    print(f"## {name}: loop {a['loop'][0]:#x}..{a['loop'][1]:#x}, {a['instructions']} instructions per {per} {unit}")
    print(f"   alu-pipe {a['alu']} ({a['alu'] / per:.2f}/{unit[:-1]})  fma-pipe {a['fma']} ({a['fma'] / per:.2f})  "
          f"lsu {a['lsu']}  control {a['control']}")
    print("   " + ", ".join(f"{k} {v}" for k, v in a["mix"].items()))


if __name__ == "__main__":
    m = mix()
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        json.dump(m, open(sys.argv[2], "w"), indent=1)
        sys.exit(0)
    print("# SASS instruction mix of the inner loops (cuobjdump -sass hla_mapper_b200/build/hlm_kernels.cu.o)")
    for k, v in SPECS.items():
        report(v[5], m[k], v[4])
