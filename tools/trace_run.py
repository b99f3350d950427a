"""Measurement tool: the stage-event timeline of one device-resident config-2 run (HLM_TRACE),
per chunk and pipeline slot, under the settings given on the command line (see grid_sweep.py).

    python tools/trace_run.py --pairs 1048576 --out gpurun_out/trace "default" "myers=3" "slots=1"
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from hla_mapper_b200 import _abi, api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--db", default="/tmp/hlm_bench/db_seed42_40k")
    ap.add_argument("--out", default="gpurun_out/trace")
    ap.add_argument("--capped", type=int, default=0)
    ap.add_argument("settings", nargs="*", default=["default"])
    args = ap.parse_args()
    w = bench.WORKLOADS["wgs_2x150"]
    db = bench.get_db(args.db, bench.DB_KW, True, lambda: None)
    rb = synth.make_reads_fast(db, args.pairs, first_pair=0, **w["reads"])
    gdb = api.Database(args.db)
    for k, setting in enumerate(args.settings):
        os.environ.pop("HLM_TUNE", None)
        os.environ.pop("HLM_SLOTS", None)
        os.environ.pop("HLM_STAGGER", None)
        os.environ["HLM_TRACE"] = f"{args.out}_{k}.txt"
        if setting != "default":
            for part in setting.split(";"):
                if part.startswith("slots="):
                    os.environ["HLM_SLOTS"] = part[6:]
                elif part.startswith("stagger="):
                    os.environ["HLM_STAGGER"] = part[8:]
                elif part:
                    os.environ["HLM_TUNE"] = part
        m = api.Mapper(gdb, 0, _abi.default_params(score_cap=args.capped))
        h = m.upload(rb)
        for _ in range(3):
            m.run_device(h)
        print(setting, m.profile()["ms_total"], flush=True)
        m.free(h)
        m.close()
        with open(os.environ["HLM_TRACE"], "a") as f:
            f.write(f"# {setting}\n")
    gdb.close()


if __name__ == "__main__":
    main()
