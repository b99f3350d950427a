"""Measurement tool: the config-2 step (2 M pairs 2x150 against the 40k-allele DB, device-resident
arm of bench.py) under different grid sizes of the persistent kernels and different numbers of
pipeline slots.  The knobs are the environment variables HLM_TUNE / HLM_SLOTS, read when a context
is created (csrc/hlm_kernels.cu, csrc/hlm_api.cu); results go to stdout as JSON lines.

    python tools/grid_sweep.py [--pairs N] [--steps K] "myers=3" "myers=3,dp=3;slots=4" ...
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from hla_mapper_b200 import _abi, api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--db", default="/tmp/hlm_bench/db_seed42_40k")
    ap.add_argument("--capped", type=int, default=0)
    ap.add_argument("settings", nargs="*")
    args = ap.parse_args()
    w = bench.WORKLOADS["wgs_2x150"]
    db = bench.get_db(args.db, bench.DB_KW, True, lambda: None)
    rb = synth.make_reads_fast(db, args.pairs, first_pair=0, **w["reads"])
    gdb = api.Database(args.db)
    ref = None
    for setting in ["default"] + list(args.settings):
        os.environ.pop("HLM_TUNE", None)
        os.environ.pop("HLM_SLOTS", None)
        os.environ.pop("HLM_STAGGER", None)
        if setting != "default":
            for part in setting.split(";"):
                if part.startswith("slots="):
                    os.environ["HLM_SLOTS"] = part[6:]
                elif part.startswith("stagger="):
                    os.environ["HLM_STAGGER"] = part[8:]
                elif part:
                    os.environ["HLM_TUNE"] = part
        p = _abi.default_params(score_cap=args.capped)
        m = api.Mapper(gdb, 0, p)
        h = m.upload(rb)
        m.run_device(h)
        m.run_device(h)
        ms = 0.0
        stage = {}
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m.run_device(h)
            pr = m.profile()
            ms += pr["ms_total"]
            for k, v in pr.items():
                if k.startswith("ms_"):
                    stage[k] = round(stage.get(k, 0) + v / args.steps, 1)
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        T = gdb.n_loci + 1
        res = m.fetch(h, rb.n_pairs, buffers=_abi.Buffers(rb.n_pairs, T, diag=False))
        sig = (int(res.min_sum.astype("int64").sum()), int(res.target_mask.astype("int64").sum()))
        if ref is None:
            ref = sig
        m.free(h)
        m.close()
        print(json.dumps({"setting": setting, "ms_per_step": round(ms / args.steps, 2), "wall_ms": round(wall, 2),
                          "same_result": sig == ref, "stage_sum_over_slots": stage}), flush=True)
    gdb.close()


if __name__ == "__main__":
    main()
