#!/usr/bin/env python
"""bench.py - read pairs/s "scored + assigned" and GCUPS of the hla-mapper hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): synthetic 30x WGS MHC-region reads, 2x150 bp,
2 000 000 pairs per GPU per step, against the synthetic IPD-IMGT/HLA-shaped database (seed 42,
24 loci + others, ~40 000 genomic alleles of 3-4 kb, ~148 Mb) in the reference's on-disk
layout.  One step = one pass of screen -> trim -> locus mask -> seed -> Myers prefilter ->
affine DP -> per-locus min -> cross-locus argmin over one batch.

  value     whole-job pairs/s with the batch already resident in HBM (hlm_run_device), device
            time from CUDA events on the library's own streams, max over ranks
  e2e       the same through the C-ABI call a user makes (hlm_run) with HOST buffers: staging
            into pinned memory, H2D, kernels and D2H of all results inside the timed region
  roofline  the dominant kernel of the step (live CUDA-event time on its launch stream) against
            its bound; `rooflines` carries every kernel of the step
  cpu_baseline  the oracle port (oracle/liborc.so, all host cores) on a bounded prefix

Multi-GPU (torchrun, one rank per GPU): pairs are sharded by rank, the database is replicated,
there is no data-path collective; torch.distributed only provides the barrier and the
max-over-ranks of the device times ("scaling": "weak").

--impl reference times the UNMODIFIED reference binary (oracle/_ref/hla-mapper-ref, shim-built
from /root/reference/src in the build container; its external `bwa` replaced by the stand-in
that calls the oracle aligner) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from hla_mapper_b200 import _abi, synth  # noqa: E402

DB_KW = dict(seed=42, n_loci=24, total_alleles=40000, n_others=2000)
READS_KW = dict(read_len=150, seed=1001, on_target=1.0)
METRIC = "read pairs/sec scored+assigned"
# integer instructions of the SASS inner loops (tools/sass_mix.py -> profiles/r1_sass_mix.txt)
DP_ALU_OPS_PER_CELL = 5.9     # 243 ALU-pipe instructions per 41-cell band row
DP_INT_OPS_PER_CELL = 10.3    # + 180 IMAD on the FMA pipe
MYERS_ALU_OPS_PER_ROW = 23.6  # 189 ALU-pipe instructions per 8 band rows (profiles/r1_sass_mix.txt)
WORKLOAD = "synthetic 30x WGS MHC-region reads, 2x150bp, 2M pairs/GPU vs 24 loci + others, 40k-allele DB seed 42"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def get_db(path, rank_is_builder, barrier):
    """the synthetic database, written once per box under `path` in the reference layout"""
    marker = os.path.join(path, ".complete")
    if rank_is_builder and not os.path.exists(marker):
        t = time.time()
        synth.make_db(path, **DB_KW)
        open(marker, "w").write("ok")
        log(f"[bench] database written in {time.time() - t:.1f}s -> {path}")
    barrier()
    # cheap in-memory description for the read simulator (no files rewritten)
    return synth.make_db(path, write=False, **DB_KW)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                 str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def peaks():
    """roofline denominators: MEASURED_PEAKS.json (driver-written) or the recipe's fallback"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- our arm
def run_b200(args):
    from hla_mapper_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, W, n = args.steps, args.warmup, args.pairs
    db = get_db(args.db, local == 0, barrier)
    t = time.time()
    # weak scaling: every rank draws its own BLOCK-aligned shard of the pair stream
    from hla_mapper_b200 import shard

    per_rank = -(-n // synth.BLOCK) * synth.BLOCK
    lo, _ = shard.shard_bounds(per_rank * world, world, synth.BLOCK)[rank]
    rb = synth.make_reads(db, n, first_pair=lo, **READS_KW)
    log(f"[bench r{rank}] {n} pairs generated in {time.time() - t:.1f}s")
    t = time.time()
    gdb = api.Database(args.db)
    p = _abi.default_params()
    if args.chunk:
        p.chunk_pairs = args.chunk
    m = api.Mapper(gdb, local, p)
    T = gdb.n_loci + 1
    log(f"[bench r{rank}] index built in {time.time() - t:.1f}s: "
        + str({k: gdb.stat(k) for k in ("alleles", "bases", "tiles", "kmers", "seed_entries")}))

    # ---- device-resident arm
    h = m.upload(rb)
    for _ in range(W):
        m.run_device(h)
    barrier()
    clk = ClockSampler(local)
    clk.start()
    stage = {}
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(K):
        m.run_device(h)  # returns after the last chunk's events completed (device synchronised)
        pr = m.profile()
        dev_ms += pr["ms_total"]
        for k, v in pr.items():
            stage[k] = stage.get(k, 0) + v
    wall_dev = time.perf_counter() - t0
    clocks = clk.stop()
    barrier()
    dev_ms = max_over_ranks(dev_ms)
    res_dev = m.fetch(h, n)
    m.free(h)

    # ---- per-kernel times: same work, one pipeline slot so that the CUDA events bracketing each
    # kernel on the launch stream are not stretched by the other slot's kernels (rank 0 only;
    # outside the timed region of `value`)
    kstage = {}
    if rank == 0:
        ps = _abi.default_params(single_slot=1)
        if args.chunk:
            ps.chunk_pairs = args.chunk
        ms_ = api.Mapper(gdb, local, ps)
        hs = ms_.upload(rb)
        ms_.run_device(hs)
        for _ in range(max(1, min(K, 2))):
            ms_.run_device(hs)
            for k, v in ms_.profile().items():
                kstage[k] = kstage.get(k, 0) + v
        for k in kstage:
            kstage[k] /= max(1, min(K, 2))
        ms_.free(hs)
        ms_.close()

    # ---- end to end through the public call with host buffers
    # what hla-mapper consumes downstream: screen flags / trims / locus mask, the (s1, s2) table
    # and the assignment; the per-alignment diagnostics are not requested
    buf = _abi.Buffers(n, T, diag=False)
    for _ in range(max(1, min(W, 2))):
        m.run(rb, buffers=buf)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(K):
        m.run(rb, buffers=buf)
        pr = m.profile()
        h2d += pr["h2d_bytes"]
        d2h += pr["d2h_bytes"]
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    for k in buf.FIELDS:  # the two arms must agree bit for bit
        if k[:-1] in ("aln", "nm", "lead", "trail"):
            continue
        assert np.array_equal(getattr(buf, k), getattr(res_dev, k)), k
    kept = int(((buf.flags & 2) != 0).sum())
    assigned = int((buf.target_mask != 0).sum())

    # ---- FASTQ text in, the reference's output files out (hlm_map_fastq), rank 0, one pass:
    # reported beside the metric, not part of it (text parsing is separate)
    files = None
    if rank == 0 and not args.no_files:
        import shutil
        import tempfile

        tmp = tempfile.mkdtemp(prefix="hlm_files_", dir=os.path.dirname(args.db.rstrip("/")))
        try:
            r1, r2 = os.path.join(tmp, "r1.fastq"), os.path.join(tmp, "r2.fastq")
            synth.write_fastq_fast(rb, r1, r2)
            os.makedirs(os.path.join(tmp, "out"))
            t0 = time.perf_counter()
            st = m.map_fastq(r1, r2, "B", os.path.join(tmp, "out"))
            dt = time.perf_counter() - t0
            files = {"value": round(n / dt, 1), "unit": "pairs/s",
                     "input_bytes": os.path.getsize(r1) + os.path.getsize(r2),
                     "what": "hlm_map_fastq: 2 plain FASTQ files -> selected / trim / addressing table / "
                             "per-gene FASTQ files, one reader thread per file",
                     **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    barrier()

    total_pairs = sum_over_ranks(float(n)) * K
    h2d_all, d2h_all = sum_over_ranks(float(h2d)), sum_over_ranks(float(d2h))  # all ranks
    value = total_pairs / (dev_ms * 1e-3)
    e2e = total_pairs / wall_e2e
    int_peak = api.int32_peak(local) if rank == 0 else None
    dp_ab = api.dp_layout_ab(local, 1 << 18, 150) if rank == 0 else None
    m.close()
    gdb.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- rooflines (rank 0's kernels; CUDA-event times of the serialised pass above)
    hbm_peak, hbm_src = peaks()
    per = lambda k: kstage[k]  # noqa: E731
    cells = kstage["n_dp_cells"]
    dp_ms, my_ms, sc_ms = per("ms_dp"), per("ms_myers"), per("ms_screen")
    gcups = cells / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else 0.0
    # INT32 roofline.  Every min/max (DPX VIADDMNMX / VIMNMX3) and logic instruction issues on the
    # ALU pipe only (64 lanes/clk/SM = the measured `viaddmnmx` rate); adds can also go to the FMA
    # pipe (IMAD).  Instruction counts per cell / per band row are read off the SASS inner loops
    # (profiles/r1_sass_mix.txt): the ALU pipe is the binding one for both kernels.
    alu_peak = int_peak["viaddmnmx"]            # G lane-op/s on the ALU pipe
    dp_alu, dp_all = DP_ALU_OPS_PER_CELL, DP_INT_OPS_PER_CELL
    dp_peak_gcups = min(alu_peak / dp_alu, int_peak["mix"] / dp_all)
    screen_bytes = kstage["bases_bytes"] + 12.0 * n
    screen_gbs = screen_bytes / (sc_ms * 1e-3) / 1e9 if sc_ms > 0 else 0.0
    myers_rows = kstage["n_myers_cols"]
    my_grows = myers_rows / (my_ms * 1e-3) / 1e9 if my_ms > 0 else 0.0
    my_peak = alu_peak / MYERS_ALU_OPS_PER_ROW
    rooflines = {
        "k_dp": {"bound": "int32", "achieved": round(gcups, 1), "peak": round(dp_peak_gcups, 1),
                 "unit": "GCUPS", "frac": round(gcups / dp_peak_gcups, 4),
                 # dram bytes read + written by one round-0 launch (65 536-pair chunk),
                 # profiles/r1i_kernels_full.txt; the kernel is ALU-bound
                 "traffic": 169649920,
                 "ms_per_step": round(dp_ms, 3), "cells_per_step": int(cells),
                 "int32_peak_glops": {k: round(v, 1) for k, v in int_peak.items()},
                 "alu_ops_per_cell": dp_alu, "int_ops_per_cell": dp_all,
                 "peak_source": "hlm_int32_peak micro-kernels measured in this run"},
        "k_myers": {"bound": "int32", "achieved": round(my_grows, 2), "peak": round(my_peak, 2),
                    "unit": "G band-rows/s", "frac": round(my_grows / my_peak, 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of one seed-index-phase launch
                    # (65 536-pair chunk), profiles/r1i_kernels_full.txt; the kernel is ALU-bound
                    "traffic": 574942720,
                    "ms_per_step": round(my_ms, 3), "rows_per_step": int(myers_rows),
                    "alu_ops_per_row": MYERS_ALU_OPS_PER_ROW},
        "k_screen": {"bound": "hbm", "achieved": round(screen_gbs, 1), "peak": hbm_peak,
                     "unit": "GB/s", "frac": round(screen_gbs / hbm_peak, 4),
                     # one launch (65 536 pairs, 40 MB algorithmic): 137 MB read + 23 MB written,
                     # profiles/r1i_kernels_full.txt - the excess is the k-mer table probes
                     "traffic": 159075840,
                     "ms_per_step": round(sc_ms, 3), "bytes_per_step": int(screen_bytes),
                     "peak_source": hbm_src,
                     "probes_per_s": round(n * 130.0 / (sc_ms * 1e-3), 0) if sc_ms > 0 else None,
                     "limiter": "130 random 8-byte probes per pair into the 134 MB k-mer table (L2 hit "
                                "55 % in profiles/r1i_kernels_full.txt), not the streaming of the reads; "
                                "the kernel is 1 % of the step"},
        "k_seed": {"ms_per_step": round(per("ms_seed"), 3),
                   "seed_candidates_per_step": int(kstage["n_seed_candidates"])},
        "k_expand": {"ms_per_step": round(per("ms_expand"), 3),
                     "candidates_per_step": int(kstage["n_candidates"])},
        "k_select": {"ms_per_step": round(per("ms_select"), 3)},
        "k_scores+k_assign": {"ms_per_step": round(per("ms_assign"), 3)},
    }
    times = {k: v.get("ms_per_step", 0) for k, v in rooflines.items()}
    dom = max(times, key=times.get)
    roof = dict(rooflines[dom])
    roof["kernel"] = dom
    roof["share_of_step"] = round(times[dom] / max(1e-9, kstage["ms_total"]), 3)

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": "pairs/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": round(dev_ms / K, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": n, "read_len": 150,
                   "db": DB_KW, "reads": READS_KW, "chunk_pairs": int(m.params.chunk_pairs) or 65536,
                   "l2": "inputs (1.2 GB/step) larger than the 126 MB L2; no flush needed",
                   "parallelism": f"pair-sharded x{world}, DB replicated, no collective"},
        "gcups": round(gcups, 1),
        "e2e": {"value": round(e2e, 1), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_all / K),
                "d2h_bytes_per_step": int(d2h_all / K)},
        "gpu_launches": int(stage["launches"]),
        "file_pipeline": files,
        "dp_layout_ab": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in dp_ab.items()},
        "clocks": clocks,
        "roofline": roof,
        "rooflines": rooflines,
        "work": {"kept_pairs": kept, "assigned_pairs": assigned,
                 "candidates_per_step": int(stage["n_candidates"] / K),
                 "dp_alignments_per_step": int(stage["n_dp"] / K),
                 "serialised_ms_per_step": round(kstage["ms_total"], 3),
                 "wall_ms_per_step_device_arm": round(wall_dev * 1e3 / K, 3)},
    }
    if not args.no_cpu and world >= 1:
        out["cpu_baseline"] = cpu_baseline(args, db, rb)
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args, db, rb):
    """the oracle port on the host cores, on a bounded prefix of the same batch"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc

    cores = os.cpu_count() or 1
    odb = orc.OracleDB(args.db)
    p = _abi.default_params()
    n = min(rb.n_pairs, args.cpu_pairs)
    sub = rb.slice(0, n)
    orc.run(odb, p, rb.slice(0, min(64, n)), mode=1, nthreads=cores)  # touch the index
    t = time.perf_counter()
    _, cells = orc.run(odb, p, sub, mode=1, nthreads=cores)
    dt = time.perf_counter() - t
    odb.close()
    return {"value": round(n / dt, 1), "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"first {n} pairs of the same batch, oracle/liborc.so fast mode (OpenMP), "
                      f"{dt:.1f}s wall",
            "gcups": round(cells / dt / 1e9, 3)}


# ------------------------------------------------------------------------------- reference arm
def run_reference_arm(args):
    """the unmodified reference binary on the host cores (rank 0 only)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import run_reference as rr

    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    db = get_db(args.db, True, lambda: None)
    n = args.ref_pairs
    rb = synth.make_reads(db, n, **READS_KW)
    tmp = os.path.join(os.path.dirname(args.db.rstrip("/")), "ref_arm")
    os.makedirs(tmp, exist_ok=True)
    r1, r2 = os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")
    synth.write_fastq(rb, r1, r2)
    K, W = args.steps, args.warmup
    kind = "reference" if rr.available() else "port"
    times = []
    for it in range(W + K):
        t = time.perf_counter()
        if kind == "reference":
            out = rr.run_reference(args.db, os.path.join(tmp, "out"), "B", r1=r1, r2=r2,
                                   threads=cores, debug=False)
            assert os.path.exists(out + "B_addressing_table.txt")
        else:
            import oracle as orc

            orc.run(orc.OracleDB(args.db), _abi.default_params(), rb, mode=1, nthreads=cores)
        dt = time.perf_counter() - t
        log(f"[bench reference] step {it}: {n} pairs in {dt:.1f}s")
        if it >= W:
            times.append(dt)
    tot = sum(times)
    value = n * K / tot
    sample = (f"{n} pairs of the same workload per step through `hla-mapper dna` (screen, trim, "
              f"sort, bwa stand-in scoring, compare), threads={cores}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "pairs/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(tot / K * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n, "db": DB_KW, "reads": READS_KW},
        "cpu_baseline": {"value": round(value, 2), "unit": "pairs/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=2_000_000, help="pairs per GPU per step")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--db", default=os.environ.get("HLM_BENCH_DB", "/tmp/hlm_bench/db_seed42_40k"))
    ap.add_argument("--cpu-pairs", type=int, default=4000)
    ap.add_argument("--ref-pairs", type=int, default=6000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-files", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # plain `python bench.py --gpus N`: launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
