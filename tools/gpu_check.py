"""First-contact GPU check: oracle vs libhlamap on small seeded inputs, then a timing pass.
Usage (on the GPU box): python tools/gpu_check.py [n_pairs_small] [n_pairs_timing]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np

from hla_mapper_b200 import _abi, api, synth
import oracle as orc


def compare(a, b, tag):
    bad = 0
    for k in a.FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        if not np.array_equal(x, y):
            idx = np.argwhere(x != y)
            print(f"[{tag}] MISMATCH {k}: {len(idx)} entries; first {idx[:5].tolist()} "
                  f"oracle={x[tuple(idx[0])]} gpu={y[tuple(idx[0])]}")
            bad += 1
    print(f"[{tag}] fields compared={len(a.FIELDS)} mismatching={bad}")
    return bad


def main():
    n_small = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    tmp = os.environ.get("HLM_TMP", "/tmp/hlm_check")
    os.makedirs(tmp, exist_ok=True)
    total_bad = 0
    print("int32 peak (G lane-ops/s):", api.int32_peak(0))
    # --- small DB, exhaustive oracle
    db = synth.make_db(os.path.join(tmp, "db_small"), seed=7, n_loci=6, total_alleles=120,
                       len_range=(600, 800), n_others=20)
    rb = synth.make_reads(db, n_small, read_len=150, seed=11, on_target=0.7, frag_mu=300, frag_sd=40)
    odb = orc.OracleDB(db.path)
    gdb = api.Database(db.path)
    print("db stats", {k: gdb.stat(k) for k in ("alleles", "bases", "tiles", "tiles_raw", "kmers", "seed_entries")})
    for variant in (0, 1):
        for clip in (0, 1):
            p = _abi.default_params(screen_variant=variant, clip_mode=clip)
            ob, _ = orc.run(odb, p, rb, mode=0)
            m = api.Mapper(gdb, 0, p)
            gb = m.run(rb)
            total_bad += compare(ob, gb, f"small dna variant={variant} clip={clip}")
            if variant == 0 and clip == 0:
                # stage-wise entry points must agree with the fused run
                sb = m.screen_trim(rb)
                m.score(rb, sb)
                m.assign(rb, sb)
                total_bad += compare(ob, sb, "stage-wise")
                h = m.upload(rb)
                m.run_device(h)
                db_ = m.fetch(h, rb.n_pairs)
                m.free(h)
                total_bad += compare(ob, db_, "device-resident")
                print("profile", m.profile())
            m.close()
    # ragged reads, single-end
    rr = synth.make_reads(db, 2000, read_len=120, seed=12, on_target=0.8, frag_mu=260, frag_sd=30, ragged=True)
    p = _abi.default_params()
    ob, _ = orc.run(odb, p, rr, mode=1)
    m = api.Mapper(gdb, 0, p, chunk_pairs=512)
    total_bad += compare(ob, m.run(rr), "ragged chunk=512")
    m.close()
    p = _abi.default_params(paired=0)
    ob, _ = orc.run(odb, p, rr, mode=1)
    m = api.Mapper(gdb, 0, p)
    total_bad += compare(ob, m.run(rr), "single-end")
    m.close()
    # --- medium DB timing + fast-oracle parity on a prefix
    t = time.time()
    db2 = synth.make_db(os.path.join(tmp, "db_med"), seed=42, n_loci=24, total_alleles=4000, n_others=200)
    print("db_med gen", round(time.time() - t, 1), "s alleles", db2.n_alleles)
    t = time.time()
    gdb2 = api.Database(db2.path)
    print("db_med open", round(time.time() - t, 1), "s", {k: gdb2.stat(k) for k in ("alleles", "bases", "tiles", "tiles_raw", "kmers", "seed_entries", "kt_slots")})
    rb2 = synth.make_reads(db2, n_time, read_len=150, seed=1001, on_target=1.0)
    p = _abi.default_params()
    m = api.Mapper(gdb2, 0, p)
    for rep in range(2):
        t = time.time()
        gb2 = m.run(rb2)
        dt = time.time() - t
        pr = m.profile()
        print(f"run {rep}: {n_time} pairs in {dt:.3f}s = {n_time / dt:.0f} pairs/s wall; profile={json.dumps(pr)}")
    h = m.upload(rb2)
    for rep in range(2):
        t = time.time()
        m.run_device(h)
        dt = time.time() - t
        pr = m.profile()
        print(f"run_device {rep}: {dt:.3f}s = {n_time / dt:.0f} pairs/s; ms_total={pr['ms_total']:.1f} screen={pr['ms_screen']:.1f} seed={pr['ms_seed']:.1f} myers={pr['ms_myers']:.1f} sel={pr['ms_select']:.1f} dp={pr['ms_dp']:.1f} assign={pr['ms_assign']:.1f} cand={pr['n_candidates']} dp={pr['n_dp']} kept={pr['n_kept']}")
    m.free(h)
    odb2 = orc.OracleDB(db2.path)
    sub = rb2.slice(0, 300)
    t = time.time()
    ob2, cells = orc.run(odb2, p, sub, mode=1)
    print("oracle fast 300 pairs", round(time.time() - t, 2), "s")
    sb = _abi.Buffers(300, gdb2.n_loci + 1)
    for k in sb.FIELDS:
        getattr(sb, k)[...] = getattr(gb2, k)[:300]
    total_bad += compare(ob2, sb, "medium prefix")
    kept = (gb2.flags & 2) != 0
    tl = rb2.truth_locus
    hit = ((gb2.target_mask >> np.maximum(tl, 0).astype(np.uint32)) & 1).astype(bool) & (tl >= 0)
    print("kept", kept.sum(), "true locus among targets", int(hit.sum()), "unique-target", int((np.array([bin(x).count('1') for x in gb2.target_mask[:20000]]) == 1).sum()), "/20000")
    m.close()
    print("TOTAL_BAD", total_bad)
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
