"""N>1 host logic on CPU: two gloo ranks shard a batch, each computes its shard (the oracle stands
in for the GPU here - it is only the checker of the sharding/gather logic), rank 0 gathers and the
result must equal the single-process result; device times reduce with MAX."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    from hla_mapper_b200 import shard

    for n in (0, 1, 7, 65536, 65537, 2_000_000):
        for world in (1, 2, 3, 4, 8):
            for block in (1, 64, 65536):
                b = shard.shard_bounds(n, world, block)
                assert b[0][0] == 0 and b[-1][1] == n
                for (lo, hi), (lo2, _) in zip(b, b[1:]):
                    assert hi == lo2 and lo <= hi
                for lo, hi in b:
                    assert lo % block == 0 or lo == n


def _worker(rank, world, port, db_path, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist

    import oracle as orc
    from hla_mapper_b200 import _abi, shard, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.make_db(db_path, seed=7, n_loci=6, total_alleles=120, len_range=(600, 800), n_others=20,
                       write=False)
    rb = synth.make_reads(db, 900, read_len=150, seed=11, on_target=0.7, frag_mu=300, frag_sd=40)
    bounds = shard.shard_bounds(rb.n_pairs, world, 64)
    lo, hi = bounds[rank]
    odb = orc.OracleDB(db_path)
    buf, _ = orc.run(odb, _abi.default_params(), rb.slice(lo, hi), mode=1, nthreads=2)
    got = shard.gather_to_rank0(dist, buf.as_dict(), bounds, rank, world)
    t = shard.max_over_ranks(dist, float(10 + rank))
    assert t == float(10 + world - 1)
    if rank == 0:
        np.savez(out_path, **got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_process(orc, small_db, tmp_path):
    import torch.multiprocessing as mp

    from hla_mapper_b200 import _abi, synth

    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 400)
    mp.start_processes(_worker, args=(2, port, small_db.path, out), nprocs=2, join=True,
                       start_method="spawn")
    got = np.load(out)
    rb = synth.make_reads(small_db, 900, read_len=150, seed=11, on_target=0.7, frag_mu=300, frag_sd=40)
    want, _ = orc.run(orc.OracleDB(small_db.path), _abi.default_params(), rb, mode=1)
    for k in want.FIELDS:
        assert np.array_equal(getattr(want, k), got[k]), k
