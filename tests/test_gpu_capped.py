"""hlm_params.score_cap = 1 (tolerance-capped scoring, dna): everything the compare step and the
writers consume must be identical to the exact mode, i.e. to the oracle:

  * screen outputs, target_mask, visible_mask, min_sum: equal;
  * every score of a VISIBLE target (the ones the addressing table prints): equal;
  * a capped score is either the exact score or 0 ("not computed"), never anything else;
  * the golden files the unmodified reference wrote: byte-identical (tests/golden/*).

The capped mode must also do less work (fewer DP alignments), or it has no reason to exist."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
import golden_check  # noqa: E402
from hla_mapper_b200 import _abi, synth  # noqa: E402

pytestmark = pytest.mark.gpu

SCREEN = ("flags", "trim_lo1", "trim_len1", "trim_lo2", "trim_len2", "locus_mask")
ASSIGN = ("target_mask", "visible_mask", "min_sum")


def check_capped_against_exact(want, got, T, paired=True):
    """want: exact results (oracle or exact GPU mode); got: capped GPU results"""
    for k in SCREEN + ASSIGN:
        x, y = getattr(want, k), getattr(got, k)
        assert np.array_equal(x, y), f"{k}: {int((x != y).sum())} differences"
    vis = ((want.visible_mask[:, None] >> np.arange(T)[None, :]) & 1).astype(bool)
    # the others column is printed from the back-filled pair (map_dna.cpp:903-915)
    loci = vis.copy()
    loci[:, T - 1] = False
    for k in ("s1", "s2") if paired else ("s1",):
        x, y = getattr(want, k), getattr(got, k)
        assert np.array_equal(x[loci], y[loci]), f"visible {k} differs"
        nz = y != 0
        assert np.array_equal(x[nz], y[nz]), f"a capped {k} is neither the exact score nor 0"
    ov = vis[:, T - 1]
    assert np.array_equal(want.others_s1[ov], got.others_s1[ov])
    if paired:
        assert np.array_equal(want.others_s2[ov], got.others_s2[ov])


@pytest.mark.parametrize("clip", [0, 1, 2])
@pytest.mark.parametrize("variant", [0, 1])
def test_capped_equals_exact_where_it_matters(orc, small_db, small_reads, clip, variant):
    from hla_mapper_b200 import api

    gdb = api.Database(small_db.path)
    p = _abi.default_params(screen_variant=variant, clip_mode=clip)
    want, _ = orc.run(orc.OracleDB(small_db.path), p, small_reads, mode=0)
    m = api.Mapper(gdb, 0, p, score_cap=1)
    got = m.run(small_reads)
    n_dp_capped = m.profile()["n_dp"]
    m.close()
    check_capped_against_exact(want, got, gdb.n_loci + 1)
    m = api.Mapper(gdb, 0, p)
    m.run(small_reads)
    n_dp_exact = m.profile()["n_dp"]
    m.close()
    gdb.close()
    assert n_dp_capped < n_dp_exact


@pytest.mark.parametrize("seed", range(6))
def test_capped_fuzz(orc, tmp_path, seed):
    """random database shape x read shape x tolerance x chunk size; tie-heavy paralogs included"""
    from hla_mapper_b200 import api

    rng = np.random.default_rng(900 + seed)
    db = synth.make_db(str(tmp_path / "db"), seed=300 + seed, n_loci=int(rng.integers(2, 9)),
                       total_alleles=int(rng.integers(20, 200)), len_range=(500, 900),
                       n_others=int(rng.integers(1, 30)),
                       paralog_div=(0.01, 0.06) if seed % 2 else (0.08, 0.15))
    L = int(rng.choice([75, 100, 150]))
    rb = synth.make_reads(db, 900, read_len=L, seed=50 + seed, on_target=0.85, frag_mu=260, frag_sd=30,
                          ragged=bool(seed % 2), lowq_tail=0.3)
    paired = seed != 4
    p = _abi.default_params(paired=int(paired), tolerance=float(rng.choice([0.02, 0.05, 0.1])),
                            mm_max=int(rng.choice([20, 12, 6])))
    want, _ = orc.run(orc.OracleDB(db.path), p, rb, mode=1)
    gdb = api.Database(db.path)
    m = api.Mapper(gdb, 0, p, score_cap=1, chunk_pairs=int(rng.choice([64, 333, 65536])))
    got = m.run(rb)
    m.close()
    check_capped_against_exact(want, got, gdb.n_loci + 1, paired)
    gdb.close()


@pytest.mark.parametrize("name", [n for n, c in cases.CASES.items() if c[2] in ("paired", "single")])
def test_capped_file_pipeline_writes_the_reference_files(name, tmp_path):
    """the files of the unmodified reference, byte for byte, from the capped mode"""
    from hla_mapper_b200 import api

    db, rb, p, mode = cases.inputs(name, str(tmp_path))
    r1, r2 = cases.write_inputs(name, rb, mode, str(tmp_path))
    gold = golden_check.load(name)
    gdb = api.Database(db.path)
    m = api.Mapper(gdb, 0, p, score_cap=1)
    out = tmp_path / "out"
    out.mkdir()
    st = m.map_fastq(r1, r2, "G", out, batch_pairs=257)
    m.close()
    gdb.close()
    assert st["n_kept"] == len(gold["trim1"])
    for fname, digest in gold["files"].items():
        assert cases.canonical_digest(str(out / fname)) == digest, fname
