"""The C-ABI library loads, exports every symbol include/hlamap.h declares, mirrors the
reference's defaults and fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hlamap.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hlm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hla_mapper_b200 import api

    names = _declared()
    assert len(names) >= 20
    assert sorted(api.SYMBOLS) == names
    L = api.lib()
    for n in names:
        assert getattr(L, n) is not None
    assert L.hlm_abi_version() == 5  # 3: hlm_params.score_cap; 4: hlm_score_alleles; 5: hlm_star_sam_read, multi file entry points


def test_struct_layouts_match_the_header():
    from hla_mapper_b200 import _abi, api

    p = _abi.HlmParams()
    api.lib().hlm_params_default(C.byref(p), 0)
    assert p.struct_size == C.sizeof(_abi.HlmParams)
    d = _abi.default_params(0)
    for k, _ in _abi.HlmParams._fields_:
        assert getattr(p, k) == getattr(d, k), k
    # main.cpp:73-86 defaults
    assert (p.minhit, p.mm_max, p.tolerance) == (3, 20, 0.05)
    assert p.mtrim_error == float(__import__("numpy").float32(0.08))
    api.lib().hlm_params_default(C.byref(p), 1)
    assert p.tolerance == 0.08 and p.rna == 1  # map_rna.cpp:116
    # hlm_batch: 12 pointer-sized fields; hlm_rna_block: 8 bytes, (gene, mate, start, len, reserved)
    assert C.sizeof(_abi.HlmBatch) == 12 * 8
    assert _abi.RNA_BLOCK.itemsize == 8 and _abi.RNA_BLOCK.fields["start"][1] == 2


def test_db_open_error_messages(tmp_path):
    from hla_mapper_b200 import api

    with pytest.raises(api.HlmError, match="Error accessing database"):
        api.Database(str(tmp_path / "nope"))
    (tmp_path / "db_dna.info").write_text("hla-mapper:3.0\nGENE:HLA-A:chr6:1:2:3:4:HLA\n")
    with pytest.raises(api.HlmError, match="not compatible"):
        api.Database(str(tmp_path))


def test_db_open_refuses_more_loci_than_the_masks_hold(tmp_path):
    from hla_mapper_b200 import api

    (tmp_path / "db_dna.info").write_text(
        "hla-mapper:4.3+\n" + "".join(f"GENE:G{i}:chr6:1:2:3:4:HLA\n" for i in range(32)) + "SIZE:50\n")
    with pytest.raises(api.HlmError, match="more than 31 GENE"):
        api.Database(str(tmp_path))


def test_db_open_parses_reference_layout(small_db):
    from hla_mapper_b200 import api

    db = api.Database(small_db.path)
    assert db.loci == small_db.loci + ["others"]
    assert db.v_size == 50
    assert db.stat("alleles") == small_db.n_alleles + len(small_db.others)
    assert db.stat("tiles") <= db.stat("tiles_raw")
    db.close()


def _no_gpu():
    from hla_mapper_b200 import api

    return api.lib().hlm_int32_peak(0, None, None, None, None) != 0


def test_no_cpu_fallback(small_db):
    """without a CUDA device the product refuses to run: there is no CPU path"""
    from hla_mapper_b200 import api

    if not _no_gpu():
        pytest.skip("a GPU is present")
    db = api.Database(small_db.path)
    with pytest.raises(api.HlmError, match="needs a CUDA device"):
        api.Mapper(db, 0)
    db.close()


def test_product_never_imports_the_oracle():
    pk = os.path.join(ROOT, "hla_mapper_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle', s), f
                assert "liborc" not in s and "import oracle" not in s and "orc_" not in s, f
