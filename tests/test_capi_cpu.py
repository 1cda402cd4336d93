"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
fails loudly without a GPU, and its host-side layout construction is exact.  No compute."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURES, ROOT, load_dataset, load_fixture
from oracle import cora_oracle as co


def test_library_exports_every_declared_symbol(lib):
    from cora_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "cora_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cora_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    for s in declared:
        assert hasattr(lib, s), s
    assert sorted(capi.SYMBOLS) == declared
    assert lib.cora_b200_version() >= 100


def test_no_cpu_fallback(lib):
    from cora_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    g, p = load_fixture("small_ra_slam_problem")
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    with pytest.raises(capi.CoraB200Error) as e:
        capi.Handle(p.d, p.n, p.m, p.n + p.l, p.Q)
    assert e.value.code == capi.ECUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """Nothing under cora_b200/ may import, include, link or execute anything under oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include[^\n]*oracle|oracle/|cora_oracle", re.M)
    for root, _, files in os.walk(os.path.join(ROOT, "cora_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert not pat.search(txt), os.path.join(root, f)


def _roundtrip(p, tile_rows=None):
    from cora_b200 import capi
    if tile_rows:
        os.environ["CORA_B200_TILE_ROWS"] = str(tile_rows)
    try:
        out, stats = capi.layout_roundtrip(p.d, p.n, p.m, p.n + p.l, p.Q)
        out_s, _ = capi.layout_roundtrip(p.d, p.n, p.m, p.n + p.l, p.Q, strips=True)
    finally:
        os.environ.pop("CORA_B200_TILE_ROWS", None)
    ref = sp.csr_matrix(p.Q); ref.eliminate_zeros(); ref.sort_indices()
    for o in (out, out_s):  # tile layout, and the strip records of the streaming kernels built from it
        o.eliminate_zeros(); o.sort_indices()
        assert np.array_equal(ref.indptr, o.indptr)
        assert np.array_equal(ref.indices, o.indices)      # integer indexing bit-exact
        assert np.array_equal(ref.data, o.data)            # values are moved, never recomputed
    return stats


@pytest.mark.parametrize("name", FIXTURES)
def test_layout_roundtrip_fixtures(lib, name):
    g, p = load_fixture(name)
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    for tr in (12, 24, 192):
        _roundtrip(p, tr)


@pytest.mark.parametrize("name", ["plaza2", "single_drone"])
def test_layout_roundtrip_datasets(lib, name):
    p = load_dataset(name)
    p.update_problem_data()
    stats = _roundtrip(p)
    # odometry chain: block-tridiagonal -> 3 block slots; landmark rows are hub groups
    assert stats["max_slots"] == 3
    assert stats["num_hub_groups"] == p.l
    assert stats["nnz_block"] + stats["nnz_spill"] + stats["nnz_hub"] + p.l + p.m >= p.Q.nnz


def test_layout_roundtrip_loop_closures_and_overflow(lib):
    """A star of loop closures around pose 0 exceeds the 8 block slots: the surplus blocks
    must land in the CSR spill, and a pose with > 64 spill entries becomes a hub group."""
    from synth import make_synthetic
    p = make_synthetic(n=40, l=3, m=25, seed=5, loop_closures=[(0, j) for j in range(2, 40, 2)])
    p.update_problem_data()
    stats = _roundtrip(p, 24)
    assert stats["max_slots"] == 8
    assert stats["nnz_spill"] + stats["nnz_hub"] > 0


def test_layout_rejects_bad_input(lib):
    from cora_b200 import capi
    g, p = load_fixture("small_ra_slam_problem")
    p.preconditioner = co.JACOBI
    p.update_problem_data()
    with pytest.raises(capi.InvalidArgument):
        capi.layout_roundtrip(4, p.n, p.m, p.n + p.l, p.Q)
