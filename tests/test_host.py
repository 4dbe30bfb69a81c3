"""CPU tests of the host-side logic and of the C-ABI library's surface (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_range_len_matches_julia_ranges():
    from tonga_b200.structs import StepRangeLen
    r = StepRangeLen(-79.477, 20, 1060.65)
    assert len(r) == 58 and abs(r.max() - 1060.523) < 1e-9 and r.min() == -79.477
    z = StepRangeLen(0.0, 20, 660.0)
    assert len(z) == 34 and z.max() == 660.0 and np.array_equal(z.vec(), np.arange(34) * 20.0)


def test_parameters_defaults_match_define_TDstructure():
    from tonga_b200.structs import define_TDstructrure
    p = define_TDstructrure()
    assert (p.debug_prior, p.plot_voronoi, p.add_yVec) == (0, 0, 1)
    assert (p.sig, p.zeta_scale, p.max_cells, p.min_cells, p.max_sig, p.interp_style, p.enforce_discon, p.prior) == (10, 50, 100, 5, 0.1, 1, 0, 1)
    assert (p.n_chains, p.n_iter, p.burn_in, p.keep_each, p.print_each) == (2, 1e3, 5e2, 1e1, 1e2)
    assert (p.max_depth, p.min_depth, p.ZnodeSpacing, p.buffer, p.XYnodeSpacing) == (660, 0, 20, 100, 20)
    assert p.zSlice == [50, 300, 500] and p.ySlice == [700, 800]


def test_pack_models_and_params(tonga):
    from tonga_b200.api import make_params, pack_models
    from tonga_b200.structs import Model
    ds, p = tonga
    a = Model(2.0, np.array([1.0, 2.0]), np.array([3.0, 4.0]), np.array([5.0, 6.0]), np.array([7.0, 8.0]))
    K, cells = pack_models([a, (np.zeros(3), np.ones(3), np.ones(3), np.ones(3))], Kcap=8)
    assert list(K) == [2, 3] and cells.shape == (2, 4, 8) and cells[0, 3, 1] == 8.0 and cells[0, 0, 2] == 0.0
    tp = make_params(p, ds)
    assert tp.xmin == ds.xVec.min() and tp.zmax == 660.0 and tp.max_cells == 100 and tp.n_actions == 4 and tp.interp_style == 1


def test_ak135_slowness_branches():
    from tonga_b200.data import ak135_slowness
    tab = np.array([[0.0, 5.0, 0], [10.0, 5.0, 0], [10.0, 8.0, 0], [20.0, 10.0, 0]])
    u = ak135_slowness(np.array([0.0, 5.0, 10.0, 15.0, 20.0, 25.0, np.nan]), tab)
    assert np.allclose(u[:6], [1 / 5, 1 / 5, 1 / 8, 1 / 9, 1 / 10, 1 / 10]) and np.isnan(u[6])


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tonga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tonga_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    """The in-tree .so loads without a GPU and exports exactly what include/tonga_b200.h declares."""
    from tonga_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    declared = _header_symbols()
    assert declared and set(declared) == set(_lib.SYMBOLS), (set(declared) ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tonga_version() == 100


def test_no_cpu_fallback_without_device(tonga):
    """Without a CUDA device every compute entry point fails loudly (there is no CPU path in the product)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tonga_b200 import _lib
    from tonga_b200.api import Context
    ds, p = tonga
    with pytest.raises(_lib.TongaError) as e:
        Context(ds, p)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, link or execute it."""
    pkg = os.path.join(ROOT, "mcmc-in-tonga_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|libtonga_oracle|oracle_np|oracle/|tonga_oracle")
    for dp_, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h", "Makefile")):
                assert not pat.search(open(os.path.join(dp_, f)).read()), f"{f} references the oracle"


def test_model_hist_export_round_trip(tmp_path):
    """checkpoint.export_model_hist / import_model_hist (the layout julia/model_hist_to_jld.jl reads), incl. the reference's
    action/accept aliasing."""
    from tonga_b200 import checkpoint as ckpt
    rng = np.random.default_rng(0)
    n, H, KC, R = 3, 4, 8, 5
    hist = dict(n_hist=np.array([4, 2, 6], np.int32), K=rng.integers(1, KC + 1, (n, H)).astype(np.int32), cells=rng.normal(size=(n, H, 4, KC)),
                phi=rng.normal(size=(n, H)), ptS=rng.normal(size=(n, H, R)), iter=np.zeros((n, H), np.int64),
                action=rng.integers(1, 5, (n, H)).astype(np.int32), accept=rng.integers(0, 2, (n, H)).astype(np.int32),
                next_action=rng.integers(0, 5, (n, H)).astype(np.int32))
    path = str(tmp_path / "mh.bin")
    assert ckpt.export_model_hist(path, hist, likelihood=-2.0) == 4 + 2 + 4  # n_hist beyond the capacity is clipped
    back = ckpt.import_model_hist(path)
    assert [len(b) for b in back] == [4, 2, 4]
    for c in range(n):
        for j, m in enumerate(back[c]):
            k = hist["K"][c, j]
            assert m["K"] == k and m["action"] == hist["action"][c, j] and m["accept"] == hist["accept"][c, j] and m["likelihood"] == -2.0
            assert np.array_equal(m["cells"], hist["cells"][c, j, :, :k]) and np.array_equal(m["ptS"], hist["ptS"][c, j])
    ckpt.export_model_hist(path, hist, likelihood=0.0, reference_aliasing=True)
    back = ckpt.import_model_hist(path)
    for c in range(n):
        for j, m in enumerate(back[c]):
            nxt = hist["next_action"][c, j]
            assert m["accept"] == 0 and m["action"] == (nxt if nxt > 0 else hist["action"][c, j])
