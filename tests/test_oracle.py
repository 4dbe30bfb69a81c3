"""CPU tests of the oracle itself (no GPU): the C restatement against its independent NumPy twin, the KATs of
v_nearest / Interpolation, and what the reference's only shipped output (model.jld) pins.

PARITY PARTLY PINNED: the reference cannot run here and has no golden vectors for the forward model (owners, t*), so those are
pinned by the twin + KATs only.  The misfit phi and the constant "likelihood" (MCsub.jl:169-182) ARE pinned by
reference-produced values: model.jld stores (ptS, tS, phi, likelihood) of 100 models, and the oracle reproduces every phi and
the likelihood to 1e-13 (test_misfit_and_likelihood_pinned_by_reference_output)."""
import os

import numpy as np
import pytest

from conftest import box_of, random_model, random_ragged

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _orc(ds, p, **kw):
    import oracle as O
    return O.make_params(box_of(ds), sig=p.sig, zeta_scale=p.zeta_scale, max_sig=p.max_sig, n_iter=p.n_iter, burn_in=p.burn_in,
                         keep_each=p.keep_each, min_cells=p.min_cells, max_cells=p.max_cells, prior=p.prior,
                         debug_prior=p.debug_prior, **kw), O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)


def test_fixture_fingerprints(tonga):
    """SURVEY.md 5.9 fingerprints of the shipped 381-ray set."""
    ds, p = tonga
    assert ds.rayX.shape == (131, 381) and ds.rayL.shape == (130, 381)
    valid = ~np.isnan(ds.rayX)
    assert valid.sum() == 16845 and (~np.isnan(ds.rayL)).sum() == 16464
    assert abs(np.nansum(ds.rayX) - 11280783.739) < 1e-3 and abs(np.nansum(ds.rayZ) - 4252266.1811) < 1e-3
    assert abs(ds.tS.sum() - 181.152908) < 1e-9 and abs(ds.allSig.sum() - 100.524106) < 1e-9
    # the authors' own frame survives in a comment: `yVec => -164.4:20:495.6` (plot_distribution.jl:40)
    assert abs(ds.yVec.min() - (-164.4623)) < 1e-9 and abs(ds.yVec.max() - 495.5377) < 1e-9
    assert len(ds.xVec) == 58 and len(ds.yVec) == 34 and len(ds.zVec) == 34
    assert (np.isnan(ds.rayL) == np.isnan(ds.rayU)).all()
    assert np.nanmin(ds.rayU) > 0.08 and np.nanmax(ds.rayU) < 0.7  # 1/Vp of ak135: 1/10.3 .. 1/1.45 s/km


def test_ray_lengths_c_vs_numpy(tonga):
    import oracle as O
    ds, _ = tonga
    rl, ru = O.ray_lengths(ds.rayX, ds.rayY, ds.rayZ, ds.U)
    assert np.array_equal(rl, ds.rayL, equal_nan=True) and np.array_equal(ru, ds.rayU, equal_nan=True)


@pytest.mark.parametrize("K", [1, 5, 27, 100])
def test_c_oracle_equals_numpy_twin_tonga(tonga, K):
    import oracle as O
    import oracle_np as N
    ds, p = tonga
    op, od = _orc(ds, p)
    rng = np.random.default_rng(K)
    mdl = random_model(rng, K, box_of(ds))
    a = O.evaluate(op, od, *mdl, want_owners=True)
    b = N.evaluate(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig, *mdl)
    assert np.array_equal(a["owners"], b["owners"])
    assert np.allclose(a["ptS"], b["ptS"], rtol=1e-12, atol=0) and abs(a["phi"] - b["phi"]) <= 1e-12 * abs(b["phi"])
    assert abs(a["likelihood"] - b["likelihood"]) <= 1e-12 * abs(b["likelihood"])
    assert a["valid"] == 1


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_c_oracle_equals_numpy_twin_ragged(seed):
    import oracle as O
    import oracle_np as N
    ds, p = random_ragged(seed)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(seed)
    for K in (1, 2, 9, 40):
        mdl = random_model(rng, K, box_of(ds))
        a = O.evaluate(op, od, *mdl, noise=1.0 + 0.25 * seed, want_owners=True)
        b = N.evaluate(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig, *mdl, noise=1.0 + 0.25 * seed)
        assert np.array_equal(a["owners"], b["owners"])
        assert np.allclose(a["ptS"], b["ptS"], rtol=1e-12, atol=0) and abs(a["phi"] - b["phi"]) <= 1e-12 * abs(b["phi"])


def test_v_nearest_kats():
    """MCsub.jl:247-263: strict <, first index wins ties, 1e9 start value, v = 0.0 when nothing is closer."""
    import oracle as O
    import oracle_np as N
    mx, my, mz, mv = np.array([9.0, 1.0, -1.0]), np.zeros(3), np.zeros(3), np.array([7.0, 11.0, 13.0])
    for f in (O.v_nearest, N.v_nearest):
        assert f(0.0, 0.0, 0.0, mx, my, mz, mv) == (11.0, 1)          # exact tie -> lowest index
        assert f(0.0, 0.0, 0.0, mx[::-1].copy(), my, mz, mv) == (7.0, 0)
        assert f(0.0, 0.0, 0.0, np.array([4e4]), np.zeros(1), np.zeros(1), np.array([5.0])) == (0.0, -1)  # d >= 1e9
        assert f(0.0, 0.0, 0.0, np.array([31622.0]), np.zeros(1), np.zeros(1), np.array([5.0])) == (5.0, 0)  # d < 1e9
        assert f(1.0, 2.0, 3.0, np.zeros(1), np.zeros(1), np.zeros(1), np.array([42.0])) == (42.0, 0)
        assert f(0.0, 0.0, 0.0, np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0)) == (0.0, -1)
    # NaN trim + slice broadcast (MCsub.jl:312-322)
    X = np.array([0.0, 1.0, 2.0, np.nan, 5.0])
    for f in (O.interpolation, N.interpolation):
        z, i = f(mx, my, mz, mv, X, [0.0], [0.0])
        assert len(z) == 3 and list(i) == [1, 1, 1]
        z, i = f(mx, my, mz, mv, np.array([np.nan, 1.0]), [0.0], [0.0])
        assert len(z) == 0


def test_evaluate_edge_rays():
    """Rays with 1 point (no segment: t* = 0), a full column with no NaN, and a point/segment count mismatch."""
    import oracle as O
    ds, p = random_ragged(9, R=6, m=5)
    ds.rayX[1:, 2] = np.nan; ds.rayY[1:, 2] = np.nan; ds.rayZ[1:, 2] = np.nan  # 1-point ray
    ds.rayL[:, 2] = np.nan; ds.rayU[:, 2] = np.nan
    op, od = _orc(ds, p)
    r = O.evaluate(op, od, *random_model(np.random.default_rng(0), 4, box_of(ds)), want_owners=True)
    assert r["valid"] == 1 and r["ptS"][2] == 0.0 and r["owners"][0, 2] >= 0 and (r["owners"][1:, 2] == -1).all()
    assert not np.isnan(ds.rayX[:, 0]).any() and r["ptS"][0] > 0
    bad = ds.rayL.copy()
    bad[0, 3] = np.nan  # Julia: DimensionMismatch in the broadcast at MCsub.jl:159
    od2 = O.Data(ds.rayX, ds.rayY, ds.rayZ, bad, ds.rayU, ds.tS, ds.allSig)
    assert O.evaluate(op, od2, *random_model(np.random.default_rng(0), 4, box_of(ds)))["valid"] == -2


def test_debug_prior_short_circuit(tonga):
    import oracle as O
    ds, p = tonga
    op, od = _orc(ds, p)
    op.debug_prior = 1
    r = O.evaluate(op, od, *random_model(np.random.default_rng(0), 7, box_of(ds)))
    assert r["phi"] == 1.0 and r["likelihood"] == 1.0


def test_chain_c_vs_numpy_replay(tonga):
    """The proposal loop: C oracle generates a stream; the NumPy twin replays it -> same accept/phi/K."""
    import oracle as O
    import oracle_np as N
    ds, p = random_ragged(4, R=19, m=12)
    p.min_cells, p.max_cells = 2, 9
    op, od = _orc(ds, p)
    rng = np.random.default_rng(1)
    mdl = random_model(rng, 4, box_of(ds))
    mb = O.ModelBuf(op.max_cells + 1, od.R).set(*mdl)
    assert O.lib().orc_evaluate(op, od.c, mb.c, None, None) == 1
    r = O.chain_run(op, od, mb, 600, g=O.rng(5))
    bx = box_of(ds)
    params = dict(xmin=bx[0], xmax=bx[1], ymin=bx[2], ymax=bx[3], zmin=bx[4], zmax=bx[5], sig=p.sig, zeta_scale=p.zeta_scale,
                  min_cells=p.min_cells, max_cells=p.max_cells)
    data = dict(rayX=ds.rayX, rayY=ds.rayY, rayZ=ds.rayZ, rayL=ds.rayL, rayU=ds.rayU, tS=ds.tS, allSig=ds.allSig)
    acc, phis, Ks, final = N.chain_replay(params, data, dict(zip(("x", "y", "z", "zeta"), mdl)), r.recs)
    assert np.array_equal(acc, r.accept) and np.array_equal(Ks, r.K)
    assert np.allclose(phis, r.phi, rtol=1e-12, atol=0)
    assert set(np.unique(r.recs["action"])) == {1, 2, 3, 4} and 0.05 < r.accept.mean() < 0.95
    assert np.array_equal(np.array(final["x"]), mb.cells()[0]) and np.array_equal(np.array(final["zeta"]), mb.cells()[3])
    assert r.K.min() == 2 and r.K.max() >= 6


def test_thinning_counts(tonga):
    """TD_inversion_function.jl:25,275-281: n_iter=1e3, burn_in=5e2, keep_each=10 -> 50 kept models (iter 509..999)."""
    import oracle as O
    ds, p = random_ragged(2, R=11, m=8)
    op, od = _orc(ds, p)
    mb = O.build_starting(op, od, O.rng(3))
    r = O.chain_run(op, od, mb, 1000, g=O.rng(4), hist_cap=64)
    assert r.n_hist == 50 and r.hist_iter[0] == 509 and r.hist_iter[49] == 999
    assert (np.diff(r.hist_iter[:50]) == 10).all()
    for j in range(50):
        assert r.hist_K[j] == r.K[r.hist_iter[j] - 1] and r.hist_phi[j] == r.phi[r.hist_iter[j] - 1]


def test_prior_sampling_is_log_uniform_in_k():
    """debug_prior = 1: with phi == 1 the chain samples the prior.  The k/(k+1) and k/(k-1) factors
    of the birth/death ratios (TD_inversion_function.jl:96,151; Byrnes & Bezada 2020 eq. 11, 16, 17) encode a LOG-uniform
    prior on nCells, p(k) ~ 1/k on [min_cells, max_cells] (the same law build_starting draws from, MCsub.jl:86-87), and zeta
    uniform on (0, zeta_scale): recovering it validates the restated RJ acceptance ratios end to end."""
    import oracle as O
    ds, p = random_ragged(1, R=5, m=4)
    p.min_cells, p.max_cells, p.debug_prior = 2, 7, 1
    op, od = _orc(ds, p)
    counts = np.zeros(8)
    zs = []
    for c in range(8):
        mb = O.build_starting(op, od, O.rng(100 + c))
        r = O.chain_run(op, od, mb, 40000, g=O.rng(200 + c))
        counts += np.bincount(r.K[2000:], minlength=8)
        zs.append(mb.cells()[3])
    frac = counts[2:8] / counts.sum()
    want = 1.0 / np.arange(2, 8)
    want /= want.sum()
    assert counts[:2].sum() == 0 and np.abs(frac - want).max() < 0.02, (frac, want)
    zs = np.concatenate(zs)
    assert zs.min() > 0 and zs.max() < 50


def test_model_jld_invariants():
    """what the reference's only shipped output pins (it is from an unshipped 487-ray set)."""
    m = np.load(os.path.join(GOLDEN, "model_jld.npz"))
    n = len(m["nCells"])
    assert n == 100 and (m["chain"] == np.repeat([0, 1], 50)).all()
    assert (m["nCells"] >= 5).all() and (m["nCells"] <= 100).all() and (m["nCells"] == m["ncell_len"]).all()
    for i in range(n):
        k = m["ncell_len"][i]
        z = m["cells"][i, 3, :k]
        assert (z > 0).all() and (z < 50).all() and np.isnan(m["cells"][i, :, k:]).all()
        assert (m["cells"][i, 2, :k] >= 0).all() and (m["cells"][i, 2, :k] <= 660).all()
    assert len(np.unique(m["likelihood"])) == 1          # F5: model-independent constant (MCsub.jl:179-180)
    assert m["likelihood"][0] > 0 and len(np.unique(np.round(m["phi"], 6))) > 50
    assert (m["accept"] == 0).all()                       # aliasing artefact, TD_inversion_function.jl:73-74 vs :280
    assert set(np.unique(m["action"])) <= {1, 2, 3, 4}    # rand(1:4): the sigma move is dead (F6)
    assert (m["zeta_xz"] == -1).all() and (m["zeta_xy"] == -1).all()
    assert m["ptS"].shape == (100, 487) and bool(m["tS_identical"])
    # phi is a weighted sum of squared residuals: non-negative and of the order of the data count
    assert (m["phi"] > 0).all() and (m["phi"] < 487).all()
    # the constant equals sum_k(-log(sig_k*sqrt(2pi))) * n  =>  its per-datum mean gives a plausible sigma
    sig_gm = np.exp(-m["likelihood"][0] / 487 / 487) / np.sqrt(2 * np.pi)
    assert 0.05 < sig_gm < 1.0


def test_misfit_and_likelihood_pinned_by_reference_output():
    """Reference-produced known answers for MCsub.jl:169-182.  model.jld (the reference's own output of a 487-ray run) stores ptS,
    tS, phi and likelihood of 100 models; phi / sum((ptS - tS)^2) = 25 for all of them, i.e. allSig = 0.2 for every ray.  With
    that, the oracle's misfit -- the function orc_evaluate itself calls, and its NumPy twin -- reproduces all 100 stored phi
    BIT FOR BIT -- operation order ((d*d)*1.0)/(sig*sig) and the left-to-right loop over the rays (:170-172) are the reference's
    -- and the likelihood constant to 1e-13 (the discarded second line, SURVEY F5, confirmed by a reference-produced value)."""
    import oracle as O
    import oracle_np as ON
    m = np.load(os.path.join(GOLDEN, "model_jld.npz"))
    assert m["ptS"].dtype == np.float64 and m["ptS"].shape == (100, 487)
    tS, sig = m["tS"], np.full(487, 0.2)
    for i in range(100):
        phi_c, like_c, lg = O.misfit(m["ptS"][i], tS, sig)
        assert phi_c == m["phi"][i], (i, phi_c, m["phi"][i])            # bit for bit
        # the likelihood is a Julia `sum(...)` (:179) whose SIMD order is unspecified: 487 equal terms, so any order agrees to a few ulp
        assert abs(like_c - m["likelihood"][i]) <= 1e-13 * like_c, (like_c, m["likelihood"][i])
        assert abs(lg - (like_c / 487 - 0.5 * phi_c)) < 1e-9 * abs(lg)
        if i % 10 == 0:
            phi_n, like_n = ON.misfit(m["ptS"][i], tS, sig)
            assert phi_n == m["phi"][i] and like_n == like_c
    # sigma is identified by the data, not assumed: a 1-ulp different sigma already breaks the equality
    assert O.misfit(m["ptS"][0], tS, np.full(487, np.nextafter(0.2, 1.0)))[0] != m["phi"][0]
