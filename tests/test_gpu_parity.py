"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): cell indices bit-exact; t* and phi within 1e-9 relative (FP64); accept/reject
decisions identical on replayed proposal streams.  The oracle sums t*/phi left to right, the device uses its
canonical tree order, hence a tolerance on the sums (never on the indices).
"""
import os
import sys

import numpy as np
import pytest

from conftest import box_of, random_model, random_ragged

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _orc(ds, p, **kw):
    import oracle as O
    return O.make_params(box_of(ds), sig=p.sig, zeta_scale=p.zeta_scale, max_sig=p.max_sig, n_iter=p.n_iter, burn_in=p.burn_in,
                         keep_each=p.keep_each, min_cells=p.min_cells, max_cells=p.max_cells, prior=p.prior,
                         debug_prior=p.debug_prior, **kw), O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)


def _flat_owners(own_mR, ctx):
    """oracle owners [m, R] -> flat CSR order used by the C ABI."""
    off = ctx.ray_offsets()
    return np.concatenate([own_mR[:off[i + 1] - off[i], i] for i in range(ctx.R)]) if ctx.R else np.zeros(0, np.int32)


@pytest.fixture(scope="module")
def tonga_ctx(tonga):
    from tonga_b200.api import Context
    ds, p = tonga
    return Context(ds, p), ds, p


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + 1e-300)


@pytest.mark.parametrize("K", [1, 5, 20, 100])
def test_evaluate_tonga381(tonga_ctx, K):
    import oracle as O
    ctx, ds, p = tonga_ctx
    op, od = _orc(ds, p)
    rng = np.random.default_rng(100 + K)
    x, y, z, zeta = random_model(rng, K, box_of(ds))
    ref = O.evaluate(op, od, x, y, z, zeta, want_owners=True)
    got = ctx.evaluate(x, y, z, zeta)
    Kb, cells = np.array([K], np.int32), np.stack([x, y, z, zeta])[None]
    gb = ctx.evaluate_batch(Kb, cells, want_owners=True)
    assert np.array_equal(gb["owners"][0], _flat_owners(ref["owners"], ctx)), "owners must be bit-exact"
    assert _close(got["ptS"], ref["ptS"]) and _close(gb["ptS"][0], ref["ptS"])
    assert _close(got["phi"], ref["phi"]) and gb["phi"][0] == got["phi"]
    assert _close(got["likelihood"], ref["likelihood"], 1e-12)
    assert _close(got["loglik_gauss"], ref["loglik_gauss"])


def test_misfit_kernel_against_reference_output():
    """Reference-PRODUCED known answers (not the oracle): model.jld, the reference's own output of a 487-ray run, stores ptS, tS and
    phi of 100 models (allSig = 0.2, see tests/test_oracle.py).  The device misfit (tg_phi_kernel through tonga_misfit -- the kernel
    every evaluate and every proposal ends in) must reproduce the stored phi to 1e-12 (its canonical summation order differs from
    the reference's left-to-right loop; the contract is 1e-9), and evaluate's model-independent "likelihood" (MCsub.jl:179, F5)
    must equal the stored constant."""
    from tonga_b200.api import Context
    m = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_jld.npz"))
    ds, p = random_ragged(3, R=487, m=12, tS=m["tS"].copy(), sig=np.full(487, 0.2))
    ctx = Context(ds, p)
    phi = ctx.misfit(m["ptS"])
    assert phi.shape == (100,) and np.all(np.abs(phi - m["phi"]) <= 1e-12 * m["phi"]), np.abs(phi / m["phi"] - 1).max()
    assert ctx.misfit(m["ptS"][:1], noise=[2.0])[0] == pytest.approx(m["phi"][0] / 4, rel=1e-12)  # hierarchical noise scales allSig
    rng = np.random.default_rng(0)
    got = ctx.evaluate(*random_model(rng, 7, box_of(ds)))
    assert abs(got["likelihood"] - m["likelihood"][0]) <= 1e-12 * m["likelihood"][0]
    ctx.close()


def test_evaluate_batch_many_models(tonga_ctx):
    import oracle as O
    ctx, ds, p = tonga_ctx
    op, od = _orc(ds, p)
    rng = np.random.default_rng(7)
    models = [random_model(rng, int(k), box_of(ds)) for k in rng.integers(1, 101, 24)]
    from tonga_b200.api import pack_models
    K, cells = pack_models(models, Kcap=104)
    noise = rng.uniform(0.5, 2.0, len(models))
    noise[:4] = 1.0
    gb = ctx.evaluate_batch(K, cells, noise=noise, want_owners=True)
    for i, mdl in enumerate(models):
        ref = O.evaluate(op, od, *mdl, noise=noise[i], want_owners=True)
        assert np.array_equal(gb["owners"][i], _flat_owners(ref["owners"], ctx))
        assert _close(gb["ptS"][i], ref["ptS"]) and _close(gb["phi"][i], ref["phi"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_evaluate_random_ragged(seed):
    import oracle as O
    from tonga_b200.api import Context
    ds, p = random_ragged(seed, R=37 + 11 * seed, m=9 + 30 * seed)
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(seed)
    for K in (1, 3, 17, 64):
        mdl = random_model(rng, K, box_of(ds))
        ref = O.evaluate(op, od, *mdl, want_owners=True)
        gb = ctx.evaluate_batch(np.array([K], np.int32), np.stack(mdl)[None], want_owners=True)
        assert np.array_equal(gb["owners"][0], _flat_owners(ref["owners"], ctx))
        assert _close(gb["ptS"][0], ref["ptS"]) and _close(gb["phi"][0], ref["phi"])
    ctx.close()


def test_evaluate_large_K_uses_bulk_copy(tonga_ctx):
    """K = 2000 nuclei (BASELINE config 3 regime): 64 KB of nuclei staged by one TMA bulk copy."""
    import oracle as O
    ctx, ds, p = tonga_ctx
    op, od = _orc(ds, p)
    rng = np.random.default_rng(5)
    mdl = random_model(rng, 2000, box_of(ds))
    ref = O.evaluate(op, od, *mdl, want_owners=True)
    gb = ctx.evaluate_batch(np.array([2000], np.int32), np.stack(mdl)[None], want_owners=True)
    assert np.array_equal(gb["owners"][0], _flat_owners(ref["owners"], ctx))
    assert _close(gb["ptS"][0], ref["ptS"]) and _close(gb["phi"][0], ref["phi"])


def test_v_nearest_kats():
    """tie -> lowest index; nothing within sqrt(1e9) -> 0.0; single nucleus; NaN trim."""
    import oracle as O
    from tonga_b200.api import Context, v_nearest
    ctx = Context(None)
    # exact tie between nuclei 1 and 2 (mirror images), nucleus 0 farther
    mx, my, mz, mv = np.array([9.0, 1.0, -1.0]), np.zeros(3), np.zeros(3), np.array([7.0, 11.0, 13.0])
    out, idx = ctx.interpolate(mx, my, mz, mv, [0.0], [0.0], [0.0])
    assert idx[0] == 1 and out[0] == 11.0
    assert O.v_nearest(0.0, 0.0, 0.0, mx, my, mz, mv) == (11.0, 1)
    # all nuclei farther than sqrt(1e9) km: v = 0.0 (MCsub.jl:249-250)
    out, idx = ctx.interpolate(np.array([4e4]), np.zeros(1), np.zeros(1), np.array([5.0]), [0.0], [0.0], [0.0])
    assert out[0] == 0.0 and idx[0] == -1
    assert O.v_nearest(0.0, 0.0, 0.0, np.array([4e4]), np.zeros(1), np.zeros(1), np.array([5.0])) == (0.0, -1)
    # single nucleus
    assert v_nearest(1.0, 2.0, 3.0, [0.0], [0.0], [0.0], [42.0]) == 42.0
    # NaN trim (MCsub.jl:312-316) and slice broadcast (:317-322)
    X = np.array([0.0, 1.0, 2.0, np.nan, 5.0])
    out, idx = ctx.interpolate(mx, my, mz, mv, X, [0.0], [0.0])
    ro, ri = O.interpolation(mx, my, mz, mv, X, [0.0], [0.0])
    assert len(out) == 3 and np.array_equal(out, ro) and np.array_equal(idx, ri)
    # random points / nuclei
    rng = np.random.default_rng(3)
    mx, my, mz, mv = rng.normal(size=(4, 57)) * 100
    X, Y, Z = rng.normal(size=(3, 1000)) * 100
    out, idx = ctx.interpolate(mx, my, mz, mv, X, Y, Z)
    ro, ri = O.interpolation(mx, my, mz, mv, X, Y, Z)
    assert np.array_equal(idx, ri) and np.array_equal(out, ro)
    ctx.close()


def _start_models(rng, n, ds, kmin=5, kmax=40):
    return [random_model(rng, int(k), box_of(ds)) for k in rng.integers(kmin, kmax + 1, n)]


def _oracle_generate(op, od, mdl, n_iter, seed, hist_cap=0):
    import oracle as O
    mb = O.ModelBuf(op.max_cells + 1, od.R).set(*mdl)
    assert O.lib().orc_evaluate(op, od.c, mb.c, None, None) == 1
    g = O.rng(seed)
    return O.chain_run(op, od, mb, n_iter, g=g, hist_cap=hist_cap), mb


@pytest.mark.parametrize("prior", [1, 2, 3])
def test_replay_oracle_stream_on_device(tonga, prior):
    """Oracle draws the proposals; the device replays them: identical accept/reject, K, final model; phi to 1e-9."""
    import copy
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = tonga
    p = copy.copy(p0)
    p.prior = prior
    p.n_iter, p.burn_in, p.keep_each = 400.0, 100.0, 10.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(11 + prior)
    n, n_iter = 6, 400
    models = _start_models(rng, n, ds)
    runs = [_oracle_generate(op, od, models[c], n_iter, 1000 + c, hist_cap=64) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n, hist_cap=64)
    K, cells = pack_models(models, Kcap=ch.KC)
    ch.set_models(K, cells)
    out = ch.run(n_iter, recs=recs, trace=True)
    st = ch.state(want_owners=True)
    hist = ch.history()
    for c, (r, mb) in enumerate(runs):
        assert np.array_equal(out["accept"][c], r.accept), f"accept/reject sequence differs (chain {c})"
        assert np.array_equal(out["K"][c], r.K)
        assert _close(out["phi"][c], r.phi)
        k = mb.K
        assert st["K"][c] == k
        for a, ref in enumerate(mb.cells()):
            assert np.array_equal(st["cells"][c, a, :k], ref), "final nuclei must be bit-identical"
        assert _close(st["ptS"][c], mb.ptS[:od.R])
        # thinning (TD_inversion_function.jl:275-281)
        assert hist["n_hist"][c] == r.n_hist == 30
        assert np.array_equal(hist["K"][c, :r.n_hist], r.hist_K[:r.n_hist])
        assert np.array_equal(hist["iter"][c, :r.n_hist], r.hist_iter[:r.n_hist])
        assert np.array_equal(hist["action"][c, :r.n_hist], r.hist_action[:r.n_hist])
        assert np.array_equal(hist["accept"][c, :r.n_hist], r.hist_accept[:r.n_hist])
        assert _close(hist["phi"][c, :r.n_hist], r.hist_phi[:r.n_hist])
        for j in range(r.n_hist):
            kj = r.hist_K[j]
            assert np.array_equal(hist["cells"][c, j, :, :kj], r.hist_cells[j, :, :kj])
        assert _close(hist["ptS"][c, :r.n_hist], r.hist_ptS[:r.n_hist])
    mm, dphi, dts = ch.verify()
    assert (mm, dphi, dts) == (0, 0.0, 0.0), "incremental state must equal a full evaluate bit for bit"
    ch.close(); ctx.close()


def test_device_stream_replayed_by_oracle(tonga):
    """The device draws (Philox) and records; the oracle replays: identical decisions and final models."""
    import copy
    import oracle as O
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 600.0, 300.0, 10.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    n, n_iter = 8, 600
    ch = Chains(ctx, n, chain_id0=3, seed=42)
    ch.build_starting()
    st0 = ch.state()
    out = ch.run(n_iter, record=True, trace=True)
    st1 = ch.state()
    assert set(np.unique(out["recs"]["action"])) <= {1, 2, 3, 4}
    for c in range(n):
        k0 = st0["K"][c]
        mb = O.ModelBuf(op.max_cells + 1, od.R).set(*[st0["cells"][c, a, :k0] for a in range(4)])
        assert O.lib().orc_evaluate(op, od.c, mb.c, None, None) == 1
        assert _close(mb.c.phi, st0["phi"][c])
        r = O.chain_run(op, od, mb, n_iter, recs=out["recs"][c])
        assert np.array_equal(r.accept, out["accept"][c])
        assert np.array_equal(r.K, out["K"][c])
        assert _close(r.phi, out["phi"][c])
        k1 = mb.K
        assert k1 == st1["K"][c]
        for a, ref in enumerate(mb.cells()):
            assert np.array_equal(st1["cells"][c, a, :k1], ref)
    it, counts = ch.stats()
    assert it == n_iter and counts[:, 0].sum() == n * n_iter
    assert (counts[:, 1] <= counts[:, 0]).all() and counts[:, 1].sum() == out["accept"].sum()
    assert ch.verify() == (0, 0.0, 0.0)
    ch.close(); ctx.close()


def test_incremental_equals_full_long_run(tonga_ctx):
    """after thousands of birth/death/change/move proposals on many chains the incrementally
    maintained owners / t* / phi equal a from-scratch evaluate, bit for bit."""
    from tonga_b200.api import Chains
    ctx, ds, p = tonga_ctx
    ch = Chains(ctx, 64, seed=7, hist_cap=0)
    ch.build_starting()
    for _ in range(3):
        ch.run(1500)
        assert ch.verify() == (0, 0.0, 0.0)
    it, counts = ch.stats()
    acc = counts[:, 1].sum(0) / np.maximum(counts[:, 0].sum(0), 1)
    assert (counts[:, 0].sum(0)[:4] > 0).all() and counts[:, 0].sum(0)[4] == 0
    assert 0.01 < acc[:4].min() and acc[:4].max() < 0.99
    st = ch.state()
    assert (st["K"] >= p.min_cells).all() and (st["K"] <= p.max_cells).all()
    ch.close()


def test_replay_edge_cases(tonga):
    """A-priori rejections (birth at max_cells, death at min_cells, zeta out of bounds, move out of the box) and exact
    behaviour at the K limits, on a small random ragged set with tiny K limits."""
    import copy
    import oracle as O
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = random_ragged(5, R=29, m=23)
    p = copy.copy(p0)
    p.min_cells, p.max_cells = 2, 6
    p.n_iter, p.burn_in, p.keep_each = 1500.0, 0.0, 50.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(2)
    n, n_iter = 5, 1500
    models = _start_models(rng, n, ds, 2, 6)
    runs = [_oracle_generate(op, od, models[c], n_iter, 50 + c) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n)
    ch.set_models(*pack_models(models, Kcap=ch.KC))
    out = ch.run(n_iter, recs=recs, trace=True)
    for c, (r, mb) in enumerate(runs):
        assert np.array_equal(out["accept"][c], r.accept) and np.array_equal(out["K"][c], r.K)
        assert _close(out["phi"][c], r.phi)
    allK = np.concatenate([r.K for r, _ in runs])
    assert allK.min() == 2 and allK.max() == 6  # both limits were hit
    assert ch.verify() == (0, 0.0, 0.0)
    ch.close(); ctx.close()


def test_sigma_move_extension(tonga):
    """Action 5 (hierarchical noise; dead code in the reference, spec in DESIGN.md): oracle stream replayed on device."""
    import copy
    import oracle as O
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = tonga
    p = copy.copy(p0)
    p.max_sig = 3.0
    ctx = Context(ds, p, n_actions=5)
    op, od = _orc(ds, p, n_actions=5)
    rng = np.random.default_rng(21)
    n, n_iter = 4, 500
    models = _start_models(rng, n, ds)
    runs = [_oracle_generate(op, od, models[c], n_iter, 900 + c) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    assert (recs["action"] == 5).any()
    ch = Chains(ctx, n)
    ch.set_models(*pack_models(models, Kcap=ch.KC))
    out = ch.run(n_iter, recs=recs, trace=True)
    st = ch.state()
    for c, (r, mb) in enumerate(runs):
        assert np.array_equal(out["accept"][c], r.accept) and np.array_equal(out["K"][c], r.K)
        assert _close(out["phi"][c], r.phi)
        assert st["noise"][c] == mb.c.noise
    assert ch.verify() == (0, 0.0, 0.0)
    ch.close(); ctx.close()


def test_debug_prior_mode(tonga):
    """debug_prior = 1 (MCsub.jl:134-136): phi == 1 always, so the chain samples the prior."""
    import copy
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    p.debug_prior = 1
    ctx = Context(ds, p)
    r = ctx.evaluate([1.0], [2.0], [3.0], [4.0])
    assert r["phi"] == 1.0 and r["likelihood"] == 1.0
    ctx.close()


def test_reference_named_api(tonga):
    import copy
    from tonga_b200 import api
    from tonga_b200.structs import Model
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 300.0, 100.0, 10.0
    rng = np.random.default_rng(0)
    x, y, z, zeta = random_model(rng, 12, box_of(ds))
    mdl = Model(12.0, x, y, z, zeta)
    mdl, ds2, valid = api.evaluate(mdl, ds, p)
    assert valid == 1 and ds2 is ds and mdl.ptS.shape == (381,) and mdl.tS is ds.tS and mdl.phi > 0
    zs = api.Interpolation(p, mdl, ds.rayX[:, 0], ds.rayY[:, 0], ds.rayZ[:, 0])
    n0 = int((~np.isnan(ds.rayX[:, 0])).sum())
    assert len(zs) == n0
    hists = api.run_chains(p, ds, range(1, 4))
    assert len(hists) == 3 and all(len(h) == 20 for h in hists)
    for h in hists:
        for m_ in h:
            assert p.min_cells <= m_.nCells <= p.max_cells and len(m_.xCell) == int(m_.nCells)
            assert (m_.zeta > 0).all() and (m_.zeta < p.zeta_scale).all()
    one = api.TD_inversion_function(p, ds, 2)
    assert len(one) == 20 and one[0].phi == hists[1][0].phi  # chain 2's stream does not depend on the batch it runs in
    m0, _, v = api.build_starting(p, ds, chain=1)
    assert v == 1 and p.min_cells <= m0.nCells <= p.max_cells


def test_sharded_chains_match_unsharded(tonga):
    """chain c's samples do not depend on how chains are split over GPUs -- two half batches with
    the right global chain ids reproduce the full batch bit for bit (Philox keyed by the global chain id)."""
    import copy
    from tonga_b200.api import Chains, Context
    from tonga_b200.dist import history_tensors, shard
    import torch
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 300.0, 100.0, 10.0
    ctx = Context(ds, p)
    full = Chains(ctx, 10, chain_id0=1, seed=99)
    full.build_starting(); full.run(300)
    hf = full.history()
    parts = []
    for r in range(2):
        s, c = shard(10, 2, r)
        ch = Chains(ctx, c, chain_id0=1 + s, seed=99)
        ch.build_starting(); ch.run(300)
        ht = history_tensors(ch, torch.device("cuda", 0))  # zero-copy views of the library's device buffers
        h = ch.history()
        assert np.array_equal(ht["phi"].cpu().numpy(), h["phi"]) and np.array_equal(ht["K"].cpu().numpy(), h["K"])
        assert np.array_equal(ht["cells"].cpu().numpy(), h["cells"]) and np.array_equal(ht["ptS"].cpu().numpy(), h["ptS"])
        parts.append(h)
    for k in ("n_hist", "K", "cells", "phi", "ptS", "iter", "action", "accept"):
        assert np.array_equal(np.concatenate([parts[0][k], parts[1][k]]), hf[k]), k


def test_fp32_screening_changes_nothing(tonga):
    """The FP32 screening pass + exact FP64 recheck of near ties must give the same chain, bit for bit, as the pure FP64
    path (owners, t*, phi, nuclei, accept decisions) -- on the Tonga set and on a set engineered to be hard for FP32:
    coordinates offset by 5e4 km (large |x| -> coarse fl32 grid) with nuclei mirrored about ray points (exact ties)."""
    import copy
    from tonga_b200.api import Chains, Context, pack_models
    from tonga_b200.data import make_datastruct
    from tonga_b200.structs import StepRangeLen
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 800.0, 400.0, 10.0
    # hard set: shift everything far from the origin
    off = 5.0e4
    dsh = make_datastruct(ds.rayX + off, ds.rayY + off, ds.rayZ, ds.U, ds.tS, ds.allSig, p,
                          box=(StepRangeLen(ds.xVec.min() + off, 20, ds.xVec.max() + off), StepRangeLen(ds.yVec.min() + off, 20, ds.yVec.max() + off), ds.zVec))
    for data in (ds, dsh):
        ctx = Context(data, p)
        outs = []
        for exact in (False, True):
            ch = Chains(ctx, 24, seed=5)
            ch.set_exact_only(exact)
            ch.build_starting()
            if data is dsh:  # plant exact ties: nucleus 1 = mirror image of nucleus 0 about a ray point
                st = ch.state(want_ptS=False)
                K, cells = st["K"].copy(), st["cells"].copy()
                pt = np.array([data.rayX[3, 7], data.rayY[3, 7], data.rayZ[3, 7]])
                for c in range(len(K)):
                    cells[c, :3, 1] = 2 * pt - cells[c, :3, 0]
                    cells[c, 2, 1] = min(max(cells[c, 2, 1], 0.0), 660.0)
                ch.set_models(K, cells)
            out = ch.run(800, record=True, trace=True)
            st = ch.state(want_owners=True)
            assert ch.verify() == (0, 0.0, 0.0)
            outs.append((out, st))
            ch.close()
        (o0, s0), (o1, s1) = outs
        assert np.array_equal(o0["accept"], o1["accept"]) and np.array_equal(o0["phi"], o1["phi"]) and np.array_equal(o0["K"], o1["K"])
        assert np.array_equal(o0["recs"], o1["recs"])
        for k in ("K", "cells", "phi", "ptS", "owners"):
            assert np.array_equal(s0[k], s1[k]), k
        ctx.close()


def test_posterior_raster_matches_oracle(tonga):
    """plot_model_hist's numerics (MCsub.jl:753-825): mean / std / mask of v_nearest over the kept models on the reference's
    own slices (xVec x zVec at y in ySlice, xVec x yVec at z in zSlice) against the oracle's Interpolation per model."""
    import copy
    import oracle as O
    from tonga_b200 import api
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 200.0, 100.0, 10.0
    p.ySlice = [100, 300]  # inside this data set's frame (the reference's 700/800 belong to another frame)
    ctx = api.Context(ds, p)
    ch = api.Chains(ctx, 6, seed=3)
    ch.build_starting(); ch.run(200)
    hist = ch.history()
    maps = api.plot_model_hist(ch, ds, p)
    assert set(maps) == {("xz", 100.0), ("xz", 300.0), ("xy", 50.0), ("xy", 300.0), ("xy", 500.0)}
    for kind, l0, shape, X, Y, Z in api.slice_nodes(ds, p):
        vals = []
        for c in range(6):
            for j in range(hist["n_hist"][c]):
                k = hist["K"][c, j]
                v, _ = O.interpolation(*[hist["cells"][c, j, a, :k] for a in range(4)], X, Y, Z)
                vals.append(v)
        vals = np.array(vals)
        m = maps[(kind, l0)]
        assert m["count"] == len(vals) == 60 and m["mean"].shape == shape == ((58, 34))
        ref_mean, ref_std = vals.mean(0).reshape(shape, order="F"), vals.std(0, ddof=1).reshape(shape, order="F")
        assert np.allclose(m["mean"], ref_mean, rtol=1e-12, atol=1e-12) and np.allclose(m["std"], ref_std, rtol=1e-9, atol=1e-9)
        sure = np.abs(ref_std - 5) > 1e-6
        assert np.array_equal(np.isnan(m["masked"])[sure], (ref_std > 5)[sure])
    ch.close(); ctx.close()


def test_posterior_agrees_with_oracle_within_monte_carlo_error():
    """north_star: 'posterior mean/sigma of 1000/Q agree within Monte-Carlo error'.  Device chains (Philox) and oracle chains
    (its own RNG) sample the same posterior on a small problem; means of nCells, phi and of zeta at fixed nodes must agree
    within 5 standard errors (between-chain scatter)."""
    import copy
    import oracle as O
    from tonga_b200 import api
    ds, p0 = random_ragged(11, R=31, m=15)
    p = copy.copy(p0)
    p.min_cells, p.max_cells = 2, 8
    p.n_iter, p.burn_in, p.keep_each = 6000.0, 2000.0, 10.0
    n = 48
    ctx = api.Context(ds, p)
    ch = api.Chains(ctx, n, seed=77)
    ch.build_starting(); ch.run(6000)
    hist = ch.history()
    rng = np.random.default_rng(0)
    bx = box_of(ds)
    X, Y, Z = rng.uniform(bx[0], bx[1], 40), rng.uniform(bx[2], bx[3], 40), rng.uniform(bx[4], bx[5], 40)
    def chain_stats(K, phi, cells):  # per-chain means over kept models
        z = np.mean([O.interpolation(*[cells[j, a, :K[j]] for a in range(4)], X, Y, Z)[0] for j in range(len(K))], axis=0)
        return np.concatenate([[K.mean(), phi.mean()], z])
    dev = np.array([chain_stats(hist["K"][c, :hist["n_hist"][c]], hist["phi"][c, :hist["n_hist"][c]], hist["cells"][c]) for c in range(n)])
    op = O.make_params(bx, n_iter=6000, burn_in=2000, keep_each=10, min_cells=2, max_cells=8)
    od = O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)
    orc = []
    for c in range(n):
        g = O.rng(500 + c)
        mb = O.build_starting(op, od, g)
        r = O.chain_run(op, od, mb, 6000, g=g, hist_cap=400)
        orc.append(chain_stats(r.hist_K[:r.n_hist], r.hist_phi[:r.n_hist], r.hist_cells))
    orc = np.array(orc)
    assert hist["n_hist"].min() == 400
    se = np.sqrt(dev.var(0, ddof=1) / n + orc.var(0, ddof=1) / n)
    zscore = np.abs(dev.mean(0) - orc.mean(0)) / np.maximum(se, 1e-12)
    assert zscore.max() < 5.0, (zscore.max(), dev.mean(0)[:2], orc.mean(0)[:2])
    ch.close(); ctx.close()


def test_parallel_tempering_keeps_the_cold_posterior():
    """Extension (BASELINE config 5): ladders of 8 temperatures with the hierarchical sigma move; swaps exchange betas.
    The beta = 1 replicas must still sample the untempered posterior: their mean nCells / phi / noise agree with plain
    chains within Monte-Carlo error, swaps do happen, and hot replicas never write history."""
    import copy
    from tonga_b200 import api
    from tonga_b200.tempering import run_tempered
    ds, p0 = random_ragged(11, R=31, m=15)
    p = copy.copy(p0)
    p.min_cells, p.max_cells, p.max_sig = 2, 8, 4.0
    p.n_iter, p.burn_in, p.keep_each = 6000.0, 2000.0, 10.0
    ctx = api.Context(ds, p, n_actions=5)
    T, L = 8, 24
    pt = api.Chains(ctx, T * L, seed=5, hist_cap=401)
    pt.build_starting()
    info = run_tempered(pt, 6000, ladder_size=T, swap_every=50, t_max=20.0, seed=1)
    assert 0.05 < info["swap_rate"] < 0.999
    for l in range(L):
        assert np.array_equal(np.sort(info["beta"][l * T:(l + 1) * T])[::-1], np.sort(info["beta"][:T])[::-1])
    hp = pt.history(want_ptS=False)
    assert hp["n_hist"].sum() == L * 400  # exactly one replica per ladder is cold at any time
    plain = api.Chains(ctx, 64, seed=6, hist_cap=401)
    plain.build_starting(); plain.run(6000)
    hq = plain.history(want_ptS=False)
    def pooled(h):
        K = np.concatenate([h["K"][c, :min(h["n_hist"][c], 401)] for c in range(len(h["n_hist"]))]).astype(float)
        phi = np.concatenate([h["phi"][c, :min(h["n_hist"][c], 401)] for c in range(len(h["n_hist"]))])
        return K, phi
    Kp, phip = pooled(hp)
    Kq, phiq = pooled(hq)
    # effective sample sizes are far below the raw counts (autocorrelation): compare with a generous 8-sigma band on an
    # assumed integrated autocorrelation time of 20 kept samples
    for a, b in ((Kp, Kq), (phip, phiq)):
        se = np.sqrt(a.var() * 20 / len(a) + b.var() * 20 / len(b))
        assert abs(a.mean() - b.mean()) < 8 * se, (a.mean(), b.mean(), se)
    pt.close(); plain.close(); ctx.close()


def test_evaluate_screening_equals_exact(tonga_ctx):
    """Full evaluate: FP32 screening + exact re-scan must give bit-identical owners / t* / phi to the pure FP64 pass, also with
    planted exact ties (mirror-image nuclei) and K = 2000 (small cells: many near ties)."""
    ctx, ds, p = tonga_ctx
    rng = np.random.default_rng(9)
    for K in (2, 7, 100, 2000):
        mdl = list(random_model(rng, K, box_of(ds)))
        if K >= 2:  # nucleus 1 = mirror image of nucleus 0 about a ray point -> exact tie at that point
            pt = np.array([ds.rayX[5, 11], ds.rayY[5, 11], ds.rayZ[5, 11]])
            for a in range(3):
                mdl[a][1] = 2 * pt[a] - mdl[a][0]
        Kb, cells = np.array([K], np.int32), np.stack(mdl)[None]
        ctx.set_exact_only(False)
        a = ctx.evaluate_batch(Kb, cells, want_owners=True)
        ctx.set_exact_only(True)
        b = ctx.evaluate_batch(Kb, cells, want_owners=True)
        ctx.set_exact_only(False)
        assert np.array_equal(a["owners"], b["owners"]) and np.array_equal(a["ptS"], b["ptS"]) and np.array_equal(a["phi"], b["phi"])


def test_evaluate_synthetic_config3_shape():
    """BASELINE config 3 style synthetic set (straight jittered rays, 150-250 points each, K up to 500) at reduced ray count:
    owners bit-exact and t*/phi within 1e-9 of the oracle."""
    import oracle as O
    from tonga_b200.api import Context
    from tonga_b200.data import synthetic_rays
    from tonga_b200.structs import parameters
    p = parameters()
    ds = synthetic_rays(400, seed=3, p=p)
    ctx = Context(ds, p)
    assert ctx.P > 60000
    op, od = _orc(ds, p)
    rng = np.random.default_rng(4)
    for K in (50, 500):
        mdl = random_model(rng, K, box_of(ds))
        ref = O.evaluate(op, od, *mdl, want_owners=True)
        gb = ctx.evaluate_batch(np.array([K], np.int32), np.stack(mdl)[None], want_owners=True)
        assert np.array_equal(gb["owners"][0], _flat_owners(ref["owners"], ctx))
        assert _close(gb["ptS"][0], ref["ptS"]) and _close(gb["phi"][0], ref["phi"])
    ctx.close()


def test_abi_error_paths_and_helpers(tonga):
    import ctypes as C
    """Error behaviour at the C ABI (codes + messages, no exceptions across it) and the small helper entry points."""
    import copy
    from tonga_b200 import _lib, api
    ds, p0 = tonga
    ctx = api.Context(ds, p0)
    # evaluate: K outside [0, Kcap]
    with pytest.raises(_lib.TongaError) as e:
        ctx.evaluate_batch(np.array([9], np.int32), np.zeros((1, 4, 4)))
    assert e.value.code == -1
    # misfit: NULL arrays -> TONGA_ERR_ARG; nothing to do for 0 models; the misfit of the model's own t* is its phi
    lib = _lib.load()
    assert lib.tonga_misfit(ctx._h, 1, None, None, None) == -1 and b"tonga_misfit" in lib.tonga_last_error()
    assert lib.tonga_misfit(ctx._h, 0, C.cast(1, C.POINTER(C.c_double)), None, C.cast(1, C.POINTER(C.c_double))) == 0
    g = ctx.evaluate_batch(np.array([3], np.int32), np.array([[[10.0, 200.0, 400.0], [5.0, 100.0, 300.0], [50.0, 150.0, 400.0], [5.0, 20.0, 40.0]]]))
    assert ctx.misfit(g["ptS"])[0] == g["phi"][0]
    # chains: run before start models -> TONGA_ERR_STATE; K above capacity -> TONGA_ERR_CAPACITY
    ch = api.Chains(ctx, 2, hist_cap=0)
    with pytest.raises(_lib.TongaError) as e:
        ch.run(1)
    assert e.value.code == -4
    with pytest.raises(_lib.TongaError) as e:
        ch.set_models(np.array([200, 5], np.int32), np.zeros((2, 4, 200)))
    assert e.value.code == -3
    ch.build_starting()
    s1, s2, cnt = ch.raster([0.0, 1.0], [0.0, 1.0], [0.0, 1.0])  # no history kept -> count 0, zero sums
    assert cnt == 0 and not s1.any() and not s2.any()
    # profile counters and pinned buffers
    ch.profile(True); ch.run(50); cyc = ch.profile(False, read=True)
    assert cyc.shape == (2, 16) and (cyc[:, :6] > 0).all()
    hist, state = ch.alloc_buffers(pinned=True)
    st = ch.state(out=state)
    assert st["K"] is state["K"] and (st["K"] >= p0.min_cells).all()
    ch.close()
    # u8 owner state supports at most 126 cells
    p = copy.copy(p0); p.max_cells = 127
    ctx2 = api.Context(ds, p)
    with pytest.raises(_lib.TongaError) as e:
        api.Chains(ctx2, 1, sampler="resident")
    assert e.value.code == -3 and "126" in str(e.value)
    chw = api.Chains(ctx2, 1)  # auto -> the streamed sampler (per-point state in HBM)
    assert chw.sampler == "streamed"
    with pytest.raises(_lib.TongaError) as e:
        chw.profile(True)
    assert e.value.code == -4
    chw.close()
    # interp_style 2 is broken in the reference (MCsub.jl:332) -> refused
    p = copy.copy(p0); p.interp_style = 2
    with pytest.raises(_lib.TongaError) as e:
        api.Context(ds, p)
    assert e.value.code == -1
    with pytest.raises(_lib.TongaError) as e:
        api.Context(ds, p, device_ingest=True)
    assert e.value.code == -1
    # device ingest: slowness matrix of the wrong shape; history in host memory asked from a device-history batch
    import dataclasses
    bad = dataclasses.replace(ds, U=ds.U[:-1]) if dataclasses.is_dataclass(ds) else None
    if bad is not None:
        with pytest.raises(_lib.TongaError):
            api.Context(bad, p0, device_ingest=True)
    chd = api.Chains(ctx, 1, hist_cap=2)
    ptrs = [C.c_void_p() for _ in range(8)]
    assert chd.lib.tonga_chains_history_host(chd._h, *[C.byref(q) for q in ptrs]) == -4
    assert chd.lib.tonga_chains_set_progress(chd._h, -1, None, None, None) == -1
    chd.close()
    with pytest.raises(_lib.TongaError) as e:
        api.Chains(ctx, 1, sampler="streamed", hist_cap=0).restore(dict(hist_cap=1, K=np.zeros(1), chain_id0=0, seed=0))
    assert e.value.code == -1
    ctx.close(); ctx2.close()


# ------------------------------------------------------------------------------------------------ wide sampler
def _check_replay(ch, out, runs, od, hist_n=None):
    st = ch.state(want_owners=False)
    hist = ch.history()
    for c, (r, mb) in enumerate(runs):
        assert np.array_equal(out["accept"][c], r.accept), f"accept/reject sequence differs (chain {c})"
        assert np.array_equal(out["K"][c], r.K)
        assert _close(out["phi"][c], r.phi)
        k = mb.K
        assert st["K"][c] == k
        for a, ref in enumerate(mb.cells()):
            assert np.array_equal(st["cells"][c, a, :k], ref), "final nuclei must be bit-identical"
        assert _close(st["ptS"][c], mb.ptS[:od.R])
        if hist_n is not None:
            assert hist["n_hist"][c] == r.n_hist == hist_n
            assert np.array_equal(hist["K"][c, :r.n_hist], r.hist_K[:r.n_hist])
            assert np.array_equal(hist["action"][c, :r.n_hist], r.hist_action[:r.n_hist])
            assert np.array_equal(hist["accept"][c, :r.n_hist], r.hist_accept[:r.n_hist])
            assert _close(hist["phi"][c, :r.n_hist], r.hist_phi[:r.n_hist])
            for j in range(r.n_hist):
                kj = r.hist_K[j]
                assert np.array_equal(hist["cells"][c, j, :, :kj], r.hist_cells[j, :, :kj])
            assert _close(hist["ptS"][c, :r.n_hist], r.hist_ptS[:r.n_hist])


@pytest.mark.parametrize("kind", ["wide", "streamed"])
@pytest.mark.parametrize("prior", [1, 2, 3])
def test_wide_sampler_replays_oracle_stream(tonga, prior, kind):
    """The wide sampler (full forward model per proposal) and the streamed sampler (per-point state in HBM, incremental
    passes) against the oracle's chain."""
    import copy
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = tonga
    p = copy.copy(p0)
    p.prior = prior
    p.n_iter, p.burn_in, p.keep_each = 300.0, 100.0, 10.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(51 + prior)
    n, n_iter = 5, 300
    models = _start_models(rng, n, ds)
    runs = [_oracle_generate(op, od, models[c], n_iter, 2000 + c, hist_cap=64) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n, hist_cap=64, sampler=kind)
    assert ch.sampler == kind
    K, cells = pack_models(models, Kcap=ch.KC)
    ch.set_models(K, cells)
    out = ch.run(n_iter, recs=recs, trace=True)
    _check_replay(ch, out, runs, od, hist_n=20)
    assert ch.verify() == (0, 0.0, 0.0)
    st = ch.state(want_owners=True)
    own = st["owners"]
    assert own.shape == (n, ctx.P) and own.min() >= 0 and (own.max(1) < st["K"]).all()
    for c in range(n):  # the maintained owners are those of a fresh forward model
        k = st["K"][c]
        ref = ctx.evaluate_batch(np.array([k], np.int32), st["cells"][c:c + 1, :, :k].copy(), want_owners=True)["owners"][0]
        assert np.array_equal(own[c], ref)
    ch.close(); ctx.close()


def test_wide_sampler_equals_resident_sampler(tonga):
    """Same seed, same global chain ids: the three samplers give bit-identical chains (proposals, decisions, phi, history)."""
    import copy
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 500.0, 200.0, 10.0
    ctx = Context(ds, p)
    res = {}
    for kind in ("resident", "wide", "streamed"):
        ch = Chains(ctx, 12, chain_id0=5, seed=99, sampler=kind)
        ch.build_starting()
        out = ch.run(250, record=True, trace=True)
        out2 = ch.run(250, record=True, trace=True)  # a second call continues the same streams
        res[kind] = (out, out2, ch.state(want_owners=True), ch.history(), ch.stats())
        assert ch.verify() == (0, 0.0, 0.0)
        ch.close()
    assert np.array_equal(res["resident"][2]["owners"], res["streamed"][2]["owners"])
    _same_chains(res["resident"], res["wide"])
    _same_chains(res["resident"], res["streamed"])
    ctx.close()


def _same_chains(ra, rb):
    (a1, a2, sa, ha, ta), (b1, b2, sb, hb, tb) = ra, rb
    for x, y in ((a1, b1), (a2, b2)):
        assert x["recs"].tobytes() == y["recs"].tobytes()
        assert np.array_equal(x["accept"], y["accept"]) and np.array_equal(x["K"], y["K"])
        assert x["phi"].tobytes() == y["phi"].tobytes(), "phi must be bit-identical (same canonical summation orders)"
    assert np.array_equal(sa["K"], sb["K"]) and sa["phi"].tobytes() == sb["phi"].tobytes() and sa["ptS"].tobytes() == sb["ptS"].tobytes()
    for c in range(12):
        k = sa["K"][c]
        assert np.array_equal(sa["cells"][c, :, :k], sb["cells"][c, :, :k])
    assert np.array_equal(ha["n_hist"], hb["n_hist"]) and (ha["n_hist"] == 30).all()
    for key in ("K", "iter", "action", "accept", "next_action"):
        assert np.array_equal(ha[key], hb[key]), key
    assert ha["phi"].tobytes() == hb["phi"].tobytes() and ha["ptS"].tobytes() == hb["ptS"].tobytes()
    assert ta[0] == tb[0] and np.array_equal(ta[1], tb[1])


@pytest.mark.parametrize("kind", ["auto", "wide"])
def test_wide_sampler_beyond_126_cells(tonga, kind):
    """max_cells = 400 (the resident sampler stops at 126): auto selects the streamed sampler; replay of the oracle's chain."""
    import copy
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = tonga
    p = copy.copy(p0)
    p.max_cells, p.min_cells = 400, 150
    p.n_iter, p.burn_in, p.keep_each = 200.0, 100.0, 20.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(77)
    n, n_iter = 3, 200
    models = _start_models(rng, n, ds, kmin=150, kmax=400)  # includes births refused at max_cells / deaths at min_cells
    models[0] = random_model(rng, 400, box_of(ds))
    models[1] = random_model(rng, 150, box_of(ds))
    runs = [_oracle_generate(op, od, models[c], n_iter, 3000 + c, hist_cap=16) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n, hist_cap=16, sampler=kind)
    assert ch.sampler == ("streamed" if kind == "auto" else kind) and ch.KC == 400
    K, cells = pack_models(models, Kcap=ch.KC)
    ch.set_models(K, cells)
    out = ch.run(n_iter, recs=recs, trace=True)
    _check_replay(ch, out, runs, od, hist_n=5)
    assert ch.verify() == (0, 0.0, 0.0)
    ch.close(); ctx.close()


@pytest.mark.parametrize("kind", ["auto", "wide"])
def test_wide_sampler_large_ray_set(kind):
    """A ray set too large for shared memory (config-3 shape scaled down: 2000 rays, ~4e5 points, up to 300 nuclei)."""
    import oracle as O
    from tonga_b200.api import Chains, Context, pack_models
    from tonga_b200.data import synthetic_rays
    from tonga_b200.structs import parameters
    p = parameters()
    p.max_cells, p.min_cells = 300, 5
    p.n_iter, p.burn_in, p.keep_each = 40.0, 20.0, 10.0
    ds = synthetic_rays(2000, seed=5, p=p)
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(8)
    n, n_iter = 2, 40
    models = [random_model(rng, k, box_of(ds)) for k in (40, 250)]
    runs = [_oracle_generate(op, od, models[c], n_iter, 4000 + c, hist_cap=8) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n, hist_cap=8, sampler=kind)
    assert ch.sampler == ("streamed" if kind == "auto" else kind)
    K, cells = pack_models(models, Kcap=ch.KC)
    ch.set_models(K, cells)
    out = ch.run(n_iter, recs=recs, trace=True)
    _check_replay(ch, out, runs, od, hist_n=2)
    assert ch.verify() == (0, 0.0, 0.0)
    ch.close(); ctx.close()


@pytest.mark.parametrize("kind", ["resident", "wide", "streamed"])
def test_debug_prior_chains(tonga, kind):
    """debug_prior = 1 (MCsub.jl:128-136, phi == 1): both samplers follow the oracle's prior-sampling chain; hierarchical
    noise (action 5, extension) rides along."""
    import copy
    from tonga_b200.api import Chains, Context, pack_models
    ds, p0 = tonga
    p = copy.copy(p0)
    p.debug_prior = 1
    p.n_iter, p.burn_in, p.keep_each = 400.0, 100.0, 10.0
    ctx = Context(ds, p)
    op, od = _orc(ds, p)
    rng = np.random.default_rng(91)
    n, n_iter = 4, 400
    models = _start_models(rng, n, ds)
    runs = [_oracle_generate(op, od, models[c], n_iter, 5000 + c) for c in range(n)]
    recs = np.stack([r.recs for r, _ in runs])
    ch = Chains(ctx, n, hist_cap=0, sampler=kind)
    K, cells = pack_models(models, Kcap=ch.KC)
    ch.set_models(K, cells)
    out = ch.run(n_iter, recs=recs, trace=True)
    for c, (r, mb) in enumerate(runs):
        assert np.array_equal(out["accept"][c], r.accept) and np.array_equal(out["K"][c], r.K)
        assert (out["phi"][c] == 1.0).all() and (r.phi == 1.0).all()
        k = mb.K
        for a, ref in enumerate(mb.cells()):
            assert np.array_equal(ch.state()["cells"][c, a, :k], ref)
    ch.close(); ctx.close()


def test_wide_sampler_sigma_move(tonga):
    """n_actions = 5 (hierarchical noise, extension): wide == streamed == resident bit for bit in generate mode."""
    import copy
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    ctx = Context(ds, p, n_actions=5)
    res = {}
    for kind in ("resident", "wide", "streamed"):
        ch = Chains(ctx, 6, seed=3, hist_cap=0, sampler=kind)
        ch.build_starting()
        res[kind] = (ch.run(300, record=True, trace=True), ch.state())
        ch.close()
    a, sa = res["resident"]
    assert (a["recs"]["action"] == 5).any()
    for kind in ("wide", "streamed"):
        b, sb = res[kind]
        assert a["recs"].tobytes() == b["recs"].tobytes() and np.array_equal(a["accept"], b["accept"])
        assert a["phi"].tobytes() == b["phi"].tobytes() and sa["noise"].tobytes() == sb["noise"].tobytes()
    ctx.close()


# ------------------------------------------------------------------------------------------------ device-side ingest (N4)
@pytest.mark.parametrize("which", ["tonga", "synthetic"])
def test_device_ingest_matches_host_flatten(tonga, which):
    """tonga_create_from_points (rayl, rayu, dt, flatten on the GPU; load_data_Tonga.jl:66-69) == tonga_create fed with the host
    arithmetic of data.ray_lengths: same sizes, same offsets, bit-identical owners / t* / phi, bit-identical chains."""
    from tonga_b200.api import Chains, Context
    from tonga_b200.data import synthetic_rays
    from tonga_b200.structs import parameters
    if which == "tonga":
        ds, p = tonga
    else:
        p = parameters()
        ds = synthetic_rays(700, seed=12, pts=(2, 90), p=p)  # ragged, including 2-point rays
    a, b = Context(ds, p), Context(ds, p, device_ingest=True)
    assert (a.R, a.P, a.S, a.Ppad) == (b.R, b.P, b.S, b.Ppad)
    assert np.array_equal(a.ray_offsets(), b.ray_offsets())
    rng = np.random.default_rng(4)
    for K in (1, 7, 60):
        mdl = random_model(rng, K, box_of(ds))
        Kb, cells = np.array([K], np.int32), np.stack(mdl)[None]
        ra, rb = a.evaluate_batch(Kb, cells, want_owners=True), b.evaluate_batch(Kb, cells, want_owners=True)
        assert np.array_equal(ra["owners"], rb["owners"])
        assert ra["ptS"].tobytes() == rb["ptS"].tobytes() and ra["phi"].tobytes() == rb["phi"].tobytes()
    outs = []
    for ctx in (a, b):
        ch = Chains(ctx, 4, seed=2, hist_cap=0)
        ch.build_starting()
        outs.append(ch.run(200, record=True, trace=True))
        assert ch.verify() == (0, 0.0, 0.0)
        ch.close()
    assert outs[0]["recs"].tobytes() == outs[1]["recs"].tobytes() and outs[0]["phi"].tobytes() == outs[1]["phi"].tobytes()
    a.close(); b.close()


# ------------------------------------------------------------------------------------------------ checkpoint / resume (N3)
@pytest.mark.parametrize("kind", ["resident", "streamed"])
def test_checkpoint_resume_is_bit_identical(tonga, tmp_path, kind):
    """TD_inversion_function.jl:40-67 / :282-294: a batch saved mid-run and resumed in a new object continues exactly like the
    uninterrupted run (same Philox streams, thinning phase, counters, history incl. the pending next_action)."""
    import copy
    from tonga_b200 import checkpoint as ckpt
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 300.0, 100.0, 7.0
    ctx = Context(ds, p)
    a = Chains(ctx, 6, chain_id0=2, seed=77, sampler=kind)
    a.build_starting()
    ta = a.run(300, trace=True)
    b = Chains(ctx, 6, chain_id0=2, seed=77, sampler=kind)
    b.build_starting()
    tb1 = b.run(143, trace=True)  # stops right after a kept model (iter 100 + 6*7 = 142 is kept at model_num 43? any phase must work)
    ckpt.save_checkpoint(str(tmp_path / "batch.npz"), b)
    b.close()
    ck = ckpt.load_checkpoint(str(tmp_path / "batch.npz"))
    assert int(ck["iter"]) == 143 and bool(ck["burnin"])
    c = ckpt.resume(ctx, ck)
    assert c.sampler == kind
    tb2 = c.run(157, trace=True)
    for key in ("accept", "phi", "K"):
        assert np.concatenate([tb1[key], tb2[key]], 1).tobytes() == ta[key].tobytes(), key
    sa, sc = a.state(), c.state()
    assert sa["phi"].tobytes() == sc["phi"].tobytes() and sa["ptS"].tobytes() == sc["ptS"].tobytes() and np.array_equal(sa["K"], sc["K"])
    ha, hc = a.history(), c.history()
    for key in ha:
        if key == "cells":
            for i in range(6):
                for j in range(ha["n_hist"][i]):
                    k = ha["K"][i, j]
                    assert np.array_equal(ha["cells"][i, j, :, :k], hc["cells"][i, j, :, :k])
        else:
            assert ha[key].tobytes() == hc[key].tobytes(), key
    assert a.stats()[0] == c.stats()[0] == 300 and np.array_equal(a.stats()[1], c.stats()[1])
    assert c.verify() == (0, 0.0, 0.0)
    # model.jld-equivalent export
    n = ckpt.export_model_hist(str(tmp_path / "model_hist.bin"), hc, likelihood=1.5)
    back = ckpt.import_model_hist(str(tmp_path / "model_hist.bin"))
    assert n == int(hc["n_hist"].sum()) and len(back) == 6
    m0 = back[3][5]
    assert m0["K"] == hc["K"][3, 5] and m0["phi"] == hc["phi"][3, 5] and np.array_equal(m0["ptS"], hc["ptS"][3, 5])
    assert np.array_equal(m0["cells"], hc["cells"][3, 5, :, :m0["K"]])
    a.close(); c.close(); ctx.close()


def test_device_normal_deviates_are_standard_normal(tonga):
    """The device draws its Gaussian proposal steps by inversion (AS 241, proposal.cuh): the first-iteration change / move steps
    of many chains, normalised by their proposal widths, must be N(0, 1) (Kolmogorov-Smirnov, mean, variance, tails)."""
    from scipy import stats
    from tonga_b200.api import Chains, Context
    ds, p = tonga
    ctx = Context(ds, p)
    n = 16384
    ch = Chains(ctx, n, seed=123, hist_cap=0)
    ch.build_starting()
    st0 = ch.state(want_ptS=False)
    recs = ch.run(1, record=True)["recs"][:, 0]
    sig_zeta = p.zeta_scale * p.sig / 100
    idx = recs["idx"]
    chg = recs["action"] == 3
    z = (recs["zeta"][chg] - st0["cells"][chg, 3, idx[chg]]) / sig_zeta
    mv = recs["action"] == 4
    box = box_of(ds)
    zs = [z]
    for a, key, lo, hi in ((0, "x", box[0], box[1]), (1, "y", box[2], box[3]), (2, "z", box[4], box[5])):
        zs.append((recs[key][mv] - st0["cells"][mv, a, idx[mv]]) / ((p.sig / 100) * (hi - lo)))
    for k, v in enumerate(zs):
        assert len(v) > 3000
        assert abs(v.mean()) < 4 / np.sqrt(len(v)) and abs(v.var() - 1) < 4 * np.sqrt(2 / len(v))
        assert stats.kstest(v, "norm").pvalue > 1e-3, f"component {k}"
        assert 0.03 < (np.abs(v) > 2).mean() < 0.06  # P(|Z| > 2) = 0.0455: the tail branch of the inversion
    # the three move components are independent draws
    c = np.corrcoef(np.stack(zs[1:]))
    assert np.abs(c[np.triu_indices(3, 1)]).max() < 5 / np.sqrt(len(zs[1]))
    ch.close(); ctx.close()


@pytest.mark.parametrize("kind", ["resident", "streamed"])
def test_host_resident_history_equals_device_history(tonga, kind):
    """TONGA_HISTORY_ON_HOST: the kernels store the kept models into mapped host memory while they run; same history, no copy."""
    import copy
    from tonga_b200.api import Chains, Context
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 400.0, 100.0, 10.0
    ctx = Context(ds, p)
    hs = []
    for host in (False, True):
        ch = Chains(ctx, 5, chain_id0=1, seed=9, sampler=kind, host_history=host)
        ch.build_starting()
        ch.run(250); ch.run(150)
        h = ch.history()
        hs.append({k: np.array(v) for k, v in h.items()})
        if host:
            s1, s2, cnt = ch.raster([100.0, 500.0], [0.0, 100.0], [50.0, 300.0])
            assert cnt == int(h["n_hist"].sum()) and np.isfinite(s1).all()
        ch.close()
    a, b = hs
    assert (a["n_hist"] == 30).all() and np.array_equal(a["n_hist"], b["n_hist"])
    for key in ("K", "phi", "ptS", "iter", "action", "accept", "next_action"):
        assert a[key].tobytes() == b[key].tobytes(), key
    for c in range(5):
        for j in range(30):
            k = a["K"][c, j]
            assert np.array_equal(a["cells"][c, j, :, :k], b["cells"][c, j, :, :k])
    ctx.close()


def test_streamed_screening_equals_exact(tonga):
    """Streamed sampler: FP32 screening + exact recheck (default) == every comparison in exact FP64 (set_exact_only), bit for bit."""
    from tonga_b200.api import Chains, Context
    ds, p = tonga
    ctx = Context(ds, p)
    outs = []
    for exact in (False, True):
        ch = Chains(ctx, 6, seed=31, hist_cap=0, sampler="streamed")
        ch.set_exact_only(exact)
        ch.build_starting()
        out = ch.run(300, record=True, trace=True)
        outs.append((out, ch.state(want_owners=True)))
        assert ch.verify() == (0, 0.0, 0.0)
        ch.close()
    (a, sa), (b, sb) = outs
    assert a["recs"].tobytes() == b["recs"].tobytes() and np.array_equal(a["accept"], b["accept"]) and a["phi"].tobytes() == b["phi"].tobytes()
    assert np.array_equal(sa["owners"], sb["owners"]) and sa["ptS"].tobytes() == sb["ptS"].tobytes()
    ctx.close()


@pytest.mark.gpu
def test_device_tempering_sweep_matches_the_numpy_mirror(tonga):
    """Extension (BASELINE config 5): the swap sweep decided on the device from the replicas' own phi / noise equals the
    NumPy mirror (same Philox counters, same order of operations) -- the property that lets every rank of a cross-GPU ladder
    take identical decisions.  Both ways of calling it: the batch's own arrays, and external all-gathered device arrays."""
    import copy
    import torch
    from tonga_b200 import api
    from tonga_b200.tempering import energy, geometric_ladder, swap_step
    ds, p0 = tonga
    p = copy.copy(p0)
    p.max_sig = 2.0
    ctx = api.Context(ds, p, n_actions=5)
    T, L = 16, 8
    ch = api.Chains(ctx, T * L, seed=3, hist_cap=0)
    ch.build_starting()
    beta = np.tile(geometric_ladder(T, 50.0), L)
    ch.set_beta(beta)
    acc = att = 0
    for step in range(6):
        ch.run(50)
        st = ch.state(want_ptS=False)
        ref, a, t = swap_step(energy(st["phi"], st["noise"], ctx.R), beta, T, step, seed=21)
        acc += a; att += t
        if step % 2 == 0:
            ch.temper_swap(T, step, seed=21)
        else:  # external arrays (what a multi-rank caller passes after its all-gather): here a copy of the batch's own
            phi_d, noise_d, beta_d = (torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (st["phi"], st["noise"], beta))
            torch.cuda.synchronize()
            ch.temper_swap(T, step, seed=21, n_all=T * L, phi_all=phi_d.data_ptr(), noise_all=noise_d.data_ptr(), beta_all=beta_d.data_ptr(), offset=0)
            ctx.synchronize()
            assert np.array_equal(beta_d.cpu().numpy(), ref)
        beta = ch.get_beta()
        assert np.array_equal(beta, ref), step
    assert ch.temper_stats() == (acc, att) and 0 < acc < att
    ck = ch.checkpoint()
    assert np.array_equal(ck["beta"], beta)  # tempered batches resume with their temperatures
    ch.close(); ctx.close()


@pytest.mark.gpu
def test_long_runs_are_cut_into_launches_without_a_trace(tonga, monkeypatch):
    """The resident sampler pre-generates the raw draws of a launch and cuts long runs into pieces (256 MB of draws at most).
    The pieces must be invisible: records, traces, history, counters and the final state equal the uncut run's bit for bit."""
    from tonga_b200 import api
    ds, p0 = tonga
    import copy
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 120.0, 40.0, 7.0
    ctx = api.Context(ds, p)
    res = []
    for cut in (None, "37"):
        if cut is None:
            monkeypatch.delenv("TONGA_RAW_ITERS", raising=False)
        else:
            monkeypatch.setenv("TONGA_RAW_ITERS", cut)
        ch = api.Chains(ctx, 6, seed=77, hist_cap=16)
        ch.build_starting()
        out = ch.run(120, record=True, trace=True)
        res.append((out, ch.state(), ch.history(), ch.stats(), ch.verify()))
        ch.close()
    (o0, s0, h0, t0, v0), (o1, s1, h1, t1, v1) = res
    assert v0 == v1 == (0, 0.0, 0.0)
    assert o0["recs"].tobytes() == o1["recs"].tobytes() and np.array_equal(o0["accept"], o1["accept"]) and o0["phi"].tobytes() == o1["phi"].tobytes()
    assert np.array_equal(o0["K"], o1["K"]) and t0[0] == t1[0] and np.array_equal(t0[1], t1[1])
    for k in ("K", "phi", "ptS"):
        assert np.array_equal(s0[k], s1[k]), k
    assert np.array_equal(h0["n_hist"], h1["n_hist"]) and (h0["n_hist"] == 11).all()
    for k in ("K", "phi", "ptS", "iter", "action", "accept", "next_action"):
        assert np.array_equal(h0[k], h1[k]), k
    ctx.close()


def _run_shards_in_threads(chs, n_iter, **kw):
    """run() of every shard in its own host thread (the accept kernel of a shard waits for the other shards' candidate passes)."""
    import threading
    outs, errs = [None] * len(chs), [None] * len(chs)

    def work(i):
        try:
            outs[i] = chs[i].run(n_iter, **kw)
        except Exception as e:  # noqa: BLE001
            errs[i] = e
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(chs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_ray_sharded_streamed_sampler_equals_unsharded(tonga, monkeypatch, world):
    """BASELINE config 3's multi-GPU mode on one GPU: `world` ray shards of a STREAMED batch (own context and stream each, the
    'peer' exchange blocks are plain device pointers here) run concurrently and exchange (t*, misfit term) of their rays per
    proposal.  Every shard must reproduce the unsharded run bit for bit: proposals, decisions, phi, K, t*, history, counters;
    the shards' own owners together are the unsharded owners."""
    import copy
    from tonga_b200.api import Chains, Context
    monkeypatch.setenv("TONGA_STREAM_TILE", "1024")   # 381-ray Tonga: 17 tiles to share out
    monkeypatch.setenv("TONGA_SHARD_TIMEOUT_MS", "20000")
    ds, p0 = tonga
    p = copy.copy(p0)
    p.n_iter, p.burn_in, p.keep_each = 120.0, 40.0, 10.0
    n = 6
    ctx = Context(ds, p)
    ref = Chains(ctx, n, chain_id0=3, seed=2024, sampler="streamed")
    ref.build_starting()
    o_ref = [ref.run(60, record=True, trace=True), ref.run(60, record=True, trace=True)]
    s_ref, h_ref, t_ref = ref.state(want_owners=True), ref.history(), ref.stats()
    assert ref.verify() == (0, 0.0, 0.0)
    ctxs = [Context(ds, p) for _ in range(world)]
    chs = [Chains(c, n, chain_id0=3, seed=2024, sampler="streamed") for c in ctxs]
    bases = [ch.shard_init(r, world)[0] for r, ch in enumerate(chs)]
    for ch in chs:
        ch.shard_connect(bases)
        ch.build_starting()
    infos = [ch.shard_info() for ch in chs]
    assert infos[0]["point0"] == 0 and infos[-1]["point1"] == ctx.P and all(a["point1"] == b["point0"] for a, b in zip(infos, infos[1:]))
    o1 = _run_shards_in_threads(chs, 60, record=True, trace=True)
    o2 = _run_shards_in_threads(chs, 60, record=True, trace=True)
    owners = np.full_like(s_ref["owners"], -2)
    for r, ch in enumerate(chs):
        for o, orf in ((o1[r], o_ref[0]), (o2[r], o_ref[1])):
            assert o["recs"].tobytes() == orf["recs"].tobytes() and np.array_equal(o["accept"], orf["accept"]) and np.array_equal(o["K"], orf["K"])
            assert o["phi"].tobytes() == orf["phi"].tobytes(), "phi must be bit-identical: the terms of all rays are summed in the canonical order on every rank"
        s, h, t = ch.state(want_owners=True), ch.history(), ch.stats()
        assert np.array_equal(s["K"], s_ref["K"]) and s["phi"].tobytes() == s_ref["phi"].tobytes() and s["ptS"].tobytes() == s_ref["ptS"].tobytes()
        assert s["cells"].tobytes() == s_ref["cells"].tobytes()
        for key in ("n_hist", "K", "iter", "action", "accept", "next_action"):
            assert np.array_equal(h[key], h_ref[key]), key
        assert h["phi"].tobytes() == h_ref["phi"].tobytes() and h["ptS"].tobytes() == h_ref["ptS"].tobytes() and (h["n_hist"] == 8).all()
        assert t[0] == t_ref[0] and np.array_equal(t[1], t_ref[1])
        assert ch.verify() == (0, 0.0, 0.0)
        mine = s["owners"] != -2
        assert not (mine & (owners != -2)).any(), "every ray point belongs to exactly one shard"
        owners[mine] = s["owners"][mine]
    assert np.array_equal(owners, s_ref["owners"])
    for ch in chs:
        ch.close()
    for c in ctxs:
        c.close()
    ref.close(); ctx.close()


@pytest.mark.gpu
def test_ray_shard_peer_timeout_is_an_error_not_a_hang(tonga, monkeypatch):
    """A shard whose peer never runs must come back with TONGA_ERR_PEER after the bounded wait."""
    from tonga_b200._lib import TongaError
    from tonga_b200.api import Chains, Context
    monkeypatch.setenv("TONGA_SHARD_TIMEOUT_MS", "300")
    ds, p = tonga
    ctx = Context(ds, p)
    a, b = Chains(ctx, 2, seed=1, sampler="streamed", hist_cap=0), Chains(ctx, 2, seed=1, sampler="streamed", hist_cap=0)
    bases = [a.shard_init(0, 2)[0], b.shard_init(1, 2)[0]]
    with pytest.raises(TongaError):
        a.build_starting(); a.run(1)          # not connected yet
    a.shard_connect(bases); b.shard_connect(bases)
    a.build_starting()
    with pytest.raises(TongaError) as e:
        a.run(2)                              # b never publishes
    assert e.value.code == -6
    a.close(); b.close(); ctx.close()


@pytest.mark.gpu
def test_ray_shards_across_two_gpus_over_cuda_ipc():
    """The real thing when the box has >= 2 GPUs: one process per GPU (torchrun), exchange blocks mapped over CUDA IPC, P2P
    stores from the candidate pass over NVLink; tools/shard_check.py asserts bit-identity with the unsharded run on every rank."""
    import json
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the 1-GPU multi-shard test above covers the protocol)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tools", "shard_check.py"), "4000", "8", "100", "30"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["parity_all_ranks"] and line["verify_all_ranks"] and line["world"] == 2
