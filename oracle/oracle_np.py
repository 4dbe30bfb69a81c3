"""NumPy twin of the C oracle -- an INDEPENDENT restatement written from the same cited reference lines.

CPU ORACLE -- test infrastructure, NOT the product (same rules as tonga_oracle.h).  Its purpose is that a single
restatement bug cannot pin itself: tests/test_oracle.py requires the C oracle and this twin to agree
(owners bit-exact, t*/phi to 1e-12) on the shipped 381-ray geometry and on random ragged ray sets.

All file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np


def v_nearest(x, y, z, mx, my, mz, mv):
    """MCsub.jl:247-263: strict `<`, first index wins, mdist starts at 1e9, v starts at 0.0."""
    v, mdist, best = 0.0, 1e9, -1
    for i in range(len(mx)):
        dx, dy, dz = mx[i] - x, my[i] - y, mz[i] - z
        distance = dx * dx + dy * dy + dz * dz  # :254 (x^2 is x*x in Julia; ** would call pow())
        if distance < mdist:  # :255
            mdist, v, best = distance, mv[i], i
    return v, best


def nearest_many(X, Y, Z, mx, my, mz, mv):
    """Vectorised v_nearest for many points: same arithmetic ((dx*dx + dy*dy) + dz*dz in float64, elementwise,
    no fused ops in NumPy), np.argmin returns the FIRST minimum (= lowest index on ties), 1e9 cut-off applied after."""
    X, Y, Z = (np.asarray(a, dtype=np.float64) for a in (X, Y, Z))
    if len(X) == 0:
        return np.zeros(0), np.zeros(0, np.int32)
    dx = mx[None, :] - X[:, None]
    dy = my[None, :] - Y[:, None]
    dz = mz[None, :] - Z[:, None]
    d = (dx * dx + dy * dy) + dz * dz
    idx = np.argmin(d, axis=1).astype(np.int32)
    dmin = d[np.arange(len(X)), idx]
    ok = dmin < 1e9  # :250,:255 -- a nucleus only wins if strictly closer than the 1e9 start value
    val = np.where(ok, np.asarray(mv)[idx], 0.0)
    return val, np.where(ok, idx, -1).astype(np.int32)


def interpolation(mx, my, mz, mv, X, Y, Z):
    """MCsub.jl:306-336 (interp_style 1)."""
    X = np.atleast_1d(np.asarray(X, dtype=np.float64))
    nan = np.flatnonzero(np.isnan(X))
    npoints = int(nan[0]) if len(nan) else len(X)  # :312-316
    Y = np.atleast_1d(np.asarray(Y, dtype=np.float64))
    Z = np.atleast_1d(np.asarray(Z, dtype=np.float64))
    if len(Y) == 1:
        Y = Y * np.ones(npoints)  # :317-319
    if len(Z) == 1:
        Z = Z * np.ones(npoints)  # :320-322
    return nearest_many(X[:npoints], Y[:npoints], Z[:npoints], np.asarray(mx), np.asarray(my), np.asarray(mz), np.asarray(mv))


def misfit(ptS, tS, allSig, noise=1.0):
    """MCsub.jl:169-182 -> (phi, likelihood).  Pinned bit for bit by the reference's own model.jld (tests/test_oracle.py)."""
    n = len(tS)
    C = 0.0
    lk = 0.0
    for k in range(n):  # :169-172
        sg = noise * allSig[k]
        df = (ptS - tS)[k]
        C += df * df * 1.0 / (sg * sg)
        lk += -math.log(sg * math.sqrt(2 * math.pi)) * n  # :179 (the second line, :180, is discarded: SURVEY F5)
    return C, lk


def evaluate(rayX, rayY, rayZ, rayL, rayU, tS, allSig, mx, my, mz, mv, noise=1.0, debug_prior=0):
    """MCsub.jl:123-185 -> dict(ptS, phi, likelihood, owners[m,R])."""
    m, n = rayX.shape  # :138
    if debug_prior == 1:
        return dict(ptS=None, phi=1.0, likelihood=1.0, owners=None)  # :128-136
    ptS = np.zeros(n)
    owners = np.full((m, n), -1, np.int32)
    for i in range(n):  # :142
        zeta0, idx = interpolation(mx, my, mz, mv, rayX[:, i], rayY[:, i], rayZ[:, i])  # :143
        owners[:len(idx), i] = idx
        rayzeta = 0.5 * (zeta0[:-1] + zeta0[1:])  # :147
        rayl = rayL[:, i]
        nan = np.flatnonzero(np.isnan(rayl))  # :150
        nseg = int(nan[0]) if len(nan) else len(rayl)
        if nseg != len(rayzeta):
            raise ValueError("DimensionMismatch (segments vs points)")
        terms = rayl[:nseg] * rayU[:nseg, i] * (rayzeta / 1000)  # :153/:159
        s = 0.0
        for t in terms:  # left-to-right (Julia's own order is unspecified; see tonga_oracle.h)
            s += t
        ptS[i] = s
    C, lk = misfit(ptS, tS, allSig, noise)
    return dict(ptS=ptS, phi=C, likelihood=lk, owners=owners)


def chain_replay(params, data, model, recs, iter0=1):
    """TD_inversion_function.jl:70-274 replaying recorded proposals (uniform prior only -- the reference default).

    params: dict(xmin..zmax, sig, zeta_scale, min_cells, max_cells, burn_in, keep_each)
    data  : dict(rayX, rayY, rayZ, rayL, rayU, tS, allSig)
    model : dict(x, y, z, zeta) (lists / arrays) -- evaluated here first
    -> (accept[n], phi[n], K[n], final model dict)
    """
    ev = lambda mdl: evaluate(data["rayX"], data["rayY"], data["rayZ"], data["rayL"], data["rayU"], data["tS"],
                              data["allSig"], *(np.asarray(mdl[k], dtype=np.float64) for k in ("x", "y", "z", "zeta")))
    model = {k: list(np.asarray(model[k], dtype=np.float64)) for k in ("x", "y", "z", "zeta")}
    model["phi"] = ev(model)["phi"]
    sig_zeta = params["zeta_scale"] * params["sig"] / 100  # :22
    acc, phis, Ks = [], [], []
    for rec in recs:
        action = int(rec["action"])
        K = len(model["x"])
        accepted = 0
        alpha, valid = 0.0, 0
        mn = None
        if action == 1:  # :76-125
            if K < params["max_cells"]:
                czeta, _ = v_nearest(rec["x"], rec["y"], rec["z"], model["x"], model["y"], model["z"], model["zeta"])
                zetanew = float(rec["zeta"])
                mn = {k: list(model[k]) for k in ("x", "y", "z", "zeta")}
                mn["x"].append(float(rec["x"])); mn["y"].append(float(rec["y"])); mn["z"].append(float(rec["z"])); mn["zeta"].append(zetanew)
                if zetanew > 0 and zetanew < params["zeta_scale"]:
                    mn["phi"] = ev(mn)["phi"]; valid = 1
                    alpha = ((K) / (K + 1)) * ((sig_zeta * math.sqrt(2 * math.pi)) / (params["zeta_scale"])) * \
                        math.exp(((czeta - zetanew) * (czeta - zetanew)) / (2 * (sig_zeta * sig_zeta)) - (mn["phi"] - model["phi"]) / 2)  # :96-97
                    alpha = min(1, alpha)
                else:
                    valid = 0
                if rec["u"] < alpha and valid == 1:
                    model, accepted = mn, 1
        elif action == 2:  # :126-181
            if K > params["min_cells"]:
                kill = int(rec["idx"])
                mn = {k: list(model[k]) for k in ("x", "y", "z", "zeta")}
                for k in ("x", "y", "z", "zeta"):
                    del mn[k][kill]
                mn["phi"] = ev(mn)["phi"]; valid = 1
                zetanew, _ = v_nearest(model["x"][kill], model["y"][kill], model["z"][kill], mn["x"], mn["y"], mn["z"], mn["zeta"])  # :146
                alpha = ((K) / (K - 1)) * ((params["zeta_scale"]) / (sig_zeta * math.sqrt(2 * math.pi))) * \
                    math.exp(-((model["zeta"][kill] - zetanew) * (model["zeta"][kill] - zetanew)) / (2 * (sig_zeta * sig_zeta)) - (mn["phi"] - model["phi"]) / 2)  # :151-152
                alpha = min(1, alpha)
                if rec["u"] < alpha and valid == 1:
                    model, accepted = mn, 1
        elif action == 3:  # :183-218
            ch = int(rec["idx"])
            mn = {k: list(model[k]) for k in ("x", "y", "z", "zeta")}
            mn["zeta"][ch] = float(rec["zeta"])
            mn["phi"] = ev(mn)["phi"]; valid = 1
            if mn["zeta"][ch] > 0 and mn["zeta"][ch] < params["zeta_scale"]:
                alpha = min(1, math.exp(-(mn["phi"] - model["phi"]) / 2))  # :196-197
            else:
                valid = 0
            if rec["u"] < alpha and valid == 1:
                model, accepted = mn, 1
        elif action == 4:  # :220-251
            mv = int(rec["idx"])
            mn = {k: list(model[k]) for k in ("x", "y", "z", "zeta")}
            if (params["xmin"] <= rec["x"] <= params["xmax"] and params["ymin"] <= rec["y"] <= params["ymax"]
                    and params["zmin"] <= rec["z"] <= params["zmax"]):  # :230-232
                mn["x"][mv], mn["y"][mv], mn["z"][mv] = float(rec["x"]), float(rec["y"]), float(rec["z"])
                mn["phi"] = ev(mn)["phi"]; valid = 1
                alpha = min(1, math.exp(-(mn["phi"] - model["phi"]) / 2))  # :241-242
            else:
                valid = 0
            if rec["u"] < alpha and valid == 1:
                model, accepted = mn, 1
        else:
            raise ValueError("numpy twin replays actions 1-4 only")
        acc.append(accepted); phis.append(model["phi"]); Ks.append(len(model["x"]))
    return np.array(acc, np.int8), np.array(phis), np.array(Ks, np.int32), model
