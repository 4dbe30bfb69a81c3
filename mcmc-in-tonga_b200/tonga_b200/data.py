"""Ray-set construction for the hot path: the arithmetic of the reference's loader, not its file I/O.

  ray_lengths()      <- load_data_Tonga.jl:66-69  (rayl = |dp| per segment, rayu = mean endpoint slowness)
  domain_from_stations() <- load_data_Tonga.jl:41-49 (box = station extent +- buffer, node spacing)
  load_tonga381()    -> DataStruct for the shipped 381-ray set (tonga_b200/datasets/tonga381.npz, made by tests/golden/make_fixtures.py)
  synthetic_rays()   -> BASELINE.json config 3 style synthetic ray sets

Documented substitutions for the shipped data: the shipped raypaths file has no slowness `u`,
so U(z) = 1/Vp_ak135(z) is synthesised from Data/ak135f.txt (linear in depth, deeper branch at
discontinuities; load_3Dvel.jl:27,32 uses P slowness); `aveatten` is absent and only feeds a dead variable
(MCsub.jl:162), so ones are passed; station positions are the ray end points (the shipped rays are in an
older local frame than the current lonlat2xy origin).
"""
from __future__ import annotations

import os

import numpy as np

from .structs import DataStruct, StepRangeLen, parameters

_DATASETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datasets")  # input data of the package (not test fixtures)


def ak135_slowness(z: np.ndarray, table: np.ndarray) -> np.ndarray:
    """U(z) = 1 / Vp(z); `table` rows are (depth km, Vp, Vs).  NaN in -> NaN out."""
    depth, vp = table[:, 0], table[:, 1]
    z = np.asarray(z, dtype=np.float64)
    out = np.full(z.shape, np.nan)
    ok = ~np.isnan(z)
    zz = np.clip(z[ok], depth[0], depth[-1])
    hi = np.searchsorted(depth, zz, side="right")  # first node strictly deeper -> deeper branch at jumps
    hi = np.clip(hi, 1, len(depth) - 1)
    lo = hi - 1
    span = depth[hi] - depth[lo]
    w = np.where(span > 0, (zz - depth[lo]) / np.where(span > 0, span, 1.0), 1.0)
    out[ok] = 1.0 / (vp[lo] + w * (vp[hi] - vp[lo]))
    return out


def ray_lengths(x: np.ndarray, y: np.ndarray, z: np.ndarray, U: np.ndarray):
    """load_data_Tonga.jl:66-69 on m x R arrays; NaN wherever an endpoint is padding."""
    with np.errstate(invalid="ignore"):
        dx, dy, dz = x[:-1] - x[1:], y[:-1] - y[1:], z[:-1] - z[1:]
        rayl = np.sqrt(dx * dx + dy * dy + dz * dz)
        rayu = 0.5 * (U[:-1] + U[1:])
    return np.asfortranarray(rayl), np.asfortranarray(rayu)


def pad_rays(npts: np.ndarray, px: np.ndarray, py: np.ndarray, pz: np.ndarray, m: int | None = None):
    """Flat points + points-per-ray -> the reference's NaN-padded m x R column-major matrices."""
    R = len(npts)
    m = int(npts.max()) if m is None else m
    off = np.concatenate([[0], np.cumsum(npts)])
    out = [np.full((m, R), np.nan, order="F") for _ in range(3)]
    for a, src in zip(out, (px, py, pz)):
        for i in range(R):
            a[:npts[i], i] = src[off[i]:off[i + 1]]
    return out


def domain_from_stations(dataX, dataY, p: parameters):
    """load_data_Tonga.jl:41-49."""
    minX, maxX = dataX.min() - p.buffer, dataX.max() + p.buffer
    minY, maxY = dataY.min() - p.buffer, dataY.max() + p.buffer
    return (StepRangeLen(minX, p.XYnodeSpacing, maxX), StepRangeLen(minY, p.XYnodeSpacing, maxY),
            StepRangeLen(p.min_depth, p.ZnodeSpacing, p.max_depth))


def make_datastruct(x, y, z, U, tS, sig, p: parameters, box=None) -> DataStruct:
    R = x.shape[1]
    rayl, rayu = ray_lengths(x, y, z, U)
    npts = (~np.isnan(x)).sum(0)
    last = np.maximum(npts - 1, 0)
    dataX, dataY = x[last, np.arange(R)], y[last, np.arange(R)]
    if box is None:
        xVec, yVec, zVec = domain_from_stations(dataX, dataY, p)
    else:
        xVec, yVec, zVec = box
    e = np.zeros((0, 0))
    return DataStruct(
        tS=np.ascontiguousarray(tS, dtype=np.float64), allaveatten=np.ones(R), allLats=np.zeros(R), allLons=np.zeros(R),
        allSig=np.ascontiguousarray(sig, dtype=np.float64), dataX=dataX, dataY=dataY, xVec=xVec, yVec=yVec, zVec=zVec,
        elonsX=e, elatsY=e, elons=e, elats=e, edep=e, coastX=e, coastY=e,
        rayX=np.asfortranarray(x), rayY=np.asfortranarray(y), rayZ=np.asfortranarray(z), rayL=rayl, rayU=rayu,
        U=np.asfortranarray(U))


def load_tonga381(path: str | None = None, p: parameters | None = None) -> DataStruct:
    """The shipped 381-ray Tonga set (BASELINE.json configs 1, 2, 4, 5)."""
    p = p or parameters()
    f = np.load(path or os.path.join(_DATASETS, "tonga381.npz"))
    x, y, z = pad_rays(f["npts"], f["px"], f["py"], f["pz"])
    U = ak135_slowness(z, f["ak135"])
    return make_datastruct(x, y, z, U, f["tStar"], f["error"], p)


_AK135_COARSE = np.array([  # fallback Vp(z) if no ak135 table is supplied (synthetic sets only)
    [0.0, 5.8], [20.0, 6.5], [35.0, 8.04], [120.0, 8.05], [210.0, 8.30], [410.0, 9.03], [410.0, 9.36],
    [660.0, 10.20], [660.0, 10.79], [760.0, 11.06]])


def load_warm_start(path: str | None = None) -> dict:
    """The warm start models of the headline benchmark (tools/make_warm_start.py): K[1024], cells[1024, 4, Kmax] -- the states
    of 1024 chains of the 381-ray inversion after 20 000 iterations.  bench.py starts both of its arms from them."""
    f = np.load(path or os.path.join(_DATASETS, "warm_start_1024.npz"))
    return {"K": np.ascontiguousarray(f["K"], dtype=np.int32), "cells": np.ascontiguousarray(f["cells"], dtype=np.float64)}


def synthetic_rays(R: int, seed: int = 3, pts=(150, 250), box=(1000.0, 1000.0, 660.0), n_true: int = 200,
                   zeta_scale: float = 50.0, jitter: float = 0.5, p: parameters | None = None,
                   ak135: np.ndarray | None = None) -> DataStruct:
    """straight rays source->receiver(z=0), equally spaced points + N(0, jitter) km,
    tS from a random `n_true`-nucleus model + N(0, sigma^2), sigma ~ U(0.04, 0.66)  (n_true = 0: noise only; fast)."""
    p = p or parameters()
    rng = np.random.default_rng(seed)
    bx, by, bz = box
    npts = rng.integers(pts[0], pts[1] + 1, size=R).astype(np.int32)
    m = int(npts.max())
    src = np.stack([rng.uniform(0, bx, R), rng.uniform(0, by, R), rng.uniform(50.0, bz, R)], 1)
    rcv = np.stack([rng.uniform(0, bx, R), rng.uniform(0, by, R), np.zeros(R)], 1)
    # all rays at once: parameter t = k / (n-1) for k < n, NaN beyond (point x ray, Fortran order like the reference)
    kk = np.arange(m)[:, None]
    valid = kk < npts[None, :]
    t = kk / np.maximum(npts[None, :] - 1, 1)
    def axis(a, b):
        v = a[None, :] + t * (b - a)[None, :] + rng.normal(0.0, jitter, (m, R))
        return np.asfortranarray(np.where(valid, v, np.nan))
    x, y = axis(src[:, 0], rcv[:, 0]), axis(src[:, 1], rcv[:, 1])
    z = axis(src[:, 2], rcv[:, 2])
    z = np.asfortranarray(np.where(valid, np.clip(z, 0.0, bz), np.nan))
    tab = ak135 if ak135 is not None else np.c_[_AK135_COARSE, np.zeros(len(_AK135_COARSE))]
    U = ak135_slowness(z, tab)
    sig = rng.uniform(0.04, 0.66, R)
    rng_box = (StepRangeLen(0.0, p.XYnodeSpacing, bx), StepRangeLen(0.0, p.XYnodeSpacing, by),
               StepRangeLen(0.0, p.ZnodeSpacing, bz))
    ds = make_datastruct(x, y, z, U, np.zeros(R), sig, p, box=rng_box)
    # "true" model -> tS (numpy nearest-nucleus forward model, host-side data synthesis only)
    tx, ty, tz = rng.uniform(0, bx, n_true), rng.uniform(0, by, n_true), rng.uniform(0, bz, n_true)
    tzeta = rng.uniform(0, zeta_scale, n_true)
    tS = np.zeros(R)
    for i0 in range(0, R if n_true > 0 else 0, 512):  # chunked nearest-nucleus forward model (host-side data synthesis only)
        sl = slice(i0, min(i0 + 512, R))
        xs, ys, zs = np.nan_to_num(x[:, sl]), np.nan_to_num(y[:, sl]), np.nan_to_num(z[:, sl])
        d = (tx[None, None] - xs[..., None]) ** 2 + (ty[None, None] - ys[..., None]) ** 2 + (tz[None, None] - zs[..., None]) ** 2
        zt = tzeta[np.argmin(d, 2)]
        seg = np.nan_to_num(ds.rayL[:, sl] * ds.rayU[:, sl]) * (0.5 * (zt[:-1] + zt[1:]) / 1000)
        tS[sl] = seg.sum(0)
    ds.tS = tS + rng.normal(0.0, sig)
    return ds
