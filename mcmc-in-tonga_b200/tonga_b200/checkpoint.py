"""Checkpoint / resume and `model.jld`-equivalent output.

The reference saves, per chain and every 10 % of the run, `model, dataStruct, iter, saved_#, model_num, model_hist, burnin` into a
JLD file (TD_inversion_function.jl:282-294) and resumes from the newest one (:40-67); at the end main_inversion.jl:18 writes
all chains' `model_hist` to `model.jld`.  No JLD/HDF5 writer exists in this image, so

  save_checkpoint / load_checkpoint / resume   keep the same information for a whole batch of chains in one `.npz`
                                               (continuing a resumed batch is bit-identical to an uninterrupted run), and
  export_model_hist                            writes the kept models in a flat little-endian binary that
                                               `julia/model_hist_to_jld.jl` turns into the reference's `model.jld`
                                               (Vector{Vector{Model}}) with Base + JLD only.
"""
from __future__ import annotations

import struct

import numpy as np

from .api import Chains, Context

MAGIC = b"TONGAMH1"


def save_checkpoint(path: str, chains: Chains) -> dict:
    """One file for the batch; returns the dict that was written (np.savez_compressed keys)."""
    ck = chains.checkpoint()
    ck["sampler"] = np.bytes_(chains.sampler)
    ck["burnin"] = np.bool_(int(ck["iter"]) >= chains.ctx.params.burn_in)  # the reference's `burnin` flag (:285 vs :292)
    ck["saved_num"] = np.minimum(ck["hist_n_hist"], chains.hist_cap)       # its `saved_#`
    np.savez_compressed(path, **ck)
    return ck


def load_checkpoint(path: str) -> dict:
    with np.load(path, allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def resume(ctx: Context, ck: dict, sampler: str | None = None) -> Chains:
    """A new batch on `ctx` that continues where the checkpoint stopped (same chains, same Philox streams)."""
    kind = sampler or (np.asarray(ck["sampler"]).item().decode() if "sampler" in ck else "auto")
    ch = Chains(ctx, len(ck["K"]), chain_id0=int(ck["chain_id0"]), seed=int(ck["seed"]), hist_cap=int(ck["hist_cap"]), sampler=kind)
    ch.restore(ck)
    return ch


def export_model_hist(path: str, hist: dict, likelihood: float, reference_aliasing: bool = False) -> int:
    """Kept models of all chains -> flat binary (little-endian):
         magic[8] | n_chains i64 | R i64 | per chain: n_models i64 | per model: K i64, action i64, accept i64, phi f64,
         likelihood f64, x[K] y[K] z[K] zeta[K] f64, ptS[R] f64.
    `hist` is Chains.history() (or the gathered ensemble with the same keys).  Returns the number of models written."""
    n, H = hist["K"].shape
    R = hist["ptS"].shape[2]
    total = 0
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<qq", n, R))
        for c in range(n):
            nm = min(int(hist["n_hist"][c]), H)
            f.write(struct.pack("<q", nm))
            for j in range(nm):
                k = int(hist["K"][c, j])
                nxt = int(hist["next_action"][c, j])
                action = nxt if (reference_aliasing and nxt > 0) else int(hist["action"][c, j])
                accept = 0 if reference_aliasing else int(hist["accept"][c, j])
                f.write(struct.pack("<qqqdd", k, action, accept, float(hist["phi"][c, j]), float(likelihood)))
                f.write(np.ascontiguousarray(hist["cells"][c, j, :, :k], dtype="<f8").tobytes())
                f.write(np.ascontiguousarray(hist["ptS"][c, j], dtype="<f8").tobytes())
                total += 1
    return total


def import_model_hist(path: str):
    """Inverse of export_model_hist -> list (chains) of lists of dict(K, action, accept, phi, likelihood, cells[4,K], ptS[R])."""
    out = []
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not a tonga model_hist file")
        n, R = struct.unpack("<qq", f.read(16))
        for _ in range(n):
            (nm,) = struct.unpack("<q", f.read(8))
            ms = []
            for _ in range(nm):
                k, action, accept, phi, like = struct.unpack("<qqqdd", f.read(40))
                cells = np.frombuffer(f.read(32 * k), dtype="<f8").reshape(4, k)
                ptS = np.frombuffer(f.read(8 * R), dtype="<f8")
                ms.append(dict(K=k, action=action, accept=accept, phi=phi, likelihood=like, cells=cells, ptS=ptS))
            out.append(ms)
    return out
