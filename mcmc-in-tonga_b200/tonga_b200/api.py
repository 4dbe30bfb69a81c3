"""Host-side mirror of the reference's hot-path interface, forwarding to libtonga_b200.so.

Same names, argument order and return values as the Julia functions they replace (file:line relative to the reference):

  evaluate(model, dataStruct, TD_parameters) -> (model, dataStruct, valid)            MCsub.jl:123-185
  Interpolation(TD_parameters, model, X, Y, Z) -> Vector{Float64}(npoints)            MCsub.jl:306-336
  v_nearest(x, y, z, mx, my, mz, mv) -> Float64                                       MCsub.jl:247-263
  build_starting(TD_parameters, dataStruct) -> (model, dataStruct, valid)             MCsub.jl:76-121
  TD_inversion_function(TD_parameters, dataStruct, chain) -> model_hist              TD_inversion_function.jl:7-305
  run_chains(TD_parameters, dataStruct, chains) -> Vector{Vector{Model}}              main_inversion.jl:15 (the pmap line)

plus the batched / low-level objects (`Context`, `Chains`) the tests and bench.py drive.  Julia is absent from this
image; `julia/TongaB200.jl` holds the same shim written in Julia.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import PROPOSAL_DTYPE, TongaError, TongaParams, check, dp, ip, lp, bp
from .structs import DataStruct, Model, parameters


def make_params(TD_parameters: parameters, dataStruct: DataStruct, n_actions: int = 4) -> TongaParams:
    p = TongaParams()
    # min(xVec...) / max(xVec...) etc., TD_inversion_function.jl:30-32,78-80
    p.xmin, p.xmax = dataStruct.xVec.min(), dataStruct.xVec.max()
    p.ymin, p.ymax = dataStruct.yVec.min(), dataStruct.yVec.max()
    p.zmin, p.zmax = dataStruct.zVec.min(), dataStruct.zVec.max()
    p.sig, p.zeta_scale, p.max_sig = float(TD_parameters.sig), float(TD_parameters.zeta_scale), float(TD_parameters.max_sig)
    p.n_iter, p.burn_in, p.keep_each = float(TD_parameters.n_iter), float(TD_parameters.burn_in), float(TD_parameters.keep_each)
    p.min_cells, p.max_cells = int(TD_parameters.min_cells), int(TD_parameters.max_cells)
    p.prior, p.debug_prior, p.interp_style = int(TD_parameters.prior), int(TD_parameters.debug_prior), int(TD_parameters.interp_style)
    p.n_actions = n_actions
    return p


def pack_models(models, Kcap=None):
    """list of (x, y, z, zeta) or Model -> K[n] int32, cells[n,4,Kcap] float64."""
    tup = [(m.xCell, m.yCell, m.zCell, m.zeta) if isinstance(m, Model) else m for m in models]
    K = np.array([len(t[0]) for t in tup], np.int32)
    Kcap = int(Kcap or max(int(K.max()) if len(K) else 1, 1))
    cells = np.zeros((len(tup), 4, Kcap))
    for i, t in enumerate(tup):
        for a in range(4):
            cells[i, a, :K[i]] = t[a]
    return K, cells


class _Pinned:
    def __init__(self, nbytes):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        check(self.lib.tonga_host_alloc(C.byref(self.ptr), nbytes))
        self.nbytes = nbytes
        self._fin = weakref.finalize(self, self.lib.tonga_host_free, self.ptr)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (tonga_host_alloc): full-speed H2D / D2H for the host-buffer entry points.
    The memory is released when the array (and every view of it) is garbage collected."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    pin = _Pinned(max(count * dtype.itemsize, 1))
    buf = (C.c_char * pin.nbytes).from_address(pin.ptr.value)
    buf._pin = pin  # ctypes instances carry a __dict__: the numpy array keeps `buf` (its base) alive, `buf` keeps `pin` alive
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


class Context:
    """Device-resident ray geometry + observations: what `dataStruct` is to evaluate()."""

    def __init__(self, dataStruct: DataStruct | None, TD_parameters: parameters | None = None, device: int = 0,
                 n_actions: int = 4, params: TongaParams | None = None, device_ingest: bool = False):
        """device_ingest=True: rayL / rayU (load_data_Tonga.jl:66-69) and the flatten are computed on the GPU from
        dataStruct.rayX/Y/Z and dataStruct.U (tonga_create_from_points); same context bit for bit, much faster for large sets."""
        self.lib = _lib.load()
        self._h = C.c_void_p()
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        if dataStruct is None:  # geometry-less context (v_nearest / Interpolation on arbitrary points only)
            z = np.zeros((1, 0), order="F")
            p = params or TongaParams()
            p.interp_style = 1
            self.R, self.m = 0, 1
            check(self.lib.tonga_create(C.byref(self._h), 1, 0, dp(z), dp(z), dp(z), dp(z), dp(z), dp(z), dp(z), C.byref(p), device))
            self.params = p
        else:
            ds = dataStruct
            rx, ry, rz = f(ds.rayX), f(ds.rayY), f(ds.rayZ)
            self.m, self.R = rx.shape
            tS, sg = np.ascontiguousarray(ds.tS, dtype=np.float64), np.ascontiguousarray(ds.allSig, dtype=np.float64)
            self.params = params or make_params(TD_parameters, ds, n_actions)
            if device_ingest:
                U = f(ds.U)
                if U.shape != rx.shape:
                    raise TongaError(-1, "U must be m x R like rayX")
                check(self.lib.tonga_create_from_points(C.byref(self._h), self.m, self.R, dp(rx), dp(ry), dp(rz), dp(U), dp(tS), dp(sg),
                                                        C.byref(self.params), device))
            else:
                rl, ru = f(ds.rayL), f(ds.rayU)
                if rl.shape != (self.m - 1, self.R) or ru.shape != rl.shape:
                    raise TongaError(-1, "rayL / rayU must be (m-1) x R")
                check(self.lib.tonga_create(C.byref(self._h), self.m, self.R, dp(rx), dp(ry), dp(rz), dp(rl), dp(ru), dp(tS), dp(sg),
                                            C.byref(self.params), device))
        R, P, S, Pp = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.tonga_info(self._h, C.byref(R), C.byref(P), C.byref(S), C.byref(Pp)))
        self.P, self.S, self.Ppad = P.value, S.value, Pp.value
        self.device = device
        self._fin = weakref.finalize(self, self.lib.tonga_destroy, self._h)
        self._chains = weakref.WeakSet()  # live Chains built on this context: they must be destroyed first

    def close(self):
        for ch in list(self._chains):
            ch.close()
        self._fin()

    def ray_offsets(self) -> np.ndarray:
        off = np.zeros(self.R + 1, np.int32)
        check(self.lib.tonga_ray_offsets(self._h, ip(off)))
        return off

    def evaluate(self, x, y, z, zeta, noise: float = 1.0):
        """One model -> dict(ptS, phi, likelihood, loglik_gauss)."""
        x, y, z, zeta = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, zeta))
        ptS = np.zeros(max(self.R, 1))
        phi, like, lg = C.c_double(), C.c_double(), C.c_double()
        check(self.lib.tonga_evaluate(self._h, len(x), dp(x), dp(y), dp(z), dp(zeta), float(noise), dp(ptS), C.byref(phi), C.byref(like), C.byref(lg)))
        return dict(ptS=ptS[:self.R], phi=phi.value, likelihood=like.value, loglik_gauss=lg.value)

    def evaluate_batch(self, K, cells, noise=None, want_ptS=True, want_owners=False):
        """K[n], cells[n,4,Kcap] -> dict(phi[n], ptS[n,R] | None, owners[n,P] | None)."""
        K = np.ascontiguousarray(K, dtype=np.int32)
        cells = np.ascontiguousarray(cells, dtype=np.float64)
        n, four, Kcap = cells.shape
        assert four == 4 and len(K) == n
        noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        ptS = np.zeros((n, self.R)) if want_ptS else None
        phi = np.zeros(n)
        owners = np.zeros((n, self.P), np.int32) if want_owners else None
        check(self.lib.tonga_evaluate_batch(self._h, n, Kcap, ip(K), dp(cells), dp(noise), dp(ptS), dp(phi), ip(owners)))
        return dict(phi=phi, ptS=ptS, owners=owners)

    def misfit(self, ptS, noise=None):
        """phi of given t* vectors ptS[n, R] against this context's tS / allSig (MCsub.jl:169-173), by the device kernel."""
        ptS = np.ascontiguousarray(np.atleast_2d(ptS), dtype=np.float64)
        assert ptS.shape[1] == self.R
        noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        phi = np.zeros(len(ptS))
        check(self.lib.tonga_misfit(self._h, len(ptS), dp(ptS), dp(noise), dp(phi)))
        return phi

    def evaluate_batch_dev(self, n, Kcap, K_ptr, cells_ptr, noise_ptr, ptS_ptr, phi_ptr, owners_ptr=None):
        """Device-pointer variant (ints from torch .data_ptr()); asynchronous on the context's stream."""
        check(self.lib.tonga_evaluate_batch_dev(self._h, n, Kcap, K_ptr, cells_ptr, noise_ptr, ptS_ptr, phi_ptr, owners_ptr))

    def synchronize(self):
        check(self.lib.tonga_synchronize(self._h))

    def set_exact_only(self, flag: bool):
        """True: full evaluate entirely in exact FP64 (default: FP32 screening + exact re-scan of near ties)."""
        check(self.lib.tonga_set_exact_only(self._h, 1 if flag else 0))

    def interpolate(self, mx, my, mz, mv, X, Y, Z):
        mx, my, mz, mv = (np.ascontiguousarray(a, dtype=np.float64) for a in (mx, my, mz, mv))
        X, Y, Z = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (X, Y, Z))
        out = np.zeros(max(len(X), 1))
        idx = np.zeros(max(len(X), 1), np.int32)
        npts = C.c_int32()
        check(self.lib.tonga_interpolate(self._h, len(mx), dp(mx), dp(my), dp(mz), dp(mv), len(X), dp(X), len(Y), dp(Y), len(Z), dp(Z),
                                         dp(out), ip(idx), C.byref(npts)))
        return out[:npts.value].copy(), idx[:npts.value].copy()

    def peak_flops(self):
        a, b = C.c_double(), C.c_double()
        check(self.lib.tonga_peak_flops(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class Chains:
    """A batch of independent RJ-MCMC chains resident on one GPU (replaces the pmap chain farm, main_inversion.jl:15)."""

    SAMPLERS = {"auto": 0, "resident": 1, "wide": 2, "streamed": 3}

    def __init__(self, ctx: Context, n_chains: int, chain_id0: int = 0, seed: int = 20260000, hist_cap: int | None = None,
                 sampler: str = "auto", host_history: bool = False):
        """sampler: "resident" (chain state in shared memory, max_cells <= 126, up to ~200k ray points), "wide" (a full
        batched forward model per proposal; any size), "streamed" (per-point state in HBM, one streaming pass per proposal; any
        ray set, 6 B per point and chain) or "auto" (resident when it fits, else streamed).  Same chains either way.
        host_history=True keeps the kept models in mapped page-locked host memory: the kernels write them there while they run
        and history() returns views of those arrays without any copy."""
        self.ctx, self.lib, self.n = ctx, ctx.lib, n_chains
        self.chain_id0, self.seed = int(chain_id0), int(seed)
        p = ctx.params
        if hist_cap is None:  # num_models_per_chain, TD_inversion_function.jl:25
            hist_cap = int((p.n_iter - p.burn_in) / p.keep_each) + 1 if p.keep_each > 0 else 0
        self.hist_cap = hist_cap
        self._h = C.c_void_p()
        self.host_history = bool(host_history)
        if host_history:
            check(self.lib.tonga_chains_create_ex(ctx._h, C.byref(self._h), n_chains, chain_id0, seed, hist_cap, self.SAMPLERS[sampler] | 0x100))
        elif sampler == "auto":  # the reference-facing entry point
            check(self.lib.tonga_chains_create(ctx._h, C.byref(self._h), n_chains, chain_id0, seed, hist_cap))
        else:
            check(self.lib.tonga_chains_create_ex(ctx._h, C.byref(self._h), n_chains, chain_id0, seed, hist_cap, self.SAMPLERS[sampler]))
        self.KC = self.lib.tonga_chains_kcap(self._h)
        self.sampler = {1: "resident", 2: "wide", 3: "streamed"}[self.lib.tonga_chains_sampler(self._h)] if hasattr(self.lib, "tonga_chains_sampler") else "resident"
        self._fin = weakref.finalize(self, self.lib.tonga_chains_destroy, self._h)
        self._ctx_keepalive = ctx
        ctx._chains.add(self)

    def close(self):
        self._fin()

    def build_starting(self):
        check(self.lib.tonga_chains_build_starting(self._h))

    def set_models(self, K, cells, noise=None):
        K = np.ascontiguousarray(K, dtype=np.int32)
        cells = np.ascontiguousarray(cells, dtype=np.float64)
        assert cells.shape[0] == self.n and cells.shape[1] == 4
        noise = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        check(self.lib.tonga_chains_set_models(self._h, cells.shape[2], ip(K), dp(cells), dp(noise)))

    def set_exact_only(self, flag: bool):
        """True: every nearest-nucleus comparison in exact FP64 (default: FP32 screening + exact recheck of near ties)."""
        check(self.lib.tonga_chains_set_exact_only(self._h, 1 if flag else 0))

    def set_beta(self, beta):
        beta = None if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
        check(self.lib.tonga_chains_set_beta(self._h, dp(beta)))

    def get_beta(self) -> np.ndarray:
        beta = np.zeros(self.n)
        check(self.lib.tonga_chains_get_beta(self._h, dp(beta)))
        return beta

    def temper_swap(self, ladder_size: int, step: int, seed: int = 0, n_all: int = 0, phi_all: int = 0, noise_all: int = 0, beta_all: int = 0,
                    offset: int = 0):
        """One even / odd sweep of tempering swaps decided on the device (tonga_chains_temper_swap).  Without pointers the
        ladders are this batch's own consecutive replicas; with DEVICE pointers (ints) of all-gathered phi / noise / beta
        arrays the ladders may span ranks (`offset` = first global replica of this batch).  Asynchronous."""
        vp = lambda a: C.c_void_p(int(a)) if a else None
        check(self.lib.tonga_chains_temper_swap(self._h, int(n_all), vp(phi_all), vp(noise_all), vp(beta_all), int(offset), int(ladder_size), int(step),
                                                int(seed) & 0xFFFFFFFFFFFFFFFF))

    def temper_stats(self, reset: bool = False):
        """(accepted, attempted) swap pairs counted on the device since the last reset."""
        a, t = C.c_int64(), C.c_int64()
        check(self.lib.tonga_chains_temper_stats(self._h, C.byref(a), C.byref(t), 1 if reset else 0))
        return a.value, t.value

    def scalar_ptrs(self):
        """Device pointers of the per-chain phi / noise / beta arrays (zero-copy collectives)."""
        ptrs = [C.c_void_p() for _ in range(3)]
        check(self.lib.tonga_chains_scalar_ptrs(self._h, *[C.byref(q) for q in ptrs]))
        return {k: q.value for k, q in zip(("phi", "noise", "beta"), ptrs)}

    # ---- ray sharding of a streamed batch (tonga_chains_shard_*; BASELINE config 3 on several GPUs)
    def shard_init(self, rank: int, world: int):
        """This rank's part of a ray-sharded batch: allocates the exchange block, returns (device pointer, bytes)."""
        base, nbytes = C.c_void_p(), C.c_uint64()
        check(self.lib.tonga_chains_shard_init(self._h, int(rank), int(world), C.byref(base), C.byref(nbytes)))
        self.shard_rank, self.shard_world = int(rank), int(world)
        return base.value, nbytes.value

    def shard_connect(self, peer_bases):
        """peer_bases[world]: device pointers (ints), valid on this rank's device, of every rank's exchange block."""
        arr = (C.c_void_p * len(peer_bases))(*[C.c_void_p(int(b) if b else 0) for b in peer_bases])
        check(self.lib.tonga_chains_shard_connect(self._h, arr))

    def shard_info(self):
        """dict(rank, world, ray0, ray1, point0, point1): the rank's own rays / points in the library's (length-sorted) order."""
        r, w, r0, r1 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        p0, p1 = C.c_int64(), C.c_int64()
        check(self.lib.tonga_chains_shard_info(self._h, C.byref(r), C.byref(w), C.byref(r0), C.byref(r1), C.byref(p0), C.byref(p1)))
        return dict(rank=r.value, world=w.value, ray0=r0.value, ray1=r1.value, point0=p0.value, point1=p1.value)

    def run(self, n_iter: int, recs: np.ndarray | None = None, record: bool = False, trace: bool = False):
        """Generate mode (device Philox) unless `recs` ([n, n_iter] PROPOSAL_DTYPE) is given (replay).
        -> dict(recs (if record), accept, phi, K (if trace))."""
        mode = 0 if recs is None else 1
        if recs is not None:
            recs = np.ascontiguousarray(recs, dtype=PROPOSAL_DTYPE)
            assert recs.shape == (self.n, n_iter), recs.shape
        elif record:
            recs = np.zeros((self.n, n_iter), PROPOSAL_DTYPE)
        acc = np.zeros((self.n, n_iter), np.int8) if trace else None
        phi = np.zeros((self.n, n_iter)) if trace else None
        K = np.zeros((self.n, n_iter), np.int32) if trace else None
        check(self.lib.tonga_chains_run(self._h, n_iter, mode, None if recs is None else recs.ctypes.data, bp(acc), dp(phi), ip(K)))
        return dict(recs=recs, accept=acc, phi=phi, K=K)

    def state(self, want_ptS=True, want_owners=False, out: dict | None = None):
        n, KC = self.n, self.KC
        if out is not None:  # caller-owned (e.g. pinned) buffers, reused across calls
            K, cells, phi, noise, ptS, owners = out["K"], out["cells"], out["phi"], out["noise"], out.get("ptS"), out.get("owners")
        else:
            K = np.zeros(n, np.int32)
            cells = np.zeros((n, 4, KC))
            phi = np.zeros(n)
            noise = np.zeros(n)
            ptS = np.zeros((n, self.ctx.R)) if want_ptS else None
            owners = np.zeros((n, self.ctx.P), np.int32) if want_owners else None
        check(self.lib.tonga_chains_get_state(self._h, KC, ip(K), dp(cells), dp(phi), dp(ptS), dp(noise), ip(owners)))
        return dict(K=K, cells=cells, phi=phi, ptS=ptS, noise=noise, owners=owners)

    def stats(self):
        it = C.c_int64()
        counts = np.zeros((self.n, 3, 5), np.int64)  # proposed / accepted / evaluated per action
        check(self.lib.tonga_chains_get_stats(self._h, C.byref(it), lp(counts)))
        return it.value, counts

    def reset(self):
        check(self.lib.tonga_chains_reset(self._h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.tonga_chains_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def alloc_buffers(self, pinned=True, want_ptS=True):
        """Reusable output buffers for history() / state() (page-locked by default)."""
        n, H, KC, R = self.n, self.hist_cap, self.KC, self.ctx.R
        mk = pinned_empty if pinned else (lambda shape, dt: np.zeros(shape, dt))
        hist = dict(n_hist=mk((n,), np.int32), K=mk((n, H), np.int32), cells=mk((n, H, 4, KC), np.float64), phi=mk((n, H), np.float64),
                    ptS=mk((n, H, R), np.float64) if want_ptS else None, iter=mk((n, H), np.int64), action=mk((n, H), np.int32),
                    accept=mk((n, H), np.int32), next_action=mk((n, H), np.int32))
        state = dict(K=mk((n,), np.int32), cells=mk((n, 4, KC), np.float64), phi=mk((n,), np.float64), noise=mk((n,), np.float64),
                     ptS=mk((n, R), np.float64) if want_ptS else None, owners=None)
        return hist, state

    def history(self, want_ptS=True, out: dict | None = None):
        n, H, KC, R = self.n, self.hist_cap, self.KC, self.ctx.R
        if self.host_history and out is None:  # zero-copy views of the library's mapped host arrays (valid until close())
            ptrs = [C.c_void_p() for _ in range(8)]
            check(self.lib.tonga_chains_history_host(self._h, *[C.byref(p) for p in ptrs]))
            n_hist = np.zeros(n, np.int32)
            check(self.lib.tonga_chains_get_history(self._h, KC, ip(n_hist), None, None, None, None, None, None, None, None))
            def view(ptr, shape, dt):
                cnt = int(np.prod(shape))
                if cnt == 0:
                    return np.zeros(shape, dt)
                buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr.value)
                return np.frombuffer(buf, dtype=dt, count=cnt).reshape(shape)
            return dict(n_hist=n_hist, K=view(ptrs[0], (n, H), np.int32), cells=view(ptrs[1], (n, H, 4, KC), np.float64),
                        phi=view(ptrs[2], (n, H), np.float64), ptS=view(ptrs[3], (n, H, R), np.float64), iter=view(ptrs[4], (n, H), np.int64),
                        action=view(ptrs[5], (n, H), np.int32), accept=view(ptrs[6], (n, H), np.int32), next_action=view(ptrs[7], (n, H), np.int32))
        if out is None:
            out = dict(n_hist=np.zeros(n, np.int32), K=np.zeros((n, H), np.int32), cells=np.zeros((n, H, 4, KC)),
                       phi=np.zeros((n, H)), ptS=np.zeros((n, H, R)) if want_ptS else None, iter=np.zeros((n, H), np.int64),
                       action=np.zeros((n, H), np.int32), accept=np.zeros((n, H), np.int32), next_action=np.zeros((n, H), np.int32))
        check(self.lib.tonga_chains_get_history(self._h, KC, ip(out["n_hist"]), ip(out["K"]), dp(out["cells"]), dp(out["phi"]),
                                                dp(out["ptS"]), lp(out["iter"]), ip(out["action"]), ip(out["accept"]),
                                                ip(out["next_action"])))
        return out

    def checkpoint(self) -> dict:
        """Everything a resumed batch needs to continue bit-identically (TD_inversion_function.jl:282-294 saves model, iter,
        saved_#, model_num, model_hist): the current models, the iteration count (= Philox counter), the thinning phase, the
        counters and the kept models."""
        st = self.state(want_ptS=False)
        it = C.c_int64()
        model_num = np.zeros(self.n, np.int64)
        pending = np.zeros(self.n, np.int32)
        check(self.lib.tonga_chains_get_progress(self._h, C.byref(it), lp(model_num), ip(pending)))
        _, counts = self.stats()
        hist = self.history()
        return dict(K=st["K"], cells=st["cells"], noise=st["noise"], beta=self.get_beta(), iter=np.int64(it.value), model_num=model_num, pending_slot=pending,
                    counts=counts, seed=np.uint64(self.seed), chain_id0=np.int64(self.chain_id0), hist_cap=np.int64(self.hist_cap),
                    **{"hist_" + k: v for k, v in hist.items()})

    def restore(self, ck: dict):
        """Inverse of checkpoint() on a batch created with the same context, chain count, chain_id0, seed and hist_cap."""
        if int(ck["hist_cap"]) != self.hist_cap or len(ck["K"]) != self.n or int(ck["chain_id0"]) != self.chain_id0 or int(ck["seed"]) != self.seed:
            raise TongaError(-1, "checkpoint does not belong to this batch (chain count, chain_id0, seed or hist_cap differ)")
        self.set_models(ck["K"], ck["cells"][:, :, :self.KC], ck["noise"])
        if "beta" in ck:  # tempered batches (extension): the replicas' inverse temperatures
            self.set_beta(ck["beta"])
        c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
        check(self.lib.tonga_chains_set_progress(self._h, int(ck["iter"]), lp(c(ck["model_num"], np.int64)), ip(c(ck["pending_slot"], np.int32)),
                                                 lp(c(ck["counts"], np.int64))))
        check(self.lib.tonga_chains_set_history(self._h, self.KC, ip(c(ck["hist_n_hist"], np.int32)), ip(c(ck["hist_K"], np.int32)),
                                                dp(c(ck["hist_cells"], np.float64)), dp(c(ck["hist_phi"], np.float64)), dp(c(ck["hist_ptS"], np.float64)),
                                                lp(c(ck["hist_iter"], np.int64)), ip(c(ck["hist_action"], np.int32)), ip(c(ck["hist_accept"], np.int32)),
                                                ip(c(ck["hist_next_action"], np.int32))))

    def raster(self, X, Y, Z):
        """plot_model_hist's accumulation (MCsub.jl:753-825) at arbitrary nodes -> (sum, sumsq, count) over all kept models."""
        X, Y, Z = (np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (X, Y, Z))
        assert len(X) == len(Y) == len(Z)
        s1, s2 = np.zeros(len(X)), np.zeros(len(X))
        cnt = C.c_int64()
        check(self.lib.tonga_chains_raster(self._h, len(X), dp(X), dp(Y), dp(Z), dp(s1), dp(s2), C.byref(cnt)))
        return s1, s2, cnt.value

    def profile(self, enable=True, read=False):
        """Per-phase cycle totals of the sampler kernel: [n, 8] = A, B, C, D+E, F, G (developer aid)."""
        cyc = np.zeros((self.n, 16), np.int64) if read else None
        check(self.lib.tonga_chains_profile(self._h, 1 if enable else 0, lp(cyc)))
        return cyc

    def verify(self):
        """Full re-evaluate of every chain on the device vs the incrementally maintained state."""
        mm, dphi, dts = C.c_int64(), C.c_double(), C.c_double()
        check(self.lib.tonga_chains_verify(self._h, C.byref(mm), C.byref(dphi), C.byref(dts)))
        return mm.value, dphi.value, dts.value

    def device_ptrs(self):
        names = ("n_hist", "hist_K", "hist_cells", "hist_phi", "hist_ptS", "state_K", "state_cells", "state_phi")
        ptrs = [C.c_void_p() for _ in names]
        check(self.lib.tonga_chains_device_ptrs(self._h, *[C.byref(p) for p in ptrs]))
        return {k: p.value for k, p in zip(names, ptrs)}


# ------------------------------------------------------------------------------------------ reference-named functions
_CTX_CACHE: dict = {}
_UTIL_CTX: list = []


_CTX_CACHE_MAX = 8


def context_for(dataStruct: DataStruct, TD_parameters: parameters, device: int = 0, n_actions: int = 4) -> Context:
    """One Context per (dataStruct, resolved parameter VALUES) pair, like the Julia shim's cache keyed on objectid(dataStruct).

    `parameters` is mutable: the key holds the bytes of the device-side tonga_params the Context would be built with, so a
    mutated (or recycled-id) parameters object gets a fresh Context instead of silently reusing stale values.  Entries whose
    dataStruct died are closed and dropped; at most _CTX_CACHE_MAX contexts stay alive (least recently used goes first)."""
    for k in [k for k, (ref, _) in _CTX_CACHE.items() if ref() is None]:
        _CTX_CACHE.pop(k)[1].close()
    key = (id(dataStruct), bytes(make_params(TD_parameters, dataStruct, n_actions)), device)
    ent = _CTX_CACHE.pop(key, None)
    if ent is None or ent[0]() is not dataStruct:
        if ent is not None:
            ent[1].close()
        ent = (weakref.ref(dataStruct), Context(dataStruct, TD_parameters, device, n_actions))
    _CTX_CACHE[key] = ent  # re-inserted at the end: dict order = recency
    while len(_CTX_CACHE) > _CTX_CACHE_MAX:
        _CTX_CACHE.pop(next(iter(_CTX_CACHE)))[1].close()
    return ent[1]


def _util_ctx() -> Context:
    if not _UTIL_CTX:
        _UTIL_CTX.append(Context(None))
    return _UTIL_CTX[0]


def v_nearest(x, y, z, mx, my, mz, mv) -> float:
    """MCsub.jl:247-263."""
    out, _ = _util_ctx().interpolate(mx, my, mz, mv, [x], [y], [z])
    return float(out[0])


def Interpolation(TD_parameters: parameters, model: Model, X, Y, Z) -> np.ndarray:  # noqa: N802
    """MCsub.jl:306-336 (interp_style 1; style 2 raises, as it does in the reference: SURVEY F7)."""
    if TD_parameters.interp_style != 1:
        raise TongaError(-1, "interp_style 2 (IDW) is broken in the reference (MCsub.jl:332 uses undefined variables)")
    out, _ = _util_ctx().interpolate(model.xCell, model.yCell, model.zCell, model.zeta, X, Y, Z)
    return out


def evaluate(model: Model, dataStruct: DataStruct, TD_parameters: parameters):
    """MCsub.jl:123-185: sets model.phi / ptS / tS / likelihood, returns (model, dataStruct, valid=1)."""
    ctx = context_for(dataStruct, TD_parameters)
    r = ctx.evaluate(model.xCell, model.yCell, model.zCell, model.zeta)
    model.phi = r["phi"]
    model.likelihood = r["likelihood"]
    if TD_parameters.debug_prior != 1:  # :134-136 returns before ptS / tS are touched
        model.ptS = r["ptS"].copy()
        model.tS = dataStruct.tS  # alias, :175
    return model, dataStruct, 1


def build_starting(TD_parameters: parameters, dataStruct: DataStruct, seed: int = 20260000, chain: int = 0):
    """MCsub.jl:76-121 with the device Philox stream of chain `chain`."""
    ctx = context_for(dataStruct, TD_parameters)
    ch = Chains(ctx, 1, chain_id0=chain, seed=seed, hist_cap=0)
    ch.build_starting()
    st = ch.state()
    k = int(st["K"][0])
    model = Model(float(k), st["cells"][0, 0, :k].copy(), st["cells"][0, 1, :k].copy(), st["cells"][0, 2, :k].copy(),
                  st["cells"][0, 3, :k].copy(), float(st["phi"][0]), st["ptS"][0].copy(), dataStruct.tS,
                  ctx.evaluate(st["cells"][0, 0, :k], st["cells"][0, 1, :k], st["cells"][0, 2, :k], st["cells"][0, 3, :k])["likelihood"],
                  -1, -1, -1.0, -1.0)
    ch.close()
    return model, dataStruct, 1


def history_to_models(hist: dict, dataStruct: DataStruct, likelihood: float, reference_aliasing: bool = False):
    """Packed history -> Vector{Vector{Model}} as `plot_model_hist` consumes it (MCsub.jl:762-767).
    reference_aliasing=True APPROXIMATES the aliasing artefact of the reference's stored models: `push!(model_hist, model)`
    (TD_inversion_function.jl:280) stores a reference, and :73-74 keep rewriting model.action / model.accept of that object on
    every later iteration until an accepted proposal rebinds `model`.  The reference therefore stores the action of the first
    later iteration whose proposal was accepted (and accept = 0); this flag stores the action of the NEXT iteration only
    (the one value the device keeps, hist_next) and accept = 0 -- exact when that proposal is accepted, an approximation
    otherwise.  Nuclei, zeta, phi and ptS are unaffected by the artefact."""
    out = []
    for c in range(len(hist["n_hist"])):
        ms = []
        for j in range(min(int(hist["n_hist"][c]), hist["K"].shape[1])):
            k = int(hist["K"][c, j])
            cl = hist["cells"][c, j]
            action = int(hist["next_action"][c, j]) if reference_aliasing and hist["next_action"][c, j] > 0 else int(hist["action"][c, j])
            accept = 0 if reference_aliasing else int(hist["accept"][c, j])
            ms.append(Model(float(k), cl[0, :k].copy(), cl[1, :k].copy(), cl[2, :k].copy(), cl[3, :k].copy(), float(hist["phi"][c, j]),
                            hist["ptS"][c, j].copy() if hist["ptS"] is not None else np.zeros(1), dataStruct.tS, likelihood,
                            action, accept, -1.0, -1.0))
        out.append(ms)
    return out


def run_chains(TD_parameters: parameters, dataStruct: DataStruct, chains, seed: int = 20260000, device: int = 0,
               reference_aliasing: bool = False):
    """Replaces `pmap(x -> TD_inversion_function(TD_parameters, dataStruct, x), 1:n_chains)` (main_inversion.jl:15):
    all chains of `chains` (an iterable of consecutive chain numbers) run batched on one GPU."""
    chains = list(chains)
    assert chains == list(range(chains[0], chains[0] + len(chains))), "chains must be consecutive"
    ctx = context_for(dataStruct, TD_parameters, device)
    ch = Chains(ctx, len(chains), chain_id0=chains[0], seed=seed)
    ch.build_starting()  # TD_inversion_function.jl:43-45
    ch.run(int(TD_parameters.n_iter))  # :70
    hist = ch.history()
    like = ctx.evaluate(*[hist["cells"][0, 0, a, :max(int(hist["K"][0, 0]), 1)] for a in range(4)])["likelihood"] if TD_parameters.debug_prior != 1 else 1.0
    out = history_to_models(hist, dataStruct, like, reference_aliasing)
    ch.close()
    return out


def TD_inversion_function(TD_parameters: parameters, dataStruct1: DataStruct, chain: int, seed: int = 20260000):  # noqa: N802
    """TD_inversion_function.jl:7-305 for one chain -> model_hist (list of Model).  Prefer run_chains() for many."""
    return run_chains(TD_parameters, dataStruct1, [chain], seed=seed)[0]


def slice_nodes(dataStruct: DataStruct, TD_parameters: parameters):
    """The node sets plot_model_hist rasterises (MCsub.jl:756,766-768,791,800-802): for every l0 in ySlice the grid
    xVec x zVec at y = l0 ("xz"), for every l0 in zSlice the grid xVec x yVec at z = l0 ("xy").
    -> list of (kind, l0, shape, X, Y, Z) with X/Y/Z flattened in Julia's comprehension order (first index fastest)."""
    xv, yv, zv = dataStruct.xVec.vec(), dataStruct.yVec.vec(), dataStruct.zVec.vec()
    out = []
    if TD_parameters.xzMap:
        for l0 in TD_parameters.ySlice:
            Xg, Zg = np.meshgrid(xv, zv, indexing="ij")
            out.append(("xz", float(l0), Xg.shape, Xg.ravel(order="F"), np.full(Xg.size, float(l0)), Zg.ravel(order="F")))
    if TD_parameters.xyMap:
        for l0 in TD_parameters.zSlice:
            Xg, Yg = np.meshgrid(xv, yv, indexing="ij")
            out.append(("xy", float(l0), Xg.shape, Xg.ravel(order="F"), Yg.ravel(order="F"), np.full(Xg.size, float(l0))))
    return out


def finish_maps(s1, s2, count, shape):
    """mean / std / masked mean exactly as MCsub.jl:774-782 (Julia `std` is the n-1 sample standard deviation)."""
    mean = s1 / count
    var = np.maximum(s2 - s1 * s1 / count, 0.0) / max(count - 1, 1)
    std = np.sqrt(var)
    mask = np.where(std > 5, np.nan, 1.0)  # :777-781
    f = lambda a: a.reshape(shape, order="F")
    return dict(mean=f(mean), std=f(std), masked=f(mask * mean), count=count)


def plot_model_hist(chains: "Chains", dataStruct: DataStruct, TD_parameters: parameters, reduce=None):
    """The numerical content of plot_model_hist(model_hist, dataStruct, TD_parameters, cmax) (MCsub.jl:753-825) from the
    device-resident history of `chains`: {("xz", 700.0): {mean, std, masked}, ...}.  `reduce(s1, s2, count)` may add the
    sums across ranks (tonga_b200.dist.allreduce_sums) before the statistics are formed; plotting itself is out of scope."""
    out = {}
    for kind, l0, shape, X, Y, Z in slice_nodes(dataStruct, TD_parameters):
        s1, s2, cnt = chains.raster(X, Y, Z)
        if reduce is not None:
            s1, s2, cnt = reduce(s1, s2, cnt)
        out[(kind, l0)] = finish_maps(s1, s2, cnt, shape)
    return out
