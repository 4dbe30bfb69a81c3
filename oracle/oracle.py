"""ctypes loader for the C oracle (oracle/libtonga_oracle.so).

CPU ORACLE -- test infrastructure, NOT the product.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module.  PARITY UNPINNED (see tonga_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "sig", "zeta_scale",
                                          "max_sig", "n_iter", "burn_in", "keep_each")] + \
               [(n, C.c_int32) for n in ("min_cells", "max_cells", "prior", "debug_prior", "interp_style", "reserved")]


class OrcData(C.Structure):
    _fields_ = [("m", C.c_int32), ("R", C.c_int32)] + [(n, c_dp) for n in ("rayX", "rayY", "rayZ", "rayL", "rayU", "tS", "allSig")]


class OrcProposal(C.Structure):
    _fields_ = [("action", C.c_int32), ("idx", C.c_int32), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("zeta", C.c_double), ("u", C.c_double)]


PROPOSAL_DTYPE = np.dtype([("action", "<i4"), ("idx", "<i4"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8"),
                           ("zeta", "<f8"), ("u", "<f8")])
assert PROPOSAL_DTYPE.itemsize == C.sizeof(OrcProposal) == 48


class OrcModel(C.Structure):
    _fields_ = [("K", C.c_int32), ("cap", C.c_int32), ("x", c_dp), ("y", c_dp), ("z", c_dp), ("zeta", c_dp),
                ("phi", C.c_double), ("likelihood", C.c_double), ("ptS", c_dp), ("action", C.c_int32),
                ("accept", C.c_int32), ("noise", C.c_double)]


class OrcRng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4), ("have_spare", C.c_int), ("spare", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libtonga_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("tonga_oracle.c", "tonga_oracle_mt.c", "tonga_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-B", "libtonga_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_v_nearest.restype = C.c_double
        L.orc_v_nearest.argtypes = [C.c_double] * 3 + [C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip]
        L.orc_interpolation.restype = C.c_int
        L.orc_interpolation.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_ip]
        L.orc_evaluate.restype = C.c_int
        L.orc_evaluate.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcData), C.POINTER(OrcModel), c_ip, c_dp]
        L.orc_misfit.restype = None
        L.orc_misfit.argtypes = [C.c_int, c_dp, c_dp, c_dp, C.c_double, c_dp, c_dp, c_dp]
        L.orc_build_starting.restype = C.c_int
        L.orc_build_starting.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcData), C.POINTER(OrcRng), C.POINTER(OrcModel)]
        L.orc_chain_run.restype = C.c_int
        L.orc_chain_run.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcData), C.POINTER(OrcModel), C.c_int64, C.c_int64,
                                    C.c_int, C.c_void_p, C.POINTER(OrcRng), C.c_void_p, c_dp, c_ip,
                                    C.c_int32, c_ip, C.POINTER(C.c_int64), c_ip, c_dp, c_dp, c_dp,
                                    C.POINTER(C.c_int64), c_ip, c_ip]
        L.orc_rng_seed.argtypes = [C.POINTER(OrcRng), C.c_uint64]
        L.orc_chain_farm.restype = C.c_int
        L.orc_chain_farm.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcData), C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                     c_dp, c_ip, C.POINTER(C.c_int64)]
        L.orc_chain_farm_from.restype = C.c_int
        L.orc_chain_farm_from.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcData), C.c_int, C.c_int64, C.c_int, C.c_uint64,
                                          c_ip, c_dp, C.c_int, c_dp, c_ip, C.POINTER(C.c_int64)]
        L.orc_ray_lengths.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def make_params(box, sig=10, zeta_scale=50, max_sig=0.1, n_iter=1e3, burn_in=5e2, keep_each=1e1, min_cells=5,
                max_cells=100, prior=1, debug_prior=0, interp_style=1, n_actions=4) -> OrcParams:
    p = OrcParams()
    p.xmin, p.xmax, p.ymin, p.ymax, p.zmin, p.zmax = [float(v) for v in box]
    p.sig, p.zeta_scale, p.max_sig = float(sig), float(zeta_scale), float(max_sig)
    p.n_iter, p.burn_in, p.keep_each = float(n_iter), float(burn_in), float(keep_each)
    p.min_cells, p.max_cells, p.prior, p.debug_prior, p.interp_style = min_cells, max_cells, prior, debug_prior, interp_style
    p.reserved = n_actions
    return p


class Data:
    """Holds Fortran-ordered copies of the reference-layout arrays and the OrcData view of them."""

    def __init__(self, rayX, rayY, rayZ, rayL, rayU, tS, allSig):
        f = lambda a: np.asfortranarray(a, dtype=np.float64)
        self.rayX, self.rayY, self.rayZ, self.rayL, self.rayU = f(rayX), f(rayY), f(rayZ), f(rayL), f(rayU)
        self.tS, self.allSig = np.ascontiguousarray(tS, dtype=np.float64), np.ascontiguousarray(allSig, dtype=np.float64)
        self.m, self.R = self.rayX.shape
        assert self.rayL.shape == (self.m - 1, self.R)
        d = OrcData()
        d.m, d.R = self.m, self.R
        d.rayX, d.rayY, d.rayZ, d.rayL, d.rayU = map(_dp, (self.rayX, self.rayY, self.rayZ, self.rayL, self.rayU))
        d.tS, d.allSig = _dp(self.tS), _dp(self.allSig)
        self.c = d


class ModelBuf:
    def __init__(self, cap: int, R: int):
        self.cap, self.R = cap, R
        self.x, self.y, self.z, self.zeta = (np.zeros(cap) for _ in range(4))
        self.ptS = np.zeros(max(R, 1))
        m = OrcModel()
        m.K, m.cap = 0, cap
        m.x, m.y, m.z, m.zeta, m.ptS = map(_dp, (self.x, self.y, self.z, self.zeta, self.ptS))
        m.phi, m.likelihood, m.action, m.accept, m.noise = -1.0, -1.0, -1, -1, 1.0
        self.c = m

    def set(self, x, y, z, zeta, noise=1.0):
        K = len(x)
        assert K <= self.cap
        self.x[:K], self.y[:K], self.z[:K], self.zeta[:K] = x, y, z, zeta
        self.c.K = K
        self.c.noise = noise
        return self

    @property
    def K(self):
        return self.c.K

    def cells(self):
        K = self.c.K
        return self.x[:K].copy(), self.y[:K].copy(), self.z[:K].copy(), self.zeta[:K].copy()


def v_nearest(x, y, z, mx, my, mz, mv):
    mx, my, mz, mv = (np.ascontiguousarray(a, dtype=np.float64) for a in (mx, my, mz, mv))
    idx = C.c_int32(-1)
    v = lib().orc_v_nearest(x, y, z, len(mx), _dp(mx), _dp(my), _dp(mz), _dp(mv), C.byref(idx))
    return v, idx.value


def interpolation(mx, my, mz, mv, X, Y, Z):
    mx, my, mz, mv, X, Y, Z = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (mx, my, mz, mv, X, Y, Z))
    out = np.zeros(max(len(X), 1))
    idx = np.zeros(max(len(X), 1), np.int32)
    n = lib().orc_interpolation(len(mx), _dp(mx), _dp(my), _dp(mz), _dp(mv), len(X), _dp(X), len(Y), _dp(Y), len(Z), _dp(Z),
                                _dp(out), _ip(idx))
    return out[:n].copy(), idx[:n].copy()


def misfit(ptS, tS, allSig, noise=1.0):
    """MCsub.jl:169-182 -> (phi, likelihood, loglik_gauss); the C function orc_evaluate itself calls."""
    ptS, tS, allSig = (np.ascontiguousarray(a, np.float64) for a in (ptS, tS, allSig))
    out = np.zeros(3)
    lib().orc_misfit(len(tS), _dp(ptS), _dp(tS), _dp(allSig), float(noise), _dp(out[0:1]), _dp(out[1:2]), _dp(out[2:3]))
    return float(out[0]), float(out[1]), float(out[2])


def evaluate(p: OrcParams, d: Data, x, y, z, zeta, noise=1.0, want_owners=False):
    """-> dict(ptS, phi, likelihood, loglik_gauss, owners[m,R] (Fortran) or None, valid)"""
    mb = ModelBuf(max(len(x), 1), d.R).set(x, y, z, zeta, noise)
    owners = np.zeros((d.m, d.R), np.int32, order="F") if want_owners else None
    lg = C.c_double(0.0)
    v = lib().orc_evaluate(C.byref(p), C.byref(d.c), C.byref(mb.c), _ip(owners) if want_owners else None, C.byref(lg))
    return dict(ptS=mb.ptS[:d.R].copy(), phi=mb.c.phi, likelihood=mb.c.likelihood, loglik_gauss=lg.value, owners=owners, valid=v)


def rng(seed: int) -> OrcRng:
    g = OrcRng()
    lib().orc_rng_seed(C.byref(g), seed)
    return g


def build_starting(p: OrcParams, d: Data, g: OrcRng, cap=None) -> ModelBuf:
    mb = ModelBuf(cap or (p.max_cells + 1), d.R)
    rc = lib().orc_build_starting(C.byref(p), C.byref(d.c), C.byref(g), C.byref(mb.c))
    if rc < 0:
        raise RuntimeError(f"orc_build_starting rc={rc}")
    return mb


class ChainRun:
    """Result of chain_run: traces + thinned history."""


def chain_run(p: OrcParams, d: Data, mb: ModelBuf, n_iter: int, iter0: int = 1, recs: np.ndarray | None = None,
              g: OrcRng | None = None, hist_cap: int = 0, n_hist: int = 0, model_num: int = 0) -> ChainRun:
    """mode replay if `recs` given (and g is None), else generate with `g` (records returned in .recs)."""
    mode = 0 if g is None else 1
    if mode == 0:
        recs = np.ascontiguousarray(recs, dtype=PROPOSAL_DTYPE)
        assert len(recs) >= n_iter
    else:
        recs = np.zeros(n_iter, PROPOSAL_DTYPE)
    out = ChainRun()
    out.accept = np.zeros(n_iter, np.int8)
    out.phi = np.zeros(n_iter)
    out.K = np.zeros(n_iter, np.int32)
    hc = max(hist_cap, 1)
    out.hist_K = np.zeros(hc, np.int32)
    out.hist_cells = np.zeros((hc, 4, mb.cap))
    out.hist_phi = np.zeros(hc)
    out.hist_ptS = np.zeros((hc, d.R))
    out.hist_iter = np.zeros(hc, np.int64)
    out.hist_action = np.zeros(hc, np.int32)
    out.hist_accept = np.zeros(hc, np.int32)
    nh = C.c_int32(n_hist)
    mn = C.c_int64(model_num)
    rc = lib().orc_chain_run(C.byref(p), C.byref(d.c), C.byref(mb.c), iter0, n_iter, mode, recs.ctypes.data,
                             C.byref(g) if g is not None else None, out.accept.ctypes.data, _dp(out.phi), _ip(out.K),
                             hist_cap, C.byref(nh), C.byref(mn), _ip(out.hist_K), _dp(out.hist_cells), _dp(out.hist_phi),
                             _dp(out.hist_ptS), out.hist_iter.ctypes.data_as(C.POINTER(C.c_int64)), _ip(out.hist_action),
                             _ip(out.hist_accept))
    if rc < 0:
        raise RuntimeError(f"orc_chain_run rc={rc}")
    out.recs, out.n_hist, out.model_num = recs, nh.value, mn.value
    return out


def chain_farm(p: OrcParams, d: Data, n_chains: int, n_iter: int, n_threads: int, seed: int = 1):
    phi = np.zeros(n_chains)
    K = np.zeros(n_chains, np.int32)
    acc = np.zeros(n_chains, np.int64)
    rc = lib().orc_chain_farm(C.byref(p), C.byref(d.c), n_chains, n_iter, n_threads, seed, _dp(phi), _ip(K),
                              acc.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc < 0:
        raise RuntimeError(f"orc_chain_farm rc={rc}")
    return phi, K, acc


def chain_farm_from(p: OrcParams, d: Data, K0, cells0, n_iter: int, n_threads: int, seed: int = 1):
    """chain_farm started from given models: K0[n], cells0[n, 4, Kcap] (x / y / z / zeta rows)."""
    K0 = np.ascontiguousarray(K0, dtype=np.int32)
    cells0 = np.ascontiguousarray(cells0, dtype=np.float64)
    n = len(K0)
    assert cells0.shape[:2] == (n, 4)
    phi = np.zeros(n)
    K = np.zeros(n, np.int32)
    acc = np.zeros(n, np.int64)
    rc = lib().orc_chain_farm_from(C.byref(p), C.byref(d.c), n, n_iter, n_threads, seed, _ip(K0), _dp(cells0), cells0.shape[2], _dp(phi),
                                   _ip(K), acc.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc < 0:
        raise RuntimeError(f"orc_chain_farm_from rc={rc}")
    return phi, K, acc


def ray_lengths(x, y, z, U):
    x, y, z, U = (np.asfortranarray(a, dtype=np.float64) for a in (x, y, z, U))
    m, R = x.shape
    rl = np.zeros((m - 1, R), order="F")
    ru = np.zeros((m - 1, R), order="F")
    lib().orc_ray_lengths(m, R, _dp(x), _dp(y), _dp(z), _dp(U), _dp(rl), _dp(ru))
    return rl, ru
