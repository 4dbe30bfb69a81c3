"""ctypes binding of libtonga_b200.so (include/tonga_b200.h) -- the same calls the Julia shim makes with `ccall`.

The library is built in-tree (mcmc-in-tonga_b200/lib/libtonga_b200.so) by `__graft_entry__.build()` /
`make -C mcmc-in-tonga_b200/csrc`.  There is no fallback: if the library is missing, or no CUDA device is
present when a compute entry point is called, a TongaError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TONGA_B200_LIB") or os.path.join(HERE, "..", "lib", "libtonga_b200.so")  # same override as julia/TongaB200.jl

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_bp = C.POINTER(C.c_int8)
c_vpp = C.POINTER(C.c_void_p)


class TongaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libtonga_b200 error {code}: {msg}")
        self.code = code


class TongaParams(C.Structure):
    """`tonga_params` (include/tonga_b200.h) <- hot-path subset of `struct parameters`, define_TDstructure.jl:1-44."""
    _fields_ = [(n, C.c_double) for n in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax", "sig", "zeta_scale",
                                          "max_sig", "n_iter", "burn_in", "keep_each")] + \
               [(n, C.c_int32) for n in ("min_cells", "max_cells", "prior", "debug_prior", "interp_style", "n_actions")]


PROPOSAL_DTYPE = np.dtype([("action", "<i4"), ("idx", "<i4"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8"),
                           ("zeta", "<f8"), ("u", "<f8")])

# every symbol include/tonga_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "tonga_last_error": (C.c_char_p, []),
    "tonga_version": (C.c_int, []),
    "tonga_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "tonga_create": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                               C.POINTER(TongaParams), C.c_int32]),
    "tonga_create_from_points": (C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.POINTER(TongaParams), C.c_int32]),
    "tonga_destroy": (None, [_P]),
    "tonga_info": (C.c_int, [_P, c_ip, c_lp, c_lp, c_lp]),
    "tonga_ray_offsets": (C.c_int, [_P, c_ip]),
    "tonga_synchronize": (C.c_int, [_P]),
    "tonga_set_exact_only": (C.c_int, [_P, C.c_int32]),
    "tonga_evaluate": (C.c_int, [_P, C.c_int32, c_dp, c_dp, c_dp, c_dp, C.c_double, c_dp, c_dp, c_dp, c_dp]),
    "tonga_evaluate_batch": (C.c_int, [_P, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "tonga_misfit": (C.c_int, [_P, C.c_int32, c_dp, c_dp, c_dp]),
    "tonga_evaluate_batch_dev": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "tonga_interpolate": (C.c_int, [_P, C.c_int32, c_dp, c_dp, c_dp, c_dp, C.c_int32, c_dp, C.c_int32, c_dp, C.c_int32, c_dp,
                                    c_dp, c_ip, c_ip]),
    "tonga_chains_create": (C.c_int, [_P, C.POINTER(_P), C.c_int32, C.c_int64, C.c_uint64, C.c_int32]),
    "tonga_chains_create_ex": (C.c_int, [_P, C.POINTER(_P), C.c_int32, C.c_int64, C.c_uint64, C.c_int32, C.c_int32]),
    "tonga_chains_sampler": (C.c_int, [_P]),
    "tonga_chains_history_host": (C.c_int, [_P] + [c_vpp] * 8),
    "tonga_chains_destroy": (None, [_P]),
    "tonga_chains_build_starting": (C.c_int, [_P]),
    "tonga_chains_set_models": (C.c_int, [_P, C.c_int32, c_ip, c_dp, c_dp]),
    "tonga_chains_set_exact_only": (C.c_int, [_P, C.c_int32]),
    "tonga_chains_set_beta": (C.c_int, [_P, c_dp]),
    "tonga_chains_get_beta": (C.c_int, [_P, c_dp]),
    "tonga_chains_temper_swap": (C.c_int, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_uint64]),
    "tonga_chains_temper_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    "tonga_chains_scalar_ptrs": (C.c_int, [_P] + [c_vpp] * 3),
    "tonga_chains_shard_init": (C.c_int, [_P, C.c_int32, C.c_int32, c_vpp, C.POINTER(C.c_uint64)]),
    "tonga_chains_shard_connect": (C.c_int, [_P, c_vpp]),
    "tonga_chains_shard_info": (C.c_int, [_P, c_ip, c_ip, c_ip, c_ip, c_lp, c_lp]),
    "tonga_shard_range": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_ip, c_ip]),
    "tonga_ipc_export": (C.c_int, [_P, C.c_char_p]),
    "tonga_ipc_open": (C.c_int, [C.c_int32, C.c_char_p, c_vpp]),
    "tonga_ipc_close": (C.c_int, [C.c_int32, _P]),
    "tonga_chains_run": (C.c_int, [_P, C.c_int64, C.c_int32, _P, c_bp, c_dp, c_ip]),
    "tonga_chains_get_state": (C.c_int, [_P, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "tonga_chains_get_stats": (C.c_int, [_P, c_lp, c_lp]),
    "tonga_chains_reset": (C.c_int, [_P]),
    "tonga_chains_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "tonga_chains_get_history": (C.c_int, [_P, C.c_int32, c_ip, c_ip, c_dp, c_dp, c_dp, c_lp, c_ip, c_ip, c_ip]),
    "tonga_chains_get_progress": (C.c_int, [_P, c_lp, c_lp, c_ip]),
    "tonga_chains_set_progress": (C.c_int, [_P, C.c_int64, c_lp, c_ip, c_lp]),
    "tonga_chains_set_history": (C.c_int, [_P, C.c_int32, c_ip, c_ip, c_dp, c_dp, c_dp, c_lp, c_ip, c_ip, c_ip]),
    "tonga_chains_raster": (C.c_int, [_P, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_lp]),
    "tonga_chains_verify": (C.c_int, [_P, c_lp, c_dp, c_dp]),
    "tonga_chains_kcap": (C.c_int, [_P]),
    "tonga_chains_device_ptrs": (C.c_int, [_P] + [c_vpp] * 8),
    "tonga_chains_profile": (C.c_int, [_P, C.c_int32, c_lp]),
    "tonga_host_alloc": (C.c_int, [C.POINTER(_P), C.c_uint64]),
    "tonga_host_free": (C.c_int, [_P]),
    "tonga_peak_flops": (C.c_int, [_P, c_dp, c_dp]),
}

_LIB = None


def load(path: str | None = None):
    """dlopen the in-tree library and bind every declared symbol (raises if the library or a symbol is missing)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = os.path.abspath(path or LIB_PATH)
    if not os.path.exists(p):
        raise TongaError(-2, f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             f"(there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise TongaError(rc, load().tonga_last_error().decode("utf-8", "replace"))


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def lp(a):
    return None if a is None else a.ctypes.data_as(c_lp)


def bp(a):
    return None if a is None else a.ctypes.data_as(c_bp)
