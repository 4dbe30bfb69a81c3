"""Host-side mirrors of the reference's data types, field for field and in the reference's order.

  parameters  <- `struct parameters`          /root/reference/define_TDstructure.jl:1-44
  DataStruct  <- `struct DataStruct`          /root/reference/DefStruct.jl:5-30
  Model       <- `mutable struct Model`       /root/reference/DefStruct.jl:32-48
  StepRangeLen<- Julia's `a:s:b` ranges used for xVec / yVec / zVec (load_data_Tonga.jl:47-49)

Julia is not available in this image, so this Python mirror plays the Julia caller above
the C-ABI library; `julia/TongaB200.jl` is the same shim written in Julia.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import numpy as np


class StepRangeLen:
    """`start:step:stop` as Julia builds it: length = floor((stop-start)/step)+1, last = start+(len-1)*step."""

    def __init__(self, start: float, step: float, stop: float):
        self.start = float(start)
        self.step = float(step)
        n = int(math.floor((float(stop) - self.start) / self.step + 1e-12)) + 1
        self.len = max(n, 0)

    def __len__(self):
        return self.len

    def __getitem__(self, i):
        if i < 0:
            i += self.len
        return self.start + i * self.step

    def vec(self) -> np.ndarray:
        return self.start + self.step * np.arange(self.len, dtype=np.float64)

    # min(xVec...) / max(xVec...) as used at TD_inversion_function.jl:30-32,78-80,230-232 and MCsub.jl:92-94
    def min(self) -> float:
        return self[0] if self.step >= 0 else self[self.len - 1]

    def max(self) -> float:
        return self[self.len - 1] if self.step >= 0 else self[0]

    def __repr__(self):
        return f"{self.start}:{self.step}:{self[self.len - 1]}"


@dataclass
class parameters:  # noqa: N801  (reference spelling)
    # ====basic parameters====  define_TDstructure.jl:3-5
    debug_prior: int = 0
    plot_voronoi: int = 0
    add_yVec: int = 1
    # ====Voronoi diagram parameters====  :8-19
    sig: int = 10
    zeta_scale: int = 50
    max_cells: int = 100
    min_cells: int = 5
    max_sig: float = 0.1
    interp_style: int = 1
    enforce_discon: int = 0
    prior: int = 1
    event_statics: int = 1
    demean: int = 1
    # =====Monte Carlo parameters=====  :23-27  (Float64 in the reference)
    n_chains: int = 2
    n_iter: float = 1e3
    burn_in: float = 5e2
    keep_each: float = 1e1
    print_each: float = 1e2
    # =====map parameters=====  :30-36
    max_depth: float = 660.0
    min_depth: float = 0.0
    rotation: int = 20
    ZnodeSpacing: int = 20
    buffer: int = 100
    XYnodeSpacing: int = 20
    # ====Cross section parameters====  :39-42
    xyMap: bool = True
    zSlice: List[int] = field(default_factory=lambda: [50, 300, 500])
    xzMap: bool = True
    ySlice: List[int] = field(default_factory=lambda: [700, 800])


def define_TDstructrure() -> parameters:  # noqa: N802  (sic, define_TDstructure.jl:46)
    """The literal values of define_TDstructure.jl:48-61."""
    return parameters()


@dataclass
class DataStruct:
    tS: np.ndarray
    allaveatten: np.ndarray
    allLats: np.ndarray
    allLons: np.ndarray
    allSig: np.ndarray
    dataX: np.ndarray
    dataY: np.ndarray
    xVec: StepRangeLen
    yVec: StepRangeLen
    zVec: StepRangeLen
    elonsX: np.ndarray
    elatsY: np.ndarray
    elons: np.ndarray
    elats: np.ndarray
    edep: np.ndarray
    coastX: np.ndarray
    coastY: np.ndarray
    rayX: np.ndarray  # m x R, Fortran order (a column is one ray), NaN tail padding
    rayY: np.ndarray
    rayZ: np.ndarray
    rayL: np.ndarray  # (m-1) x R
    rayU: np.ndarray
    U: np.ndarray


@dataclass
class Model:
    nCells: float
    xCell: np.ndarray
    yCell: np.ndarray
    zCell: np.ndarray
    zeta: np.ndarray
    phi: float = -1.0
    ptS: np.ndarray = field(default_factory=lambda: np.zeros(1))
    tS: np.ndarray = field(default_factory=lambda: np.zeros(1))
    likelihood: float = -1.0
    action: int = -1
    accept: int = -1
    zeta_xz: float = -1.0
    zeta_xy: float = -1.0

    def copy(self) -> "Model":
        return Model(self.nCells, self.xCell.copy(), self.yCell.copy(), self.zCell.copy(), self.zeta.copy(),
                     self.phi, self.ptS.copy(), self.tS, self.likelihood, self.action, self.accept,
                     self.zeta_xz, self.zeta_xy)
