import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mcmc-in-tonga_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tonga():
    """The shipped 381-ray Tonga set as (DataStruct, parameters)."""
    from tonga_b200.data import load_tonga381
    from tonga_b200.structs import define_TDstructrure
    p = define_TDstructrure()
    return load_tonga381(p=p), p


def random_ragged(seed, R=23, m=17, box=(100.0, 80.0, 60.0), full_some=True, tS=None, sig=None):
    """Small random ragged ray set in the reference layout (m x R, NaN tail padding)."""
    from tonga_b200.data import make_datastruct
    from tonga_b200.structs import StepRangeLen, parameters
    rng = np.random.default_rng(seed)
    npts = rng.integers(2, m + 1, R)
    if full_some:
        npts[0] = m       # a ray with no NaN at all
        npts[1] = 2       # shortest ray that still has a segment
    x = np.full((m, R), np.nan, order="F"); y = x.copy(); z = x.copy()
    for i in range(R):
        n = npts[i]
        a = rng.uniform(0, 1, 3) * box
        b = rng.uniform(0, 1, 3) * box
        t = np.linspace(0, 1, n)[:, None]
        pt = a + t * (b - a) + rng.normal(0, 0.3, (n, 3))
        x[:n, i], y[:n, i], z[:n, i] = pt.T
    U = np.where(np.isnan(z), np.nan, 0.1 + 0.001 * np.nan_to_num(z))
    p = parameters()
    bx = (StepRangeLen(-10.0, 5.0, box[0] + 10), StepRangeLen(-10.0, 5.0, box[1] + 10), StepRangeLen(0.0, 5.0, box[2]))
    ds = make_datastruct(x, y, z, U, rng.uniform(0.1, 1.0, R) if tS is None else tS, rng.uniform(0.05, 0.5, R) if sig is None else sig, p, box=bx)
    return ds, p


def random_model(rng, K, box):
    return (rng.uniform(box[0], box[1], K), rng.uniform(box[2], box[3], K), rng.uniform(box[4], box[5], K),
            rng.uniform(0, 50, K))


def box_of(ds):
    return (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
