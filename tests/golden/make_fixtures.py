"""Regenerate the data files derived from the reference's shipped data: the package's input data set and the test fixture.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_fixtures.py

Outputs
  mcmc-in-tonga_b200/tonga_b200/datasets/tonga381.npz   (package input data, loaded by tonga_b200.data.load_tonga381)
                 the 381-ray Tonga geometry + observations (Data/381raypaths.jld, Data/381traces.jld)
                 and the ak135 velocity table (Data/ak135f.txt) used to synthesise the slowness the
                 shipped raypaths file lacks.
  tests/golden/model_jld.npz  the 100 stored models of the reference's only shipped *output*, model.jld (2 chains x 50),
                 It belongs to an unshipped 487-ray data set, so the forward model cannot be replayed on it,
                 but (ptS, tS) -> phi and the constant "likelihood" (MCsub.jl:169-182) can: with allSig = 0.2 the stored phi
                 of all 100 models and the likelihood are reproduced bit for bit (tests/test_oracle.py).

These are DATA (inputs/outputs of the reference), not reference source.
"""
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATASETS = os.path.join(HERE, "..", "..", "mcmc-in-tonga_b200", "tonga_b200", "datasets")
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
from jld_min import JLDFile  # noqa: E402

REF = os.environ.get("TONGA_REFERENCE", "/root/reference")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def tonga381():
    rp = os.path.join(REF, "Data/381raypaths.jld")
    tp = os.path.join(REF, "Data/381traces.jld")
    r = JLDFile(rp)
    # HDF5 dims are the reverse of the Julia dims: the file holds Julia 381x131 (ray x point) arrays,
    # i.e. numpy (131, 381) = [point, ray], which is the point x ray orientation evaluate() expects.
    x, y, z = r.f64("x_n"), r.f64("y_n"), r.f64("z_n")
    assert x.shape == (131, 381)
    nan = np.isnan(x)
    assert (nan == np.isnan(y)).all() and (nan == np.isnan(z)).all()
    npts = (~nan).sum(0).astype(np.int32)
    for i in range(381):  # NaN padding is tail-only
        assert not nan[:npts[i], i].any() and nan[npts[i]:, i].all()
    # fingerprints recorded in SURVEY.md 5.9
    assert abs(np.nansum(x) - 11280783.739) < 1e-3
    assert abs(np.nansum(y) - 1709944.4563) < 1e-3
    assert abs(np.nansum(z) - 4252266.1811) < 1e-3
    flat = lambda a: np.concatenate([a[:npts[i], i] for i in range(381)])
    t = JLDFile(tp)
    out = dict(
        npts=npts, px=flat(x), py=flat(y), pz=flat(z),
        tStar=t.f64_via_refs("tStar"), error=t.f64_via_refs("error"),
        EventDepth=t.f64_via_refs("EventDepth"),
        ak135=np.loadtxt(os.path.join(REF, "Data/ak135f.txt"), delimiter=","),
        sha256_raypaths=np.array(sha(rp)), sha256_traces=np.array(sha(tp)),
    )
    assert abs(out["tStar"].sum() - 181.152908) < 1e-6
    assert abs(out["error"].sum() - 100.524106) < 1e-6
    np.savez_compressed(os.path.join(DATASETS, "tonga381.npz"), **out)
    print("tonga381.npz: P =", int(npts.sum()), "rays =", len(npts))


def model_jld():
    mp = os.path.join(REF, "model.jld")
    m = JLDFile(mp)
    chains = m.refs("model")
    recs = []
    for c, cref in enumerate(chains):
        for j, mref in enumerate(m.refs(int(cref))):
            raw = m.raw(int(mref))
            assert len(raw) == 104, len(raw)
            nC, = struct.unpack_from("<d", raw, 0)
            rx, ry, rz, rzeta = struct.unpack_from("<4Q", raw, 8)
            phi, = struct.unpack_from("<d", raw, 40)
            rptS, rtS = struct.unpack_from("<2Q", raw, 48)
            like, = struct.unpack_from("<d", raw, 64)
            action, accept = struct.unpack_from("<2q", raw, 72)
            zxz, zxy = struct.unpack_from("<2d", raw, 88)
            recs.append(dict(chain=c, j=j, nCells=nC, x=m.f64(rx).ravel(), y=m.f64(ry).ravel(),
                             z=m.f64(rz).ravel(), zeta=m.f64(rzeta).ravel(), phi=phi,
                             ptS=m.f64(rptS).ravel(), tS=m.f64(rtS).ravel(), likelihood=like,
                             action=action, accept=accept, zeta_xz=zxz, zeta_xy=zxy))
    n = len(recs)
    kmax = max(len(r["x"]) for r in recs)
    R = len(recs[0]["ptS"])
    cells = np.full((n, 4, kmax), np.nan)
    for i, r in enumerate(recs):
        k = len(r["x"])
        cells[i, 0, :k], cells[i, 1, :k], cells[i, 2, :k], cells[i, 3, :k] = r["x"], r["y"], r["z"], r["zeta"]
    out = dict(
        chain=np.array([r["chain"] for r in recs], np.int32),
        nCells=np.array([r["nCells"] for r in recs]),
        ncell_len=np.array([len(r["x"]) for r in recs], np.int32),
        cells=cells,
        phi=np.array([r["phi"] for r in recs]),
        likelihood=np.array([r["likelihood"] for r in recs]),
        action=np.array([r["action"] for r in recs], np.int64),
        accept=np.array([r["accept"] for r in recs], np.int64),
        zeta_xz=np.array([r["zeta_xz"] for r in recs]),
        zeta_xy=np.array([r["zeta_xy"] for r in recs]),
        ptS=np.stack([r["ptS"] for r in recs]),  # float64: (ptS, tS) -> phi is a reference-produced known answer (allSig = 0.2)
        tS=recs[0]["tS"],
        tS_identical=np.array(all((r["tS"] == recs[0]["tS"]).all() for r in recs)),
        sha256=np.array(sha(mp)),
    )
    np.savez_compressed(os.path.join(HERE, "model_jld.npz"), **out)
    print("model_jld.npz:", n, "models, kmax", kmax, "R", R)


if __name__ == "__main__":
    tonga381()
    model_jld()
