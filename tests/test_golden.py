"""Committed fixtures from the reference itself (tests/golden/*.npz, written by tools/make_golden.py
from oracle/_ref): the oracle must reproduce them on CPU, and the CUDA path (through the C ABI) must
reproduce them on the GPU.  Tolerances are stated at each assert; integer results are exact."""
import os

import numpy as np
import pytest

from oracle import pm_oracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name)))


def pofk_from_sums(g, box, nsample):
    n = g["pofk_n"]
    good = n > 0
    p = np.zeros_like(n)
    k = np.zeros_like(n)
    p[good] = g["pofk_sum"][good] / n[good] * box ** 3 - (box / nsample) ** 3
    k[good] = g["pofk_ksum"][good] / n[good] * 2 * np.pi / box
    return p, k, n


def step_kwargs(model, g):
    a, om, box = float(g["a"]), float(g["omega"]), float(g["box"])
    if model == "fofr":
        pc, c, m2 = po.fofr_scalars(a, om, box, float(g["par_fofr0"]), float(g["par_nfofr"]))
        return dict(omega=om, a=a, phi_crit=pc, coupling=c, massterm2=m2)
    if model == "dgp":
        c, f0 = po.dgp_scalars(a, om, float(g["par_rcH0_DGP"]))
        return dict(coupling=c, dgp_fac0=f0, rsmooth=float(g["par_Rsmooth_global"]))
    return None


# ----------------------------------------------------------------------------- CPU: oracle vs golden

@pytest.mark.parametrize("model", ["lcdm", "fofr", "dgp"])
def test_oracle_reproduces_reference_step(model):
    g = load("step_%s.npz" % model)
    N, box = int(g["N"]), float(g["box"])
    cfg = g["pofk_cfg"]
    pk = dict(nbins=int(cfg[0]), bintype=int(cfg[1]), subtract_shotnoise=int(cfg[2]), kmin_hmpc=float(cfg[3]), kmax_hmpc=float(cfg[4]))
    out = po.get_displacements(g["pos"], N, N, box, model="none" if model == "lcdm" else model, mg=step_kwargs(model, g), pofk=pk)
    assert np.abs(out["density_k"] - g["density_k"]).max() / np.abs(g["density_k"]).max() < 1e-13
    for a in range(3):
        assert np.abs(out["force_grids"][a] - g["force"][a]).max() / np.abs(g["force"][a]).max() < 1e-12
    # Disp is float32: identical up to the last bit where the double value sits on a rounding boundary
    assert np.abs(out["disp"] - g["disp"]).max() <= 2.5e-7 * np.abs(g["disp"]).max()
    assert np.allclose(out["sumDxyz"], g["sumDxyz"], rtol=0, atol=1e-9 * np.abs(g["disp"]).max())
    p, k, n = out["pofk"]
    pr, kr, nr = pofk_from_sums(g, box, N)
    assert np.array_equal(n, nr)
    assert np.allclose(p, pr, rtol=1e-11, atol=1e-11 * (box / N) ** 3)
    assert np.allclose(k, kr, rtol=1e-13)


def test_oracle_reproduces_reference_kick_drift():
    g = load("kickdrift.npz")
    v, d, sv = po.kick(g["vel"], g["disp"], g["D"], g["D2"], g["sumDxyz"], float(g["omega"]), 1, float(g["A"]), float(g["dda"]),
                       float(g["ddDddy"]), float(g["ddD2ddy"]))
    assert np.array_equal(v.view(np.uint32), g["vel_after"].view(np.uint32))
    assert np.array_equal(d.view(np.uint32), g["disp_after"].view(np.uint32))
    assert np.allclose(sv, g["sumxyz"], rtol=0, atol=1e-14)
    p = po.drift(g["pos"], v, g["D"], g["D2"], g["sumxyz"], float(g["box"]), 1, float(g["dyyy"]), float(g["deltaD"]), float(g["deltaD2"]))
    assert np.array_equal(p.view(np.uint32), g["pos_after"].view(np.uint32))


def _oracle_run(g, stepper):
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    pos, vel, D, D2 = g["pos0"].copy(), g["vel0"].copy(), g["D"], g["D2"]
    cfg = g["pofk_cfg"]
    pk = dict(nbins=int(cfg[0]), bintype=int(cfg[1]), subtract_shotnoise=int(cfg[2]), kmin_hmpc=float(cfg[3]), kmax_hmpc=float(cfg[4]))
    pks = []
    for (A, dda, ddD, ddD2, dyyy, dD, dD2) in g["steps"]:
        pc, c, m2 = po.fofr_scalars(A, om, box, float(g["fofr0"]), float(g["nfofr"]))
        out = po.get_displacements(pos, N, N, box, model="fofr", mg=dict(omega=om, a=A, phi_crit=pc, coupling=c, massterm2=m2), pofk=pk)
        vel, _, sv = po.kick(vel, out["disp"], D, D2, out["sumDxyz"], om, 1, A, dda, ddD, ddD2)
        pos = po.drift(pos, vel, D, D2, sv, box, 1, dyyy, dD, dD2)
        pks.append(out["pofk"])
    return pos, vel, pks


def test_oracle_reproduces_reference_run():
    """3 COLA steps with f(R) screening from the reference's own ICs (seed 5001)."""
    g = load("run_fofr.npz")
    N, box = int(g["N"]), float(g["box"])
    pos, vel, pks = _oracle_run(g, None)
    assert np.array_equal(g["id0"], g["id1"])
    dp = np.abs(pos.astype(np.float64) - g["pos1"])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < 2e-5 * box / N                   # float32 positions (ulp(60) = 3.8e-6): a few ulp
    assert np.abs(vel - g["vel1"]).max() < 1e-5 * np.abs(g["vel1"]).max()
    for it, (p, k, n) in enumerate(pks):
        s = g["pofk_sums"][it]
        assert np.array_equal(n, s[1])
        good = n > 0
        pr = s[0][good] / n[good] * box ** 3 - (box / N) ** 3
        assert np.allclose(p[good], pr, rtol=1e-6, atol=1e-6 * (box / N) ** 3)


# ----------------------------------------------------------------------------- GPU: CUDA path vs golden

@pytest.mark.gpu
@pytest.mark.parametrize("model", ["lcdm", "fofr", "dgp"])
@pytest.mark.parametrize("gb", [8, 4])
def test_cuda_reproduces_reference_step(mgp, require_gpu, model, gb):
    g = load("step_%s.npz" % model)
    N, box = int(g["N"]), float(g["box"])
    mid = {"lcdm": mgp.MODEL_NONE, "fofr": mgp.MODEL_FOFR, "dgp": mgp.MODEL_DGP}[model]
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), model=mid, include_screening=1, grid_bytes=gb)
    cfg = g["pofk_cfg"]
    pm.set_pofk(int(cfg[0]), int(cfg[1]), int(cfg[2]), float(cfg[3]), float(cfg[4]))
    pm.upload_particles(g["pos"], g["vel"], g["D"], g["D2"])
    kw = step_kwargs(model, g) or {}
    kw.pop("omega", None)
    s = pm.scalars(compute_pofk=1, **({"a": float(g["a"])} | kw))
    sumD = pm.GetDisplacements(s)
    got = pm.download_particles()
    order = np.argsort(got["id"])
    disp = pm.download_disp()[order]
    tol = 3e-7 if gb == 8 else 2e-4        # f64 grids: float32 Disp rounding only; f32 grids: single-precision FFTs
    assert np.abs(disp - g["disp"]).max() / np.abs(g["disp"]).max() < tol
    assert np.abs(sumD - g["sumDxyz"]).max() < tol * np.abs(g["disp"]).max()
    p, k, n = pm.step_power_spectrum()
    pr, kr, nr = pofk_from_sums(g, box, N)
    assert np.array_equal(n, nr)           # mode counts: exact
    good = nr > 0
    rel = np.abs(p[good] - pr[good]) / (np.abs(pr[good]) + (box / N) ** 3)
    assert rel.max() < (1e-11 if gb == 8 else 1e-4)
    pm.close()


@pytest.mark.gpu
def test_cuda_reproduces_reference_kick_drift(mgp, require_gpu):
    """Bit-exact Kick / Drift against the reference's own output."""
    g = load("kickdrift.npz")
    N, box = int(g["N"]), float(g["box"])
    pm = mgp.PM(N, N, box, omega=float(g["omega"]), sort_particles=0)
    pm.upload_particles(g["pos"], g["vel"], g["D"], g["D2"])
    pm.upload_disp(g["disp"])
    sv = pm.Kick(float(g["A"]), float(g["dda"]), float(g["ddDddy"]), float(g["ddD2ddy"]), sumDxyz=g["sumDxyz"])
    got = pm.download_particles()
    assert np.array_equal(got["vel"].view(np.uint32), g["vel_after"].view(np.uint32))
    assert np.allclose(sv, g["sumxyz"], rtol=0, atol=1e-13)
    pm.Drift(float(g["dyyy"]), float(g["deltaD"]), float(g["deltaD2"]), sumxyz=g["sumxyz"])
    got = pm.download_particles()
    assert np.array_equal(got["pos"].view(np.uint32), g["pos_after"].view(np.uint32))
    pm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gb", [8, 4])
def test_cuda_reproduces_reference_run(mgp, require_gpu, gb):
    """3 full COLA steps (f(R) + screening, in-step P(k)) from the reference's ICs: particle IDs exact,
    positions / velocities / P(k) within float32 tolerance of the reference's own result."""
    g = load("run_fofr.npz")
    N, box, om = int(g["N"]), float(g["box"]), float(g["omega"])
    pm = mgp.PM(N, N, box, omega=om, model=mgp.MODEL_FOFR, include_screening=1, grid_bytes=gb)
    cfg = g["pofk_cfg"]
    pm.set_pofk(int(cfg[0]), int(cfg[1]), int(cfg[2]), float(cfg[3]), float(cfg[4]))
    pm.upload_particles(g["pos0"], g["vel0"], g["D"], g["D2"], g["id0"])
    for it, (A, dda, ddD, ddD2, dyyy, dD, dD2) in enumerate(g["steps"]):
        pc, c, m2 = po.fofr_scalars(A, om, box, float(g["fofr0"]), float(g["nfofr"]))
        pm.GetDisplacements(pm.scalars(a=A, phi_crit=pc, coupling=c, massterm2=m2, compute_pofk=1))
        p, k, n = pm.step_power_spectrum()
        s = g["pofk_sums"][it]
        assert np.array_equal(n, s[1])
        good = n > 0
        pr = s[0][good] / n[good] * box ** 3 - (box / N) ** 3
        assert np.allclose(p[good], pr, rtol=1e-4, atol=1e-4 * (box / N) ** 3)     # north-star P(k) tolerance
        pm.Kick(A, dda, ddD, ddD2)
        pm.Drift(dyyy, dD, dD2)
    got = pm.download_particles()
    order = np.argsort(got["id"])
    assert np.array_equal(got["id"][order], np.sort(g["id1"]))                      # particle IDs: exact
    ref_order = np.argsort(g["id1"])
    dp = np.abs(got["pos"][order].astype(np.float64) - g["pos1"][ref_order])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < (2e-5 if gb == 8 else 2e-3) * box / N
    assert np.abs(got["vel"][order] - g["vel1"][ref_order]).max() < (1e-5 if gb == 8 else 1e-3) * np.abs(g["vel1"]).max()
    pm.close()
