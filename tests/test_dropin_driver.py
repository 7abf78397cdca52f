"""The drop-in itself: the reference's own C driver (main.c + cosmo.c + read_param.c + ...) linked
against libmgpicola_cuda.so through adapter/auxPM_cuda.c, run on the same parameter file as the
unmodified CPU reference (oracle/_ref), outputs compared: in-code P(k) files of every step and the
final GADGET snapshot (IDs exact, positions / velocities within float tolerance)."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_gadget(path):
    """GADGET-1 snapshot as written by Output() (main.c:915-997): header, pos, vel, ids (u64 with PARTICLE_ID)."""
    with open(path, "rb") as f:
        def block():
            n = struct.unpack("i", f.read(4))[0]
            b = f.read(n)
            assert struct.unpack("i", f.read(4))[0] == n
            return b
        hdr = block()
        npart = struct.unpack("6I", hdr[:24])[1]
        pos = np.frombuffer(block(), np.float32).reshape(-1, 3)
        vel = np.frombuffer(block(), np.float32).reshape(-1, 3)
        ids = np.frombuffer(block(), np.uint64)
        assert pos.shape[0] == npart == ids.size
    return pos, vel, ids


def read_pofk(path):
    return np.loadtxt(path, comments="#").reshape(-1, 4)


def _exe(kind, variant):
    p = os.path.join(ROOT, "oracle", "_ref", "MG_PICOLA_%s" % variant) if kind == "cpu" else \
        os.path.join(ROOT, "adapter", "_build", "MG_PICOLA_CUDA_%s" % variant)
    if not os.path.exists(p):
        pytest.skip("%s not built (needs /root/reference at build time)" % p)
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("variant,model,tolx", [("lcdm", "fofr", 3e-5), ("dgp", "dgp", 3e-5), ("lcdm_sp", "fofr", 5e-3)])
def test_driver_with_cuda_library_matches_cpu_reference(require_gpu, tmp_path, variant, model, tolx):
    import bench
    N, box, nsteps = 32, 100.0, 6
    runs = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, model, nsteps)
        r = subprocess.run([_exe(kind, variant), pf], capture_output=True, text=True, cwd=wd, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[kind] = wd
    out_c, out_g = os.path.join(runs["cpu"], "output"), os.path.join(runs["gpu"], "output")
    pk_c = sorted(f for f in os.listdir(out_c) if f.startswith("pofk_") and f.endswith("_CDM.txt"))
    pk_g = sorted(f for f in os.listdir(out_g) if f.startswith("pofk_") and f.endswith("_CDM.txt"))
    assert pk_c == pk_g and len(pk_c) >= nsteps              # one file per force evaluation, same names
    shot = (box / N) ** 3
    for f in pk_c:
        a, b = read_pofk(os.path.join(out_c, f)), read_pofk(os.path.join(out_g, f))
        assert a.shape == b.shape
        assert np.array_equal(a[:, 0], b[:, 0])               # bin centres
        # files carry %10.5f: compare to the print resolution plus the stated P(k) tolerance
        assert np.all(np.abs(a[:, 1] - b[:, 1]) <= 2e-5 + (1e-4 if variant == "lcdm_sp" else 1e-8) * (np.abs(a[:, 1]) + shot))
        assert np.all(np.abs(a[:, 2] - b[:, 2]) <= 2e-5)
    snap = [f for f in os.listdir(out_c) if f.startswith("bench_z0p000")]
    assert snap
    pc, vc, ic = read_gadget(os.path.join(out_c, snap[0]))
    pg, vg, ig = read_gadget(os.path.join(out_g, snap[0]))
    oc, og = np.argsort(ic), np.argsort(ig)
    assert np.array_equal(ic[oc], ig[og]) and np.array_equal(ic[oc], np.arange(N ** 3, dtype=np.uint64))
    dp = np.abs(pc[oc].astype(np.float64) - pg[og])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < tolx * box / N
    assert np.abs(vc[oc] - vg[og]).max() < tolx * np.abs(vc).max() * 10


@pytest.mark.gpu
@pytest.mark.parametrize("variant,merged", [("fofr", 0), ("fofr", 1), ("dgp_sd", 0)])
def test_scale_dependent_driver_matches_cpu_reference(require_gpu, tmp_path, variant, merged):
    """SCALEDEPENDENT builds (MODEL=FOFR, DGP -DSCALEDEPENDENT) with use_lcdm_growth_factors = 0 and RSD multipoles
    every step: the driver's own assign_displacment_field_to_particles / compute_RSD_powerspectrum call sites,
    served by the CUDA library, against the unmodified CPU reference."""
    import bench
    N, box, nsteps = 32, 100.0, 5
    model = "fofr" if variant == "fofr" else "dgp"
    runs = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, model, nsteps, lcdm_growth=0)
        txt = open(pf).read().replace("pofk_compute_rsd_pofk 0", "pofk_compute_rsd_pofk 1")
        open(pf, "w").write(txt)
        env = dict(os.environ, MGP_SD_MERGED=str(merged))
        r = subprocess.run([_exe(kind, variant), pf], capture_output=True, text=True, cwd=wd, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[kind] = wd
    out_c, out_g = os.path.join(runs["cpu"], "output"), os.path.join(runs["gpu"], "output")
    shot = (box / N) ** 3
    for suffix, ncol in (("_CDM.txt", 4), (".txt", 7)):
        fc = sorted(f for f in os.listdir(out_c) if f.startswith("pofk_") and f.endswith(suffix) and (ncol == 4 or "RSD" in f))
        fg = sorted(f for f in os.listdir(out_g) if f.startswith("pofk_") and f.endswith(suffix) and (ncol == 4 or "RSD" in f))
        assert fc == fg and len(fc) >= nsteps - 1
        for f in fc:
            a = np.loadtxt(os.path.join(out_c, f), comments="#").reshape(-1, ncol)
            b = np.loadtxt(os.path.join(out_g, f), comments="#").reshape(-1, ncol)
            assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
            # P(k) [, k_mean, Delta] or P0, P2, P4: print resolution + 1e-6 relative; the merged fields differ from the
            # reference's two separately rounded floats by one float32 ulp per particle and step: 2e-5 (north star: 1e-4)
            rel = 2e-5 if merged else 1e-6
            for col in range(1, 4):
                assert np.all(np.abs(a[:, col] - b[:, col]) <= 2e-5 + rel * (np.abs(a[:, col]) + shot)), (f, col)
    snap = [f for f in os.listdir(out_c) if f.startswith("bench_z0p000")]
    assert snap
    pc, vc, ic = read_gadget(os.path.join(out_c, snap[0]))
    pg, vg, ig = read_gadget(os.path.join(out_g, snap[0]))
    oc, og = np.argsort(ic), np.argsort(ig)
    assert np.array_equal(ic[oc], ig[og])
    dp = np.abs(pc[oc].astype(np.float64) - pg[og])
    dp = np.minimum(dp, box - dp)
    assert dp.max() < (3e-4 if merged else 3e-5) * box / N
    assert np.abs(vc[oc] - vg[og]).max() < (3e-3 if merged else 3e-4) * np.abs(vc).max()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,model,lcdm_growth,ranks", [("lcdm", "fofr", 1, 2), ("fofr", "fofr", 0, 2), ("lcdm", "fofr", 1, 8)])
def test_driver_on_ranks_matches_cpu_reference_on_ranks(require_gpu, tmp_path, variant, model, lcdm_growth, ranks):
    """The reference's C driver as the multi-rank program it is, one rank per GPU, bound to the CUDA library
    (adapter/_build/MG_PICOLA_CUDA_<v>_mp, ranks = processes of the multi-process MPI stand-in, NCCL id handed round
    with MPI_Bcast) against the unmodified CPU reference on the same number of ranks: every P(k) file, and per rank
    the snapshot file -- the SAME particle IDs on the same rank (slab ownership, auxPM.c:151-153) at positions within
    the single-GPU tolerance.  Every rank takes GPU `rank % device count` by itself (adapter/auxPM_cuda.c)."""
    import torch
    if torch.cuda.device_count() < ranks:
        pytest.skip("needs %d GPUs" % ranks)
    import bench
    from oracle import mprun
    N, box, nsteps = 32, 100.0, 5
    exe = {"cpu": mprun.exe_path(variant), "gpu": os.path.join(ROOT, "adapter", "_build", "MG_PICOLA_CUDA_%s_mp" % variant)}
    out = {}
    for kind in ("cpu", "gpu"):
        if not os.path.exists(exe[kind]):
            pytest.skip("%s not built (needs /root/reference at build time)" % exe[kind])
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, model, nsteps, lcdm_growth=lcdm_growth)
        rc, so, errs = mprun.run([exe[kind], pf], ranks, scratch_mb=mprun.scratch_mb_for(N), timeout=900, cwd=wd)
        assert rc == 0, (kind, rc, so[-2000:], errs)
        out[kind] = os.path.join(wd, "output")
    shot = (box / N) ** 3
    pk_c = sorted(f for f in os.listdir(out["cpu"]) if f.startswith("pofk_") and f.endswith("_CDM.txt"))
    pk_g = sorted(f for f in os.listdir(out["gpu"]) if f.startswith("pofk_") and f.endswith("_CDM.txt"))
    assert pk_c == pk_g and len(pk_c) >= nsteps
    for f in pk_c:
        a, b = read_pofk(os.path.join(out["cpu"], f)), read_pofk(os.path.join(out["gpu"], f))
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0])
        assert np.all(np.abs(a[:, 1] - b[:, 1]) <= 2e-5 + 1e-8 * (np.abs(a[:, 1]) + shot))
    # per rank: the same particle IDs (slab ownership) at positions within the single-GPU tolerance; a particle within that
    # tolerance of a slab boundary may legitimately sit on either side
    rank_c, rank_g = np.zeros(N ** 3, np.int64), np.zeros(N ** 3, np.int64)
    pos_c, pos_g = np.zeros((N ** 3, 3)), np.zeros((N ** 3, 3))
    seen_c, seen_g = np.zeros(N ** 3, bool), np.zeros(N ** 3, bool)
    for r in range(ranks):
        pc, vc, ic = read_gadget(os.path.join(out["cpu"], "bench_z0p000.%d" % r))
        pg, vg, ig = read_gadget(os.path.join(out["gpu"], "bench_z0p000.%d" % r))
        assert not seen_c[ic].any() and not seen_g[ig].any()
        seen_c[ic], seen_g[ig] = True, True
        rank_c[ic], rank_g[ig] = r, r
        pos_c[ic], pos_g[ig] = pc, pg
    assert seen_c.all() and seen_g.all()                      # nobody lost, nobody duplicated
    tol = 3e-5 * box / N
    dp = np.abs(pos_c - pos_g)
    dp = np.minimum(dp, box - dp)
    assert dp.max() < tol
    moved = rank_c != rank_g
    if moved.any():
        xs = pos_c[moved, 0] * N / box                        # cells; slab boundaries are multiples of N / ranks cells
        edge = np.abs(xs / (N // ranks) - np.round(xs / (N // ranks))) * (N // ranks)
        assert (edge < 2 * tol * N / box).all()
    assert moved.sum() <= 4


@pytest.mark.gpu
def test_lightcone_driver_matches_cpu_reference(require_gpu, tmp_path):
    """-DLIGHTCONE -DUNFORMATTED build: the reference's driver with Drift_Lightcone's particle loop served by
    mgp_drift_lightcone (adapter/lightcone_c.patch) against the unmodified reference on a whole lightcone run (three
    ordinary steps down to the redshift the cone starts at, four lightcone steps to z = 0, 125 replicates).  Every image
    the reference writes must be written by the CUDA build too, replicate file by replicate file, at the same exit
    position and velocity; the info file (lightcone.c:570-646) must agree."""
    from scipy.spatial import cKDTree
    import lightcone_case as lcc
    N, box = 32, 100.0
    runs = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = lcc.write_lightcone_paramfile(wd, N, box, 0.06, 3, 4)
        r = subprocess.run([_exe(kind, "lcdm_lc"), pf], capture_output=True, text=True, cwd=wd, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[kind] = os.path.join(wd, "output")
    info_c = np.loadtxt(os.path.join(runs["cpu"], "bench_lightcone.info"), comments="#").reshape(-1, 8)
    info_g = np.loadtxt(os.path.join(runs["gpu"], "bench_lightcone.info"), comments="#").reshape(-1, 8)
    assert np.array_equal(info_c[:, :7], info_g[:, :7])                    # file numbers and slice corners
    assert info_c[:, 7].sum() > 50000
    # a particle within float rounding of the cone's surface may leave one step earlier or later, or start just outside
    assert np.all(np.abs(info_c[:, 7] - info_g[:, 7]) <= 2 + 1e-3 * info_c[:, 7])
    files_c = sorted(f for f in os.listdir(runs["cpu"]) if "_lightcone." in f and not f.endswith(".info"))
    files_g = sorted(f for f in os.listdir(runs["gpu"]) if "_lightcone." in f and not f.endswith(".info"))
    assert set(files_c) ^ set(files_g) <= {f for f in set(files_c) | set(files_g)
                                            if info_c[int(f.rsplit(".", 1)[1]), 7] + info_g[int(f.rsplit(".", 1)[1]), 7] <= 2}
    total = unmatched = 0
    for f in sorted(set(files_c) & set(files_g)):
        a, b = lcc.read_lightcone_file(os.path.join(runs["cpu"], f)), lcc.read_lightcone_file(os.path.join(runs["gpu"], f))
        n = int(f.rsplit(".", 1)[1])
        assert a.shape[0] == info_c[n, 7] and b.shape[0] == info_g[n, 7]      # Noutput bookkeeping (lightcone.c:456)
        if a.shape[0] == 0 or b.shape[0] == 0:
            unmatched += a.shape[0]
            continue
        d, j = cKDTree(b[:, :3].astype(np.float64)).query(a[:, :3].astype(np.float64))
        good = (d < 3e-5 * box / N * 20) & (np.abs(a[:, 3:] - b[j, 3:]).max(axis=1) < 3e-4 * np.abs(a[:, 3:]).max() * 10)
        total += a.shape[0]
        unmatched += int((~good).sum())
    assert total > 50000 and unmatched <= 5 + 2e-4 * total, (total, unmatched)


@pytest.mark.gpu
def test_matchmaker_driver_matches_cpu_reference(require_gpu, tmp_path):
    """-DMATCHMAKER_HALOFINDER build: main.c's FoF hook (main.c:830-872) served by mgp_fof_find, the catalogue written by
    the reference's own mm_snap_io.c.  (a) Exact: the numpy restatement of MatchMaker applied to the positions of the
    CUDA driver's own snapshot must give the catalogue the CUDA driver wrote -- np, centres, rms radii and axis ratios bit
    for bit (the snapshot holds the very floats the finder saw; its velocities carry -<Vel>, so velocity moments are
    compared to that accuracy only).  (b) Against the unmodified reference's run: positions differ by float rounding after
    ten steps, so a few marginal links differ: the same number of halos and the same mass in them to a per cent."""
    import bench
    import fof_case as fc
    from oracle import pm_oracle as po
    N, box, nsteps = 64, 100.0, 10
    runs = {}
    for kind in ("cpu", "gpu"):
        wd = str(tmp_path / kind)
        pf = bench.write_paramfile(wd, N, box, "lcdm", nsteps, extra=fc.MM_TAGS)
        r = subprocess.run([_exe(kind, "lcdm_mm"), pf], capture_output=True, text=True, cwd=wd, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        runs[kind] = os.path.join(wd, "output")
    hc = fc.read_halo_file(os.path.join(runs["cpu"], "matchmaker_bench_z0.000.dat"))
    hg = fc.read_halo_file(os.path.join(runs["gpu"], "matchmaker_bench_z0.000.dat"))
    assert hc.size > 200
    # (a) the catalogue of the CUDA run is what MatchMaker gives on the particles of the CUDA run
    snap = [f for f in os.listdir(runs["gpu"]) if f.startswith("bench_z0p000")]
    pos, vel, ids = read_gadget(os.path.join(runs["gpu"], snap[0]))
    zero = np.zeros_like(pos)
    Hubble = 100.0                                          # HUBBLE * UnitTime_in_s with the units of the parameter file
    mass_part = 3.0 * bench.OMEGA * Hubble * Hubble * box ** 3 / (8.0 * np.pi * 43.0071 * float(N) ** 3)   # main.c:853
    out, _ = po.fof_halos([dict(pos=pos, vel=vel, D=zero, D2=zero, local_p_start=0)], 1.0, 1.0, box, 3.0, 0.2, 20, mass_part, N, 0)
    a, b = fc.canonical(out[0]), fc.canonical(hg)
    assert a.size == b.size and np.array_equal(a["np"], b["np"])
    for f in ("x_avg", "x_rms", "b", "c"):
        assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), f
    assert np.allclose(a["m_halo"], b["m_halo"], rtol=1e-4)
    vscale = np.abs(b["v_rms"]).max()
    assert np.abs(a["v_avg"] - b["v_avg"]).max() < 2e-3 * vscale and np.abs(a["v_rms"] - b["v_rms"]).max() < 2e-3 * vscale
    # (b) against the unmodified reference
    assert abs(hc.size - hg.size) <= 2 + 0.01 * hc.size
    assert abs(int(hc["np"].sum()) - int(hg["np"].sum())) <= 0.01 * hc["np"].sum()
    assert abs(int(hc["np"][0]) - int(hg["np"][0])) <= 0.02 * hc["np"][0]
