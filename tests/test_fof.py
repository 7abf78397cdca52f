"""FoF halo finder on the fly (MatchMaker: mm_main.c, mm_fof.c).

Chain of evidence:
  CPU  the numpy restatement (oracle/pm_oracle.py::fof_halos) reproduces the UNMODIFIED reference
       (oracle/_ref/libmgpicola_ref_lcdm_mm.so, MatchMaker() called with the reference's own struct) on one task: every
       float of every FoFHalo record identical; and the reference run as a 2-task program (oracle/_ref/MG_PICOLA_lcdm_mm_mp)
       halo by halo, which pins the strip exchange and the redundancy rule;
  CPU  the GPU path's own sequence of steps and functors (csrc/fof_impl.cuh, csrc/fof.cuh), run on the host
       (tests/host/fof_emul.cu), give the reference's catalogue bit for bit on one task and the oracle's on 2 and 4
       emulated tasks, with and without COLA and with scale-dependent velocity fields;
  GPU  mgp_fof_find / mgp_fof_get against the reference's catalogue, bit for bit."""
import os
import struct

import numpy as np
import pytest

import fof_case as fc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N1D, BOX = 32, 100.0


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    pos, vel, D, D2 = fc.make_particles(N1D, BOX, 3)
    r = fc.reference_halos(str(tmp_path_factory.mktemp("fof")), pos, vel, D, D2, N1D, BOX)
    if r is None:
        pytest.skip("oracle/_ref MatchMaker build missing (make -C oracle; needs /root/reference)")
    ref, cfg = r
    assert ref.size >= 40 and ref["np"].min() >= cfg["np_min"] and np.any(ref["np"] == cfg["np_min"])
    return dict(pos=pos, vel=vel, D=D, D2=D2, ref=ref, cfg=cfg)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    lib = fc.build_emulation(str(tmp_path_factory.mktemp("fofemul")))
    if lib is None:
        pytest.skip("nvcc not available")
    return lib


def _oracle(tasks, cfg, use_cola=1, sd=False):
    from oracle import pm_oracle as po
    return po.fof_halos(tasks, cfg["norm_pos"], cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"],
                        cfg["mass_part"], N1D, use_cola, cfg["dDdy"], cfg["dD2dy"], sd)


def test_oracle_matches_reference_bit_for_bit(case):
    c = case
    out, info = _oracle([dict(pos=c["pos"], vel=c["vel"], D=c["D"], D2=c["D2"], local_p_start=0)], c["cfg"])
    fc.assert_same_halos(out[0], c["ref"], exact_vectors=False)       # numpy's eigh against the stand-in's Jacobi rotations
    assert info[0]["n_toleft"] > 500                                  # the strip (and its duplicates on one task) is in play


def test_oracle_without_cola_matches_reference(tmp_path):
    pos, vel, D, D2 = fc.make_particles(N1D, BOX, 5)
    r = fc.reference_halos(str(tmp_path), pos, vel, D, D2, N1D, BOX, use_cola=0)
    if r is None:
        pytest.skip("oracle/_ref MatchMaker build missing")
    ref, cfg = r
    out, _ = _oracle([dict(pos=pos, vel=vel, D=D, D2=D2, local_p_start=0)], cfg, use_cola=0)
    fc.assert_same_halos(out[0], ref, exact_vectors=False)


def _read_gadget_pos(path):
    with open(path, "rb") as f:
        def block():
            n = struct.unpack("i", f.read(4))[0]
            b = f.read(n)
            assert struct.unpack("i", f.read(4))[0] == n
            return b
        block()
        return np.frombuffer(block(), np.float32).reshape(-1, 3).copy()


def test_oracle_on_two_tasks_matches_reference_on_two_ranks(tmp_path):
    """The reference as a 2-rank program: a short run, halos per rank (mm_output_pernode 1, binary).  The oracle gets every
    rank's particle positions from the rank's own snapshot file (lengthfac = 1: the very floats MatchMaker saw) and must
    find the same halos on the same rank with the same position-only properties (the snapshot's velocities are not
    MatchMaker's, so velocity moments are not compared here; the one-task test pins them)."""
    from oracle import mprun
    import bench
    if not mprun.available("lcdm_mm"):
        pytest.skip("oracle/_ref multi-rank MatchMaker build missing")
    wd = str(tmp_path)
    tags = fc.MM_TAGS.replace("mm_output_pernode 0", "mm_output_pernode 1")
    pf = bench.write_paramfile(wd, N1D, BOX, "lcdm", 8, extra=tags)
    rc, out, errs = mprun.run([mprun.exe_path("lcdm_mm"), pf], 2, scratch_mb=mprun.scratch_mb_for(N1D), timeout=600)
    assert rc == 0, out[-2000:] + str(errs)
    od = os.path.join(wd, "output")
    tasks, refs = [], []
    for r in range(2):
        p = _read_gadget_pos(os.path.join(od, "bench_z0p000.%d" % r))
        z = np.zeros_like(p)
        tasks.append(dict(pos=p, vel=z, D=z, D2=z, local_p_start=r * (N1D // 2)))
        refs.append(fc.read_halo_file(os.path.join(od, "matchmaker_bench_z0.000.%d.dat" % r)))
    assert sum(h.size for h in refs) >= 10
    cfg = dict(fc.FOF_DEFAULTS, norm_vel=1.0, dDdy=0.0, dD2dy=0.0)
    from oracle import pm_oracle as po
    # particle mass as main.c:853 forms it does not enter the position-only fields; np_min, b_fof, dx_extra as the tags
    outs, info = po.fof_halos(tasks, 1.0, 1.0, BOX, 3.0, 0.2, 20, cfg["mass_part"], N1D, 1, 0.0, 0.0)
    for r in range(2):
        a, b = fc.canonical(outs[r]), fc.canonical(refs[r])
        assert a.size == b.size and np.array_equal(a["np"], b["np"])
        for f in ("x_avg", "x_rms", "b", "c"):
            assert np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32)), (r, f)


def test_gpu_steps_on_the_host_match_reference(case, emul):
    """csrc/fof_impl.cuh::find_halos with the host back end, one task: the reference's catalogue bit for bit, the
    eigenvectors included (the device code carries the same Jacobi rotations as the stand-in for gsl_eigen_symmv)."""
    c = case
    res = fc.emulated_halos(emul, [dict(pos=c["pos"], vel=c["vel"], D=c["D"], D2=c["D2"], local_p_start=0)], N1D, N1D, c["cfg"],
                            BOX * c["cfg"]["norm_pos"])
    fc.assert_same_halos(res[0], c["ref"])


@pytest.mark.parametrize("ntask", [2, 4])
@pytest.mark.parametrize("mode", ["cola", "nocola", "sd"])
def test_gpu_steps_on_emulated_tasks_match_oracle(case, emul, ntask, mode):
    c = case
    tasks = fc.split_tasks(c["pos"], c["vel"], c["D"], c["D2"], ntask, N1D, BOX)
    use_cola, sd = (0 if mode == "nocola" else 1), mode == "sd"
    orc, info = _oracle(tasks, c["cfg"], use_cola, sd)
    em = fc.emulated_halos(emul, tasks, N1D, N1D, c["cfg"], BOX * c["cfg"]["norm_pos"], use_cola, int(sd))
    assert sum(h.size for h in em) >= 40
    for r in range(ntask):
        fc.assert_same_halos(em[r], orc[r], exact_vectors=False)


@pytest.mark.parametrize("max_cells", [1, 300, 20000])
def test_gpu_steps_do_not_depend_on_the_search_cells(case, emul, max_cells, monkeypatch):
    """The search cells are as fine as memory allows (csrc/fof.cuh::geometry); with any budget -- one cell per direction,
    cells of several linking lengths -- the catalogue is the reference's bit for bit."""
    c = case
    monkeypatch.setenv("MGP_FOF_MAX_CELLS", str(max_cells))
    res = fc.emulated_halos(emul, [dict(pos=c["pos"], vel=c["vel"], D=c["D"], D2=c["D2"], local_p_start=0)], N1D, N1D, c["cfg"],
                            BOX * c["cfg"]["norm_pos"])
    fc.assert_same_halos(res[0], c["ref"])
    tasks = fc.split_tasks(c["pos"], c["vel"], c["D"], c["D2"], 2, N1D, BOX)
    monkeypatch.delenv("MGP_FOF_MAX_CELLS")
    fine = fc.emulated_halos(emul, tasks, N1D, N1D, c["cfg"], BOX * c["cfg"]["norm_pos"])
    monkeypatch.setenv("MGP_FOF_MAX_CELLS", str(max_cells))
    coarse = fc.emulated_halos(emul, tasks, N1D, N1D, c["cfg"], BOX * c["cfg"]["norm_pos"])
    for a, b in zip(fine, coarse):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_gpu_steps_edge_cases(emul):
    """No halo at all (uniform particles), np_min above every group, a rank without particles in its strip."""
    rng = np.random.default_rng(11)
    n = 16 ** 3
    pos = rng.uniform(0, BOX, (n, 3)).astype(np.float32)
    z = np.zeros((n, 3), np.float32)
    cfg = dict(fc.FOF_DEFAULTS, dx_extra=1.0)
    res = fc.emulated_halos(emul, [dict(pos=pos, vel=z, D=z, D2=z, local_p_start=0)], 16, 16, cfg, BOX)
    assert res[0].size == 0
    pos2, vel2, D, D2 = fc.make_particles(16, BOX, 9, nblobs=8)
    res = fc.emulated_halos(emul, [dict(pos=pos2, vel=vel2, D=D, D2=D2, local_p_start=0)], 16, 16, dict(cfg, np_min=100000), BOX)
    assert res[0].size == 0
    pos3 = pos2.copy()
    pos3[:, 0] = np.clip(pos3[:, 0], 10.0, 90.0)                       # nobody within dx_extra of x = 0
    from oracle import pm_oracle as po
    tasks = [dict(pos=pos3, vel=vel2, D=D, D2=D2, local_p_start=0)]
    orc, info = po.fof_halos(tasks, 1.0, cfg["norm_vel"], BOX, 1.0, 0.2, 20, cfg["mass_part"], 16, 1, cfg["dDdy"], cfg["dD2dy"])
    assert info[0]["n_toleft"] == 0
    res = fc.emulated_halos(emul, tasks, 16, 16, cfg, BOX)
    fc.assert_same_halos(res[0], orc[0], exact_vectors=False)


@pytest.mark.gpu
def test_cuda_halo_finder_matches_reference(mgp, require_gpu, case):
    c, cfg = case, case["cfg"]
    pm = mgp.PM(N1D, N1D, BOX, grid_bytes=8, use_cola=1, sort_particles=0)
    pm.upload_particles(c["pos"], c["vel"], c["D"], c["D2"])
    h = pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"],
                      cfg["mass_part"], cfg["dDdy"], cfg["dD2dy"])
    assert np.all(np.diff(h["np"]) <= 0)                               # by decreasing np (mm_fof.c:459)
    fc.assert_same_halos(h, c["ref"])
    # the particle store is untouched, and a second call gives the same catalogue (nothing depends on the thread schedule)
    got = pm.download_particles(("pos", "vel"))
    assert np.array_equal(got["pos"], c["pos"]) and np.array_equal(got["vel"], c["vel"])
    h2 = pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"],
                       cfg["mass_part"], cfg["dDdy"], cfg["dD2dy"])
    assert np.array_equal(h.view(np.uint8), h2.view(np.uint8))
    # nor on the size of the search cells (as fine as memory allows by default)
    for max_cells in ("300", "20000"):
        os.environ["MGP_FOF_MAX_CELLS"] = max_cells
        try:
            h3 = pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"],
                               cfg["mass_part"], cfg["dDdy"], cfg["dD2dy"])
        finally:
            del os.environ["MGP_FOF_MAX_CELLS"]
        assert np.array_equal(h.view(np.uint8), h3.view(np.uint8))
    pm.close()


@pytest.mark.gpu
def test_cuda_halo_finder_after_steps_and_edge_cases(mgp, require_gpu, case):
    """After sorted steps (the store is in cell order, not the upload order) the catalogue is the oracle's on the
    downloaded particles; np_min above every group gives none; a strip of half the box is refused (mm_main.c:214)."""
    from oracle import pm_oracle as po
    c, cfg = case, case["cfg"]
    pm = mgp.PM(N1D, N1D, BOX, grid_bytes=8, use_cola=1, sort_particles=1, deposit_mode=mgp.DEPOSIT_ATOMIC)
    pm.upload_particles(c["pos"], c["vel"], c["D"], c["D2"])
    pm.GetDisplacements()
    pm.Kick(0.8, 1e-4, 1.0, -0.4)
    pm.Drift(1e-3, 1e-3, -1e-3)
    got = pm.download_particles()
    h = pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"],
                      cfg["mass_part"], cfg["dDdy"], cfg["dD2dy"])
    orc, _ = po.fof_halos([dict(pos=got["pos"], vel=got["vel"], D=got["D"], D2=got["D2"], local_p_start=0)], cfg["norm_pos"],
                          cfg["norm_vel"], BOX * cfg["norm_pos"], cfg["dx_extra"], cfg["b_fof"], cfg["np_min"], cfg["mass_part"],
                          N1D, 1, cfg["dDdy"], cfg["dD2dy"])
    assert h.size >= 30
    fc.assert_same_halos(h, orc[0], exact_vectors=False)
    assert pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX, cfg["dx_extra"], cfg["b_fof"], 10 ** 6, cfg["mass_part"]).size == 0
    with pytest.raises(mgp.MgpError) as e:
        pm.MatchMaker(cfg["norm_pos"], cfg["norm_vel"], BOX, 0.5 * BOX, cfg["b_fof"], cfg["np_min"], cfg["mass_part"])
    assert e.value.code == -1
    pm.close()
