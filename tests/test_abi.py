"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/mgpicola.h declares, the ctypes structs match the header, and compute entry points
fail loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(mgp):
    L = mgp.load_library()
    names = mgp.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libmgpicola_cuda.so does not export %s" % n
    out = subprocess.run(["nm", "-D", "--defined-only", mgp.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mgp_[a-z0-9_]+)", out))
    assert set(names) <= exported
    # nothing but the mgp_* surface is exported with C linkage from our own objects
    assert L.mgp_version() >= 100


def test_header_compiles_as_c_and_struct_sizes_match(mgp, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mgpicola.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mgp_config), '
                   'sizeof(mgp_pofk_config), sizeof(mgp_step_scalars), sizeof(mgp_lightcone_step), sizeof(mgp_fof_config), '
                   'sizeof(mgp_fof_halo)); return 0;}\n')
    exe = tmp_path / "t"
    inc = os.path.join(os.path.dirname(mgp.HEADER_PATH))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(mgp.Config), C.sizeof(mgp.PofkConfig), C.sizeof(mgp.StepScalars), C.sizeof(mgp.LightconeStep),
                     C.sizeof(mgp.FofConfig), mgp.FOF_HALO_DTYPE.itemsize]


def test_no_cpu_fallback(mgp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour cannot be exercised")
    with pytest.raises(mgp.MgpError) as e:
        mgp.PM(16, 16, 10.0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "mg-picola-public_b200")
    for dp, dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/", "oracle/").lower() or f == "__init__.py" and "oracle" not in txt, \
                    "%s mentions the oracle: the product path must not depend on test infrastructure" % f


def test_cosmology_host_scalars_match_reference():
    """Host scalar set-up used by bench.py (growth factors, Sq, Sphi) against cosmo.c run here."""
    from oracle import ref_lib
    if not ref_lib.available("lcdm"):
        pytest.skip("oracle/_ref not built")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import tempfile
    from mgpicola_b200 import cosmology
    pf = bench.write_paramfile(tempfile.mkdtemp(prefix="mgp_cos_"), 16, 50.0, "fofr", 10)
    r = ref_lib.RefLib("lcdm")
    with ref_lib._silenced(True):
        r.init_from_paramfile(pf)
    L = r.lib
    cos = cosmology.LCDM(bench.OMEGA, bench.Z_INIT)
    for a in (0.1, 0.25, 0.5, 0.77, 1.0):
        for nm in ("growth_D", "growth_D2", "growth_dDdy", "growth_dD2dy", "growth_ddDddy", "growth_ddD2ddy"):
            assert abs(getattr(cos, nm)(a) / getattr(L, nm)(a) - 1) < 2e-6, nm      # reference ODE eps = 1e-7 (cosmo.c:38)
        assert abs(cos.Sq(a, a * 1.1, a * 1.05) / L.Sq(a, a * 1.1, a * 1.05) - 1) < 1e-6
        assert abs(cos.Sphi(a, a * 1.05, a) / L.Sphi(a, a * 1.05, a) - 1) < 1e-12
        s = cosmology.fofr_step_scalars(a, bench.OMEGA, 50.0, 1e-5, 1.0)
        assert abs(s["coupling"] - L.coupling_function(a)) < 1e-15


def test_schedule_matches_reference_loop():
    from mgpicola_b200 import cosmology
    seq = cosmology.schedule(9.0, [(0.0, 10)])
    assert len(seq) == 11                                  # N steps -> N + 1 force evaluations (main.c:397-399)
    assert seq[-1]["final"] and seq[-1]["output"]
    assert abs(seq[0]["AF"] - (0.1 + 0.045)) < 1e-15 and abs(seq[0]["AFF"] - 0.19) < 1e-15
    assert abs(seq[1]["AI"] - seq[0]["AF"]) < 1e-15
    seq2 = cosmology.schedule(9.0, [(1.0, 4), (0.0, 4)])
    out = [s for s in seq2 if s["output"]]
    assert len(out) == 2 and abs(out[0]["A"] - 0.5) < 1e-12 and abs(out[0]["AF"] - out[0]["A"]) < 1e-15
