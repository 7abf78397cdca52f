"""Developer tool (GPU, a few seconds, no torch import): the last two library changes of round 2 against their checkers --
the lightcone counting / drift passes against the reference's rows (tests/lightcone_case.py) and mgp_simple_pofk with
Npart != Ngrid^3 against the restatement pinned to the compiled tool."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()
import mgpicola_b200 as mgp
import lightcone_case as lcc
from oracle import pm_oracle as po

case = lcc.reference_case(tempfile.mkdtemp(prefix="qc"))
i, s = case["inputs"], case["scalars"]
N = i["nmesh"]
pm = mgp.PM(N, N, i["box"], omega=case["omega"], grid_bytes=8, use_cola=i["use_cola"], sort_particles=0)
pm.upload_particles(i["pos"], i["vel"], i["D"], i["D2"], i["ids"])
want = np.array([r.shape[0] for r in case["ref_rows"]], np.uint64)
cnt = pm.lightcone_count(s, case["reps"], i["sumxyz"])
print("lightcone counts equal:", bool(np.array_equal(cnt, want)), int(cnt.sum()))
rows = pm.Drift_Lightcone(s, case["reps"], i["sumxyz"], pinned=True)
ok = all(g.shape == r.shape and np.array_equal(lcc.sort_rows(g).view(np.uint32), lcc.sort_rows(r).view(np.uint32))
         for g, r in zip(rows, case["ref_rows"]))
after = pm.download_particles(("pos",))["pos"]
print("lightcone rows bit-identical:", ok, "positions:", bool(np.array_equal(after.view(np.uint32), case["ref_pos"].view(np.uint32))))
pm.close()

from test_gpu_parity import OMEGA, make_particles
N, box = 32, 100.0
pos, vel, D, D2 = make_particles(N, box, 13, clustered=True)
keep = N ** 3 - 9000
pm = mgp.PM(N, N, box, omega=OMEGA, grid_bytes=8, sort_particles=0)
pm.upload_particles(pos[:keep], vel[:keep], D[:keep], D2[:keep])
for scheme in ("NGP", "CIC", "TSC"):
    p, n = pm.simple_pofk(scheme, subtract_shotnoise=True)
    pr, nr = po.simple_pofk(pos[:keep], N, box, scheme, subtract_shotnoise=True)
    good = nr > 0
    print("simple_pofk", scheme, "modes equal:", bool(np.array_equal(n, nr)), "max rel err: %.2e" % (np.abs(p[good] - pr[good]).max() / np.abs(pr[good]).max()))
pm.close()
print("seconds: %.1f" % (time.time() - t0))
