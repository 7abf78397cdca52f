"""Writes tests/golden/input_power_spectrum.npz from the reference's bundled CAMB table
(files/input_power_spectrum.dat, 691 rows of k [h/Mpc], P(k) [(Mpc/h)^3]).  Run in the build
container (the GPU box has no /root/reference); the fixture is what bench.py and the tests read."""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
tab = np.loadtxt(os.path.join(ref, "files", "input_power_spectrum.dat"))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_power_spectrum.npz")
np.savez_compressed(out, k=tab[:, 0], P=tab[:, 1])
print(out, tab.shape)
