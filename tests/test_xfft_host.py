"""CPU check of the fused x-transform + slab-exchange kernels (csrc/xfft.cuh, csrc/xfft_mixed.cuh): the programs under
tests/host/ run the phase functions the kernels consist of, thread by thread with a barrier between phases, for several
emulated ranks and compare with a direct DFT in long double (both directions, double and float, every Nmesh the library
instantiates, the tile widths it picks and a few others)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prog", ["xfft_emul", "xfft_mixed_emul"])
def test_xfft_phases_match_direct_dft(tmp_path, prog):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / prog)
    subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                    os.path.join(ROOT, "mg-picola-public_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tests", "host", prog + ".cu")],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]
