"""Host-side mirror of the reference's slab bookkeeping (2LPT.c:47-176, auxPM.c:151-153).

The CUDA library computes the same numbers in mgp_create(); this module exists so that the Python
host side (bench.py, run.py, the multi-rank tests) can split particle sets and check ownership
without a GPU.  Integer logic only -- results are exact, not approximate."""
import numpy as np


def block(nmesh, nranks):
    """FFTW-MPI's default 1-D block: ceil(Nmesh / NTask) planes per task (fftw_mpi_local_size_3d, 2LPT.c:50)."""
    return (nmesh + nranks - 1) // nranks


def layout(nmesh, nsample, nranks, rank):
    """(Local_nx, Local_x_start, Local_np, Local_p_start) of `rank`."""
    b = block(nmesh, nranks)
    x0 = rank * b
    nx = 0 if x0 >= nmesh else min(b, nmesh - x0)
    npl, p0 = 0, nsample
    for i in range(nsample):
        slab = int(float(i * nmesh) / float(nsample))          # initialize_parts, 2LPT.c:124
        if x0 <= slab < x0 + nx:
            npl += 1
            p0 = min(p0, i)
    return nx, x0, npl, p0


def slab_to_task(nmesh, nranks):
    b = block(nmesh, nranks)
    return np.minimum(np.arange(nmesh) // b, nranks - 1).astype(np.int32)


def owner_of(pos_x, nmesh, box, nranks):
    """Slab_to_task[(int)(Pos[0] * Nmesh / Box)] for float32 positions (auxPM.c:151-153)."""
    X = (np.asarray(pos_x, dtype=np.float32).astype(np.float64) * (np.float64(nmesh) / np.float64(box))).astype(np.int64)
    X = np.clip(X, 0, nmesh - 1)
    return slab_to_task(nmesh, nranks)[X]


def lagrangian_owner(nsample, nmesh, nranks):
    """Task that creates Lagrangian plane i (Part_to_task, 2LPT.c:160-172)."""
    i = np.arange(nsample)
    slab = (i * nmesh / float(nsample)).astype(np.int64)
    return slab_to_task(nmesh, nranks)[slab]
