"""Launcher of the multi-process MPI stand-in (standins/shim_mpi_mp.c): starts `nranks` copies of a reference
executable built with it (oracle/_ref/MG_PICOLA_<variant>_mp), all mapping one zero-filled file under /dev/shm.

TEST INFRASTRUCTURE ONLY: the CPU baseline of bench.py (the reference's multi-rank path on the box's host cores) and a
multi-rank parity oracle.

    python -m oracle.mprun <nranks> <executable> <args...>
"""
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def exe_path(variant):
    return os.path.join(REF_DIR, "MG_PICOLA_%s_mp" % variant)


def available(variant):
    return os.path.exists(exe_path(variant))


def _die_with_parent():
    """preexec: SIGKILL this rank when the launcher dies (a rank waiting in a barrier would otherwise spin for ever)."""
    try:
        import ctypes
        ctypes.CDLL("libc.so.6").prctl(1, 9)          # PR_SET_PDEATHSIG, SIGKILL
    except Exception:
        pass


def run(cmd, nranks, slot_mb=64, scratch_mb=64, cwd=None, timeout=3600, env=None, rank_env=None, line_cb=None):
    """Runs `cmd` (list) on `nranks` ranks.  Returns (returncode, stdout of rank 0, tail of every rank's stderr).
    scratch_mb must hold one complex half-spectrum of the mesh: Nmesh^2 (Nmesh/2+1) * 16 bytes (double grids).
    line_cb(line, t_seconds_since_start): called for every line of rank 0's stdout as it arrives; returning True stops
    the run (all ranks are killed, returncode 0): bench.py times the reference's iterations this way."""
    import shutil
    # pages appear when touched: the scratch grid in full, of the message slots only what the messages use
    touched = (scratch_mb << 20) + nranks * (8 << 20)
    shm_dir = "/dev/shm"
    try:
        if not os.path.isdir(shm_dir) or shutil.disk_usage(shm_dir).free < 2 * touched:
            shm_dir = tempfile.gettempdir()          # a file-backed shared mapping works as well (page cache)
    except OSError:
        shm_dir = tempfile.gettempdir()
    fd, shm = tempfile.mkstemp(prefix="mgpshim_", dir=shm_dir)
    try:
        os.ftruncate(fd, 8192 + nranks * (slot_mb << 20) + (scratch_mb << 20) + 8192)     # sparse file
        os.close(fd)
        procs, errfiles = [], []
        for r in range(nranks):
            e = dict(os.environ, **(env or {}))
            if rank_env is not None:
                e.update(rank_env(r))                   # e.g. one GPU per rank: {"MGP_DEVICE": str(r)}
            e.update(MGPSHIM_RANK=str(r), MGPSHIM_SIZE=str(nranks), MGPSHIM_SHM=shm, MGPSHIM_SLOT_MB=str(slot_mb),
                     MGPSHIM_SCRATCH_MB=str(scratch_mb))
            out = subprocess.PIPE if r == 0 else subprocess.DEVNULL
            errf = tempfile.TemporaryFile(mode="w+")          # a file, not a pipe: nobody drains it while the ranks run
            errfiles.append(errf)
            procs.append(subprocess.Popen(cmd, cwd=cwd, env=e, stdout=out, stderr=errf, text=True, preexec_fn=_die_with_parent))
        t0 = time.time()
        # rank 0's stdout can be large: drain it in this thread while polling the others
        import threading
        chunks = []
        stop = []

        def drain():
            if line_cb is None:
                chunks.append(procs[0].stdout.read())
                return
            for line in procs[0].stdout:
                chunks.append(line)
                if not stop and line_cb(line, time.time() - t0):
                    stop.append(1)
        th = threading.Thread(target=drain, daemon=True)
        th.start()
        rc = None
        while True:
            if stop:
                rc = 0
                break
            codes = [p.poll() for p in procs]
            if any(c not in (None, 0) for c in codes):                  # a rank died or aborted: the others wait forever
                rc = next(c for c in codes if c not in (None, 0))
                break
            if all(c == 0 for c in codes):
                rc = 0
                break
            if time.time() - t0 > timeout:
                rc = -9
                break
            time.sleep(0.05)
        for p in procs:
            if p.poll() is None:
                p.kill()
        th.join(timeout=5)
        errs = []
        for p, f in zip(procs, errfiles):
            try:
                p.wait(timeout=5)
                f.seek(0)
                errs.append(f.read()[-2000:])
                f.close()
            except Exception:
                errs.append("")
        return rc, "".join(chunks), errs
    finally:
        try:
            os.unlink(shm)
        except OSError:
            pass


def scratch_mb_for(nmesh, grid_bytes=8):
    return int(nmesh * nmesh * (nmesh // 2 + 1) * 2 * grid_bytes / (1 << 20)) + 2


if __name__ == "__main__":
    n = int(sys.argv[1])
    rc, out, errs = run(sys.argv[2:], n, scratch_mb=int(os.environ.get("MGPSHIM_SCRATCH_MB", "1100")))
    sys.stdout.write(out)
    for r, e in enumerate(errs):
        if e.strip():
            sys.stderr.write("[rank %d stderr] %s\n" % (r, e))
    sys.exit(rc if rc >= 0 else 1)
