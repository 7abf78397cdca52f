"""Builds libmgpicola_cuda.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(OUT_DIR, "libmgpicola_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
        os.path.join(ROOT, "include", "mgpicola.h")]
    objs, jobs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OUT_DIR, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or not os.path.exists(LIB):
        # NCCL: link the copy PyTorch bundles when there is one (2.28.x), so that a process that loads both
        # this library and torch sees ONE libnccl.so.2; the system library (2.27.x) otherwise.
        nccl = ["-lnccl"]
        try:
            import importlib.util
            sp = importlib.util.find_spec("nvidia.nccl")
            d = os.path.join(list(sp.submodule_search_locations)[0], "lib")
            if os.path.exists(os.path.join(d, "libnccl.so.2")):
                nccl = ["-Xlinker", os.path.join(d, "libnccl.so.2"), "-Xlinker", "-rpath," + d]
        except Exception:
            pass
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + [
            "-lcufft"] + nccl + ["-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
