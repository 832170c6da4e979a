"""
Build recipe for libgpsig_b200.so (the C-ABI library declared in include/gpsig_b200.h).

Compiles every gpsig_b200/csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU) and links them into
gpsig_b200/lib/libgpsig_b200.so.  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libgpsig_b200.so")
BUILDDIR = os.path.join(HERE, "build")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())  # not the absolute path: the GPU box runs from another directory
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile and link; returns the path of the shared library.  Skips work when sources are unchanged."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "gpsig_b200.h")]
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BUILDDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libgpsig_b200.stamp")
    digest = _digest(deps)
    if not force and os.path.exists(LIBPATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIBPATH
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILDDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIBPATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return LIBPATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
