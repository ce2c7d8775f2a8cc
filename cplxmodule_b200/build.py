"""In-tree build of the native library ``cplxmodule_b200/csrc/libcplxk.so``.

``nvcc`` cross-compiles for sm_100a without a GPU, so this runs on a CPU-only
box; the resulting ``.so`` is git-ignored and ships to the GPU box with the
working tree.  Usage: ``python -m cplxmodule_b200.build [--force] [--verbose]``.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CPLXK_BUILD_TAG=<tag> [CPLXK_BUILD_FLAGS="-D..."]: a side build csrc/libcplxk_<tag>.so for same-box A/B
# runs and instrumented measurements (loaded with CPLXK_LIB=<path>); never the shipped library
_TAG = os.environ.get("CPLXK_BUILD_TAG", "")
OUT = os.path.join(CSRC, f"libcplxk_{_TAG}.so" if _TAG else "libcplxk.so")
OBJ = os.path.join(CSRC, f"_obj_{_TAG}" if _TAG else "_obj")
SOURCES = ["api.cu", "kl.cu", "fwd_simt.cu", "fwd_tc.cu", "fwd_tc3.cu", "fwd_lin3.cu", "conv.cu", "conv_tc.cu", "bwd.cu", "bilinear.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build cplxmodule_b200/csrc/libcplxk.so")


def _fingerprint():
    h = hashlib.sha256()
    h.update(os.environ.get("CPLXK_DEBUG_BUILD", "0").encode())
    h.update(os.environ.get("CPLXK_BUILD_FLAGS", "").encode())
    names = sorted(os.listdir(CSRC)) + ["../../include/cplxk.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    stamp = os.path.join(OBJ, "stamp")
    fp = _fingerprint()
    if not force and os.path.exists(OUT) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == fp:
                return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()
    flags = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                    "--expt-relaxed-constexpr", "-I", os.path.join(HERE, "..", "include")]
    if os.environ.get("CPLXK_DEBUG_BUILD") == "1":   # measurement aids (CPLXK_DBG); never shipped
        flags += ["-DCPLXK_DEBUG"]
    flags += os.environ.get("CPLXK_BUILD_FLAGS", "").split()
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as f:
        f.write(fp)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
