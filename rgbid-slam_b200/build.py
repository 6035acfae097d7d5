"""Builds librgbid_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The library has no torch / Python dependency: plain pointers and sizes in, status codes out
(include/rgbid_b200.h).  cudart is linked statically, so the .so travels to the GPU box as is.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# RGBID_BUILD_TAG=<tag> (with RGBID_EXTRA_NVCC_FLAGS) builds a kernel variant next to the product library:
# build_<tag>/ and lib/librgbid_b200_<tag>.so, loaded with RGBID_LIB=<that path> (capi.py).  Variants are built here
# and travel to the GPU box, so that no GPU time is spent compiling.
_TAG = os.environ.get("RGBID_BUILD_TAG", "")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(LIBDIR, "librgbid_b200%s.so" % ("_" + _TAG if _TAG else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["image_ops.cu", "warp_ops.cu", "calib_ops.cu", "scale_est.cu", "gn_system.cu", "api.cu", "aligner.cu", "tracker.cu"]

# numeric flags of the reference build (CMakeLists.txt:110) so that per-pixel float arithmetic is compiled
# the same way as the reference's own kernels
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--ftz=true", "--prec-div=false", "--prec-sqrt=false",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))] + [
        os.path.join(HERE, "..", "include", "rgbid_b200.h")]


def build(force=False, verbose=False):
    extra = os.environ.get("RGBID_EXTRA_NVCC_FLAGS", "").split()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in _deps())
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s + ".o")
        if force or _newer(src, obj) or newest_hdr > os.path.getmtime(obj):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(cc, jobs))
    objs = [os.path.join(OBJDIR, s + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
