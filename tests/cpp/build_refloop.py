"""Test infrastructure (not product): builds tests/cpp/librefloop_dropin.so -- the reference-order host loop of
oracle/ref_shim.cpp compiled UNCHANGED against the drop-in header include/rgbid_b200/internal.hpp and linked with
librgbid_b200.so.  Used by tests/test_refloop_gpu.py only; __graft_entry__.build() calls it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIBDIR = os.path.join(ROOT, "rgbid-slam_b200", "lib")


def build_refloop():
    src = os.path.join(ROOT, "oracle", "ref_shim.cpp")
    orc = os.path.join(ROOT, "oracle", "oracle.c")
    out = os.path.join(ROOT, "tests", "cpp", "librefloop_dropin.so")
    inc = os.path.join(ROOT, "tests", "cpp", "refloop_include")
    deps = [src, orc, os.path.join(ROOT, "oracle", "oracle.h"), os.path.join(inc, "internal.h"),
            os.path.join(ROOT, "include", "rgbid_b200", "internal.hpp"), os.path.join(ROOT, "include", "rgbid_b200", "device_array.hpp"),
            os.path.join(ROOT, "include", "rgbid_b200.h"), os.path.join(LIBDIR, "librgbid_b200.so")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    obj_c = out + ".oracle.o"
    r = subprocess.run([gcc, "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-c", orc, "-o", obj_c], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the refloop library failed:\n" + r.stderr)
    cmd = [gxx, "-O2", "-fPIC", "-shared", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "oracle"),
           "-I" + inc, "-I" + os.path.join(ROOT, "include"), src, obj_c, "-o", out, "-L" + LIBDIR, "-lrgbid_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    os.remove(obj_c)
    if r.returncode != 0:
        raise RuntimeError("the reference-order host loop does not compile against include/rgbid_b200/internal.hpp:\n" + r.stderr)
    return out


if __name__ == "__main__":
    print(build_refloop())
