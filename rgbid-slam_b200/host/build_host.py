"""Builds the C++ drop-in test program (tests/cpp/test_dropin) against the header-only host layer
(include/rgbid_b200/*.hpp, rgbid-slam_b200/host/*.hpp) and librgbid_b200.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "test_dropin")
LIBDIR = os.path.join(ROOT, "rgbid-slam_b200", "lib")


def build():
    deps = [SRC, os.path.join(HERE, "visodo.hpp"), os.path.join(HERE, "settings.hpp"),
            os.path.join(ROOT, "include", "rgbid_b200", "internal.hpp"),
            os.path.join(ROOT, "include", "rgbid_b200", "device_array.hpp"), os.path.join(ROOT, "include", "rgbid_b200.h"),
            os.path.join(LIBDIR, "librgbid_b200.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        build_app()
        return OUT
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-I/usr/local/cuda/include", SRC, "-o", OUT, "-L" + LIBDIR, "-lrgbid_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building test_dropin failed:\n" + r.stderr)
    build_app()
    return OUT


def build_app():
    """apps/rgbid_slam_app: the RGBID_SLAMapp-compatible evaluation driver (TUM sequences in, pose log out)."""
    src, out = os.path.join(ROOT, "apps", "rgbid_slam_app.cpp"), os.path.join(ROOT, "apps", "rgbid_slam_app")
    deps = [src, os.path.join(HERE, "visodo.hpp"), os.path.join(HERE, "settings.hpp"), os.path.join(HERE, "tum_io.hpp"),
            os.path.join(HERE, "keyframe.hpp"), os.path.join(LIBDIR, "librgbid_b200.so")]
    deps = [d for d in deps if os.path.exists(d)]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-I/usr/local/cuda/include", src, "-o", out, "-L" + LIBDIR, "-lrgbid_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lz", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building rgbid_slam_app failed:\n" + r.stderr)
    return out


if __name__ == "__main__":
    print(build())
