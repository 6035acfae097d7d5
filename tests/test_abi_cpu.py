"""The C-ABI library builds, loads and exports every symbol include/rgbid_b200.h declares (no compute calls:
there is no GPU here), and the host-only entry points behave like the reference's host code."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rgbid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rgbid_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    assert len(syms) >= 45
    for must in ("rgbid_build_system", "rgbid_warp_invdepth", "rgbid_warp_intensity", "rgbid_sigma_nu_student",
                 "rgbid_aligner_run", "rgbid_tracker_track", "rgbid_integrate_warped_frame", "rgbid_visibility_ratio"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    from rgbid_slam_b200 import capi
    lib = capi.load()
    for s in declared_symbols():
        assert hasattr(lib, s), "librgbid_b200.so does not export " + s
        assert s in capi.PROTOTYPES, "capi.py does not bind " + s
    assert lib.rgbid_version() == 100
    assert lib.rgbid_status_string(-1).decode().startswith("numerical failure")
    assert b"argument" in lib.rgbid_status_string(-2)


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "rgbid_b200.h")).read()
    assert src.count("src/internal.h:") >= 20 and "src/visodo.cpp:" in src and "src/keyframe_align.cpp:" in src


@pytest.mark.parametrize("rows,cols,ns,want", [(480, 640, 10000, (120, 160, 4)), (240, 320, 10000, (120, 160, 2)),
                                               (120, 160, 10000, (120, 160, 1)), (60, 80, 10000, (60, 80, 1)),
                                               (480, 640, 19200, (120, 160, 4)), (960, 1280, 10000, (120, 160, 8)),
                                               (480, 640, 9999999, (480, 640, 1)), (15, 21, 10, (15, 21, 1))])
def test_error_geometry_matches_reference_rule(built, rows, cols, ns, want):
    """computeErrorGridStride, src/cuda/sigmaFuncs.cu:711-747."""
    from rgbid_slam_b200 import capi
    import oracle as orc
    lib = capi.load()
    kr, kc, s = C.c_int(), C.c_int(), C.c_int()
    assert lib.rgbid_error_geometry(rows, cols, ns, C.byref(kr), C.byref(kc), C.byref(s)) == 0
    assert (kr.value, kc.value, s.value) == want == orc.error_geometry(rows, cols, ns)


def test_bad_arguments_are_reported_not_fatal(built):
    from rgbid_slam_b200 import capi
    lib = capi.load()
    assert lib.rgbid_ctx_sync(None) == capi.ERR_ARG
    assert lib.rgbid_error_geometry(0, 10, 5, None, None, None) == capi.ERR_ARG
    assert lib.rgbid_aligner_create(None, None, None) == capi.ERR_ARG
    assert lib.rgbid_ctx_launch_count(None) == 0


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under rgbid-slam_b200/ or include/ may reference it."""
    bad = []
    for base in ("rgbid-slam_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dp or os.sep + "lib" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"\bimport oracle\b|from oracle\b|oracle\.h|liboracle|libref_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_missing_library_fails_loudly(built, monkeypatch):
    from rgbid_slam_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/librgbid_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        capi.load()
