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


def _c_layout(header, include_dir, pairs, tmp_path):
    """sizeof / offsetof of every mirrored struct as the C compiler sees them."""
    import subprocess
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "%s"' % header, "int main(void) {"]
    for cname, cls in pairs:
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in cls._fields_:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0]))
    src.append("return 0; }")
    c, exe = tmp_path / "layout.c", tmp_path / "layout"
    c.write_text("\n".join(src))
    r = subprocess.run(["gcc", "-I", include_dir, str(c), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr  # a field named in the ctypes mirror does not exist in the header
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    return {k: int(v) for k, v in (ln.split() for ln in out.splitlines())}


def _check_layout(got, pairs):
    for cname, cls in pairs:
        assert got[cname] == C.sizeof(cls), (cname, got[cname], C.sizeof(cls))
        for f in cls._fields_:
            assert got["%s.%s" % (cname, f[0])] == getattr(cls, f[0]).offset, (cname, f[0])


def test_ctypes_mirrors_have_the_layout_of_the_header(tmp_path):
    """include/rgbid_b200.h is plain C: every struct of the boundary, field by field, against rgbid-slam_b200/capi.py."""
    from rgbid_slam_b200 import capi
    pairs = [("rgbid_system_params", capi.SystemParams), ("rgbid_intr", capi.Intr), ("rgbid_depth_dist", capi.DepthDist),
             ("rgbid_custom_calibration", capi.CustomCalibration), ("rgbid_align_config", capi.AlignConfig),
             ("rgbid_iter_trace", capi.IterTrace), ("rgbid_tracker_config", capi.TrackerConfig),
             ("rgbid_frame_result", capi.FrameResult), ("rgbid_keyframe_handoff", capi.KeyframeHandoff)]
    _check_layout(_c_layout("rgbid_b200.h", os.path.join(ROOT, "include"), pairs, tmp_path), pairs)
    # a zeroed configuration is the reference's shipped one (pyrFirst), not enum value 0 of the reference (WARP_FIRST)
    assert capi.AlignConfig().warp_first == 0


def test_oracle_ctypes_mirrors_have_the_layout_of_its_header(tmp_path):
    import oracle as orc
    pairs = [("orc_system_params", orc.SystemParams), ("orc_align_config", orc.AlignConfig), ("orc_pyramids", orc.Pyramids),
             ("orc_iter_trace", orc.IterTrace), ("orc_frame_stats", orc.FrameStats)]
    _check_layout(_c_layout("oracle.h", os.path.join(ROOT, "oracle"), pairs, tmp_path), pairs)


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
