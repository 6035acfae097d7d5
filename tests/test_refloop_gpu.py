"""The boundary, proven by the compiler: oracle/ref_shim.cpp -- the host loop that drives RGBID_SLAM::device::* in the
reference's own order (src/visodo.cpp:1041-1415, src/keyframe_align.cpp:178-350, trackNewFrame's image preparation,
covisibility and fusion calls) -- is compiled a second time, UNCHANGED, against include/rgbid_b200/internal.hpp
(through tests/cpp/refloop_include/internal.h) and linked with librgbid_b200.so (tests/cpp/librefloop_dropin.so, built by
tests/cpp/build_refloop.py).  The same translation unit built on the reference's own kernels is
oracle/_ref/libref_oracle.so; both must produce the same results."""
import contextlib
import ctypes as C
import os

import numpy as np
import pytest

from util import cuda, pair_maps, rot_angle, sums_rel_err
import oracle as orc
from oracle import ref as refk
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tests", "cpp", "librefloop_dropin.so")


@contextlib.contextmanager
def dropin_build():
    """oracle.ref bound to the drop-in build of the same shim instead of the reference build."""
    assert os.path.exists(DROPIN), "tests/cpp/librefloop_dropin.so was not built (see __graft_entry__.build)"
    real = refk._lib
    l = C.CDLL(DROPIN)
    for n in ("ref_pyr_down", "ref_gradient", "ref_bilateral", "ref_warp_invdepth", "ref_warp_intensity",
              "ref_warp_invdepth_weighted", "ref_integrate_warped_frame", "ref_visibility_ratio", "ref_build_system"):
        getattr(l, n).restype = C.c_float
    assert l.ref_init(0) == 0
    refk._lib = l
    try:
        yield
    finally:
        refk._lib = real


def _align(P, rows, cols, levels, mode, its):
    i = P["intr"]
    cfg = orc.make_config(rows, cols, levels, mode, its, i["fx"], i["fy"], i["cx"], i["cy"])
    return refk.align(cfg, refk.prepare_keyframe(cuda(P["WA"]), cuda(P["IA"]), levels, mode == orc.MODE_TRACKER),
                      refk.prepare_current(cuda(P["WB"]), cuda(P["IB"]), levels))


@pytest.mark.parametrize("mode,levels,its", [(orc.MODE_TRACKER, 3, [10, 5, 3]), (orc.MODE_ALIGN, 4, [5, 5, 3, 0])])
def test_reference_order_loop_on_the_dropin_header(built, mode, levels, its):
    rows, cols = 480, 640
    P = pair_maps(seed=7001 + mode, rows=rows, cols=cols, noise=True)
    with dropin_build():
        mine = _align(P, rows, cols, levels, mode, its)
    i = P["intr"]
    ocfg = orc.make_config(rows, cols, levels, mode, its, i["fx"], i["fy"], i["cx"], i["cy"])
    cpu = orc.align(ocfg, orc.prepare_keyframe(P["WA"], P["IA"], levels, mode == orc.MODE_TRACKER),
                    orc.prepare_current(P["WB"], P["IB"], levels))
    others = [("CPU oracle", cpu)]
    if refk.available():
        others.append(("reference build of the same TU", _align(P, rows, cols, levels, mode, its)))
    assert mine["status"] == 0
    for name, o in others:
        dt, ang = float(np.linalg.norm(mine["t"] - o["t"])), rot_angle(mine["R"], o["R"])
        e = sums_rel_err(mine["trace"][0]["sums27"], o["trace"][0]["sums27"])
        print("drop-in build vs %s: |dt| = %.2e m, angle = %.2e rad, first-iteration sums %.2e" % (name, dt, ang, e))
        assert dt < 1e-4 and ang < 1e-4 and e < 1e-5
        assert mine["trace"][0]["nu_depthinv"] == o["trace"][0]["nu_depthinv"]
        assert len(mine["trace"]) == len(o["trace"])


def test_reference_order_tracker_on_the_dropin_header(built):
    """trackNewFrame's whole call sequence (ingest, pyramids, keyframe preparation with the bilateral filter, alignment,
    covisibility, fusion, vertex / normal maps) through the reference-named functions of the drop-in header."""
    rows, cols, n = 240, 320, 10
    seq = synth.make_sequence(seed=7100, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    with dropin_build():
        ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="ref")
        mine = [ot.track(seq["depth"][k].cuda(), seq["rgb"][k].cuda()) for k in range(n)]
        fused = ot.intW.cpu().numpy()
    oc = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu")
    want = [oc.track(seq["depth"][k].numpy().astype(np.uint16), seq["rgb"][k].numpy()) for k in range(n)]
    worst = 0.0
    for k, (a, b) in enumerate(zip(mine, want)):
        assert a["status"] == b["status"] == 0
        assert a["new_odo_keyframe"] == b["new_odo_keyframe"] and a["new_integr_keyframe"] == b["new_integr_keyframe"], k
        dt, ang = float(np.linalg.norm(a["t"] - b["t"])), rot_angle(a["R"], b["R"])
        worst = max(worst, dt, ang)
        assert dt < 1e-4 and ang < 1e-4, (k, dt, ang)
    m = ~(np.isnan(fused) | np.isnan(oc.intW))
    assert np.mean(np.isnan(fused) == np.isnan(oc.intW)) > 0.999
    assert np.mean(np.abs(fused[m] - oc.intW[m]) / oc.intW[m] < 1e-4) > 0.999
    print("reference-order tracker on the drop-in header vs CPU oracle: worst pose difference %.2e" % worst)
