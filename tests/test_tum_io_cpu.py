"""SURVEY section 8 f1, host side: the TUM sequence reader of the drop-in driver (rgbid-slam_b200/host/tum_io.hpp, built into
apps/rgbid_slam_app) against OpenCV's decoder and numpy -- PNG decoding (8-bit RGB, 16-bit grey), the 0.2 depth
scaling with round-half-to-even, both association formats with the reference's trailing-entry quirk, and the
Eigen-style quaternion of the pose log."""
import os
import subprocess

import numpy as np
import pytest

from rgbid_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "apps", "rgbid_slam_app")


def checksum(a):
    flat = a.reshape(-1).astype(np.uint64)
    return int((flat * (np.arange(flat.size, dtype=np.uint64) % 251 + 1)).sum())


@pytest.mark.parametrize("use_match_file", [True, False])
def test_sequence_reader_matches_opencv(built, tmp_path, use_match_file):
    import cv2
    assert os.path.exists(APP), "apps/rgbid_slam_app was not built (see __graft_entry__.build)"
    seq = synth.make_sequence(seed=5, n_frames=3, rows=48, cols=64, noise=True)
    folder = str(tmp_path / "fr_synth")
    synth.write_tum_sequence(seq, folder)
    # odd depth values so that v * 0.2 hits exact halves (2.5 -> 2, 7.5 -> 8: half to even)
    d0 = cv2.imread(os.path.join(folder, "depth", "1000.000000.png"), cv2.IMREAD_UNCHANGED)
    d0[0, :8] = np.array([12, 13, 37, 38, 62, 63, 65535, 0], dtype=np.uint16)
    cv2.imwrite(os.path.join(folder, "depth", "1000.000000.png"), d0)
    if not use_match_file:  # folder mode reads *_associated.txt (three header lines, like rgb.txt / depth.txt)
        os.rename(os.path.join(folder, "rgb.txt"), os.path.join(folder, "rgb_associated.txt"))
        os.rename(os.path.join(folder, "depth.txt"), os.path.join(folder, "depth_associated.txt"))
    cmd = [APP, "-check_io", "-eval", folder + "/"] + (["-match_file", "matches.txt"] if use_match_file else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    # the files end with a newline: both readers append one empty association (tools/evaluation.cpp:172-181, :195-199)
    assert lines[0] == "associations 4"
    assert lines[4] == "frame 3 grab failed"
    for k in range(3):
        tok = lines[1 + k].split()
        ts = 1000.0 + k / 30.0
        assert abs(float(tok[2]) - ts) < 1e-5 and abs(float(tok[3]) - ts) < 1e-5 and tok[4:6] == ["48", "64"]
        name = "%.6f.png" % ts
        d = cv2.imread(os.path.join(folder, "depth", name), cv2.IMREAD_UNCHANGED)
        want_d = np.clip(np.rint(d.astype(np.float64) * 0.2), 0, 65535).astype(np.uint16)  # convertTo(.., 0.2)
        c = cv2.imread(os.path.join(folder, "rgb", name))[:, :, ::-1]
        assert int(tok[6]) == checksum(want_d) and int(tok[7]) == checksum(np.ascontiguousarray(c))
    if k == 2:
        assert want_d.max() > 0
    # Eigen::Quaternionf(R) for R = that fixed rotation: compare with the closed form
    q = np.array([float(v) for v in lines[5].split()[1:]])
    R = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    w = np.sqrt(1 + np.trace(R)) / 2
    want = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    assert np.allclose(q, want, atol=1e-5)
