"""the RGBID_SLAMapp-compatible driver (apps/rgbid_slam_app) on a synthetic TUM-format sequence --
PNG + association files in, "<stamp> tx ty tz qx qy qz qw" pose log out -- against the restated trackNewFrame on the
CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from util import rot_angle
from oracle.tracker import OracleTracker
from rgbid_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "apps", "rgbid_slam_app")


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_app_on_tum_sequence(built, tmp_path):
    assert os.path.exists(APP), "apps/rgbid_slam_app was not built (see __graft_entry__.build)"
    rows, cols, n = 240, 320, 6
    seq = synth.make_sequence(seed=11, n_frames=n, rows=rows, cols=cols, noise=True)
    intr = seq["intr"]
    folder = str(tmp_path / "rgbd_dataset_synth")
    synth.write_tum_sequence(seq, folder)
    calib = tmp_path / "calibration.ini"
    calib.write_text("[CALIBRATION]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n" % (intr["fx"], intr["fy"], intr["cx"], intr["cy"]))
    config = tmp_path / "visodo.ini"
    config.write_text("[VISODO]\nM_ESTIMATOR = Student\nSIGMA_ESTIMATOR = sigmaML\nWARP_ORDER = pyrFirst\nIMAGE_FILTERING = none\n")
    log = tmp_path / "poses.txt"
    r = subprocess.run([APP, "-eval", folder + "/", "-match_file", "matches.txt", "-calib", str(calib), "-config", str(config),
                        "-o", str(log)], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "lost 0" in r.stdout
    rowsf = [[float(v) for v in ln.split()] for ln in log.read_text().splitlines()]
    assert len(rowsf) == n
    depth = seq["depth"].numpy().astype(np.uint16)
    rgb = seq["rgb"].numpy()
    ot = OracleTracker(rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu")
    for k in range(n):
        o = ot.track(depth[k], rgb[k])
        stamp, t, q = rowsf[k][0], np.array(rowsf[k][1:4]), np.array(rowsf[k][4:8])
        assert abs(stamp - (1000.0 + k / 30.0)) < 1e-5
        # the log is written with 6 decimals (ios::fixed, default precision)
        assert np.linalg.norm(t - o["t"]) < 1e-4 and rot_angle(quat_to_R(q), o["R"]) < 1e-4

    # default log name: "<dataset>_poses.txt" in the working directory (tools/RGBID_SLAMapp.cpp:414-428)
    r = subprocess.run([APP, "-eval", folder + "/", "-match_file", "matches.txt", "-calib", str(calib), "-config", str(config),
                        "-n", "2"],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(tmp_path / "rgbd_dataset_synth_poses.txt")
    # the two time logs of tools/evaluation.cpp:353-420 next to it
    misc = (tmp_path / "rgbd_dataset_synth_misc.txt").read_text().splitlines()
    assert [ln.split(":")[0] for ln in misc] == ["Mean time per frame", "Std time per frame", "Max time per frame"]
    assert float(misc[2].split(":")[1]) >= float(misc[0].split(":")[1]) > 0.0
    kft = (tmp_path / "rgbd_dataset_synth_kf_times.txt").read_text().splitlines()
    assert kft[0].split() == ["ObtainKeyframe", "ProcessKeyframeTotal", "Segmentation", "DescriptionBoW", "LoopDetection", "PoseGraphOptim"]
    assert all(len(ln.split()) == 6 for ln in kft[1:])

    # the tracker thread (VisodoTracker::start / operator(), the default) and -inline must write the same poses
    log2 = tmp_path / "poses_inline.txt"
    r = subprocess.run([APP, "-eval", folder + "/", "-match_file", "matches.txt", "-calib", str(calib), "-config", str(config),
                        "-o", str(log2), "-inline"], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert log2.read_text() == log.read_text()


def test_app_with_custom_calibration_file(built, tmp_path):
    """config_data/calibration_custom.ini dialect ([RGB_CALIBRATION], [DEPTH_CALIBRATION] custom_registration=1,
    [STEREO_DEPTH2RGB] with a multi-line dRc): the driver must take the custom-calibration ingest path and keep tracking."""
    rows, cols, n = 240, 320, 5
    seq = synth.make_sequence(seed=12, n_frames=n, rows=rows, cols=cols, noise=True)
    folder = str(tmp_path / "rgbd_dataset_custom")
    synth.write_tum_sequence(seq, folder)
    i = seq["intr"]
    calib = tmp_path / "calibration_custom.ini"
    calib.write_text(
        "[RGB_CALIBRATION]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\nkd=-0.01630   0.0   0.0   0.0  0.0\n\n"
        "[DEPTH_CALIBRATION]\ncustom_registration=1\nfx=%r\nfy=%r\ncx=%r\ncy=%r\nkd=-0.02711   0.0   0.0   0.0  0.0\n"
        "c1 = 0.98954\nc0 = -1.2618e-03\n"
        "q0 =  7.0023e-03   1.0844e-02  -6.0580e-01   1.2602e+00  -2.3050e-03   1.6084e-02   2.1441e-02  -1.8073e-02  -3.6722e-02\n"
        "q1 =  -6.7052e-03  -1.9692e-03   5.5808e-01  -1.2327e+00   1.2714e-02  -2.0804e-02  -7.5163e-03   3.1985e-02   4.8632e-02\n\n"
        "[STEREO_DEPTH2RGB]\ndRc=\n0.9999    0.0143    0.0060\n-0.0143    0.9999   -0.0018\n-0.0060    0.0017    1.0000\n"
        "t_dc=0.0263595 -0.0000973  0.0002853\n"
        % (i["fx"], i["fy"], i["cx"], i["cy"], i["fx"] * 1.06, i["fy"] * 1.06, i["cx"] + 1.4, i["cy"] + 1.2))
    # the [VISODO] section of the reference's shipped config_data/visodoRGBDconfig.ini (without -config the tracker runs
    # with the code default WARP_ORDER = warpFirst, like the reference's app)
    config = tmp_path / "visodo.ini"
    config.write_text("[VISODO]\nM_ESTIMATOR = Student\nSIGMA_ESTIMATOR = sigmaML\nWARP_ORDER = pyrFirst\nIMAGE_FILTERING = none\n")
    log = tmp_path / "poses.txt"
    r = subprocess.run([APP, "-eval", folder + "/", "-match_file", "matches.txt", "-calib", str(calib), "-config", str(config),
                        "-o", str(log)],
                       capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "frames %d" % n in r.stdout and "lost 0" in r.stdout
    plain = tmp_path / "plain.ini"
    plain.write_text("[CALIBRATION]\nfx=%r\nfy=%r\ncx=%r\ncy=%r\n" % (i["fx"], i["fy"], i["cx"], i["cy"]))
    log2 = tmp_path / "poses_plain.txt"
    r2 = subprocess.run([APP, "-eval", folder + "/", "-match_file", "matches.txt", "-calib", str(plain), "-config", str(config),
                         "-o", str(log2)],
                        capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r2.returncode == 0
    a = np.array([[float(v) for v in ln.split()] for ln in log.read_text().splitlines()])
    b = np.array([[float(v) for v in ln.split()] for ln in log2.read_text().splitlines()])
    assert a.shape == b.shape == (n, 8)
    assert np.abs(a[1:, 1:4] - b[1:, 1:4]).max() > 1e-6   # a different camera model was really used ...
    assert np.abs(a[:, 1:4] - b[:, 1:4]).max() < 0.05     # ... on the same motion
