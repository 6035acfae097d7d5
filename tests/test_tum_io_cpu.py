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


@pytest.mark.parametrize("damage", ["truncated_ihdr", "huge_size", "idat_before_ihdr", "cut_file"])
def test_malformed_png_is_a_failed_grab_not_a_crash(built, tmp_path, damage):
    """The PNG reader runs on user-supplied dataset folders: every chunk length is checked before it is used."""
    import struct
    import zlib
    assert os.path.exists(APP), "apps/rgbid_slam_app was not built (see __graft_entry__.build)"
    seq = synth.make_sequence(seed=6, n_frames=2, rows=48, cols=64, noise=False)
    folder = str(tmp_path / "fr_bad")
    synth.write_tum_sequence(seq, folder)
    path = os.path.join(folder, "depth", "1000.000000.png")
    raw = open(path, "rb").read()

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xffffffff)
    sig, rest = raw[:8], raw[8:]
    ihdr_body = rest[8:8 + 13]
    after_ihdr = rest[8 + 13 + 4:]
    if damage == "truncated_ihdr":      # IHDR of 8 bytes: d[8], d[9], d[12] would be read past the chunk
        bad = sig + chunk(b"IHDR", ihdr_body[:8]) + after_ihdr
    elif damage == "huge_size":         # 2^31 - 1 by 2^31 - 1 pixels
        bad = sig + chunk(b"IHDR", struct.pack(">II", 0x7fffffff, 0x7fffffff) + ihdr_body[8:]) + after_ihdr
    elif damage == "idat_before_ihdr":
        bad = sig + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + rest
    else:                               # file cut in the middle of a chunk
        bad = raw[:len(raw) // 2]
    open(path, "wb").write(bad)
    r = subprocess.run([APP, "-check_io", "-eval", folder + "/", "-match_file", "matches.txt"], capture_output=True, text=True,
                       timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout + r.stderr)
    lines = r.stdout.strip().splitlines()
    assert lines[1] == "frame 0 grab failed", lines[:3]
    assert lines[2].startswith("frame 1 ") and "grab failed" not in lines[2]
