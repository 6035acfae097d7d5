"""Deterministic synthetic RGB-D scenes: a smooth textured height field rendered by
ray casting from arbitrary camera poses, with TUM-style invalid depth.  Used by tests (CPU torch) and by
bench.py (rendered on the GPU, outside the timed region).  Plumbing only -- nothing here is on the hot path.

Conventions: camera pose (R_wc, t_wc) maps camera to scene coordinates, X_w = R_wc X_c + t_wc; the relative
pose of frame j with respect to frame i is T_ij = T_i^-1 T_j (X_i = R_ij X_j + t_ij), which is the tracker's
_{KF}T^{cur} (src/visodo.cpp:1458-1460).  Depth is written as uint16 millimetres (what VisodoTracker::depth_
receives after the application's 0.2 scaling of TUM's 5000-per-metre PNGs, tools/evaluation.cpp:285).
"""
import math

import numpy as np
import torch

FREIBURG1 = dict(fx=525.0, fy=525.0, cx=319.5, cy=239.5)  # config_data/calibration_freiburg1.ini


def intrinsics_for(rows, cols):
    s = cols / 640.0
    return dict(fx=525.0 * s, fy=525.0 * s, cx=(cols - 1) / 2.0, cy=(rows - 1) / 2.0)


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * (K @ K)


class Scene:
    def __init__(self, seed=20261017, amplitude=0.25, n_tex=24):
        rng = np.random.default_rng(seed)
        self.seed = seed
        self.amp = amplitude
        self.phi = rng.uniform(0, 2 * math.pi, 2)
        self.kx = 2 * math.pi / 2.4
        self.ky = 2 * math.pi / 2.4
        self.tilt = 0.3 / 2.4
        self.a = rng.uniform(2.0, 10.0, n_tex)
        freq = np.exp(rng.uniform(math.log(1.0), math.log(30.0), n_tex))  # cycles per metre
        ang = rng.uniform(0, 2 * math.pi, n_tex)
        self.fu = freq * np.cos(ang)
        self.fv = freq * np.sin(ang)
        self.ph = rng.uniform(0, 2 * math.pi, n_tex)

    # height field Z(X, Y) and its partial derivatives
    def surface(self, X, Y):
        s = torch.sin(self.kx * X + self.phi[0])
        c = torch.cos(self.ky * Y + self.phi[1])
        Z = 2.0 + self.amp * s * c + self.tilt * X
        Zx = self.amp * self.kx * torch.cos(self.kx * X + self.phi[0]) * c + self.tilt
        Zy = -self.amp * self.ky * s * torch.sin(self.ky * Y + self.phi[1])
        return Z, Zx, Zy

    def texture(self, X, Y):
        I = torch.full_like(X, 127.0)
        for a, fu, fv, ph in zip(self.a, self.fu, self.fv, self.ph):
            I = I + a * torch.sin(2 * math.pi * (fu * X + fv * Y) + ph)
        return I.clamp(0.0, 255.0)

    def render(self, R_wc, t_wc, rows=480, cols=640, intr=None, device="cpu", frame_id=0, invalid=True, noise=False):
        """Returns (depth_mm uint16 [rows, cols], rgb uint8 [rows, cols, 3], depth_m float64, intensity float64)."""
        intr = intr or intrinsics_for(rows, cols)
        dt = torch.float64
        R = torch.as_tensor(np.asarray(R_wc, dtype=np.float64), device=device)
        t = torch.as_tensor(np.asarray(t_wc, dtype=np.float64), device=device)
        v, u = torch.meshgrid(torch.arange(rows, dtype=dt, device=device), torch.arange(cols, dtype=dt, device=device),
                              indexing="ij")
        dx, dy = (u - intr["cx"]) / intr["fx"], (v - intr["cy"]) / intr["fy"]
        dwx = R[0, 0] * dx + R[0, 1] * dy + R[0, 2]
        dwy = R[1, 0] * dx + R[1, 1] * dy + R[1, 2]
        dwz = R[2, 0] * dx + R[2, 1] * dy + R[2, 2]
        lam = torch.full_like(dx, 2.0)
        for _ in range(12):  # Newton on g(lam) = Xw_z - Z(Xw_x, Xw_y)
            Xw, Yw, Zw = t[0] + lam * dwx, t[1] + lam * dwy, t[2] + lam * dwz
            Z, Zx, Zy = self.surface(Xw, Yw)
            g = Zw - Z
            gp = dwz - Zx * dwx - Zy * dwy
            lam = lam - g / gp
        Xw, Yw = t[0] + lam * dwx, t[1] + lam * dwy
        inten = self.texture(Xw, Yw)
        depth = lam.clone()
        rng = np.random.default_rng((self.seed * 1000003 + frame_id * 7919 + 17) % (2 ** 63))
        if noise:
            inten = inten + torch.as_tensor(rng.normal(0.0, 1.0, (rows, cols)), device=device)
            w = 1.0 / depth + torch.as_tensor(rng.normal(0.0, 0.001, (rows, cols)), device=device)
            depth = 1.0 / w
            inten = inten.clamp(0.0, 255.0)
        depth_mm = torch.round(depth * 1000.0).clamp(0, 65535)
        if invalid:
            nb_r, nb_c = rows // 8, cols // 8
            blocks = torch.as_tensor(rng.random((nb_r, nb_c)) < 0.05, device=device)
            mask = blocks.repeat_interleave(8, 0).repeat_interleave(8, 1)
            depth_mm = torch.where(mask, torch.zeros_like(depth_mm), depth_mm)
            depth_mm[:, :10] = 0
        depth_u16 = depth_mm.to(torch.int32).to(torch.uint16) if hasattr(torch, "uint16") else depth_mm.to(torch.int16)
        rgb = torch.stack([inten, inten * 0.95, inten * 0.9], dim=-1).round().clamp(0, 255).to(torch.uint8)
        return depth_u16, rgb, depth, inten


def trajectory(n_frames, seed=0, trans_step=0.012, rot_step_deg=0.6):
    """Smooth camera trajectory; per-frame increments stay below 0.02 m and 1.0 degree."""
    rng = np.random.default_rng(seed + 991)
    pa = rng.uniform(0, 2 * math.pi, 6)
    R, t = np.eye(3), np.zeros(3)
    poses = [(R.copy(), t.copy())]
    for k in range(1, n_frames):
        v = trans_step * np.array([math.sin(2 * math.pi * k / 60 + pa[0]), math.cos(2 * math.pi * k / 45 + pa[1]),
                                   0.5 * math.sin(2 * math.pi * k / 90 + pa[2])])
        w = math.radians(rot_step_deg) * np.array([math.sin(2 * math.pi * k / 50 + pa[3]),
                                                    math.cos(2 * math.pi * k / 70 + pa[4]),
                                                    0.7 * math.sin(2 * math.pi * k / 40 + pa[5])])
        dR = _so3_exp(w)
        t = t + R @ v
        R = R @ dR
        poses.append((R.copy(), t.copy()))
    return poses


def relative_pose(pose_i, pose_j):
    """T_ij = T_i^-1 T_j: X_i = R_ij X_j + t_ij."""
    Ri, ti = pose_i
    Rj, tj = pose_j
    return Ri.T @ Rj, Ri.T @ (tj - ti)


def random_pose(rng, max_trans=0.02, max_rot_deg=1.0):
    v = rng.normal(size=3)
    v = v / np.linalg.norm(v) * rng.uniform(0, max_trans)
    w = rng.normal(size=3)
    w = w / np.linalg.norm(w) * math.radians(rng.uniform(0, max_rot_deg))
    return _so3_exp(w), v


def make_pair(seed=20261017, rows=480, cols=640, device="cpu", max_trans=0.02, max_rot_deg=1.0, noise=False, invalid=True):
    """Frame A at the origin and frame B under a random small SE(3); returns dict with raw and float maps."""
    scene = Scene(seed)
    rng = np.random.default_rng(seed + 5)
    R_ab, t_ab = random_pose(rng, max_trans, max_rot_deg)
    dA, cA, _, _ = scene.render(np.eye(3), np.zeros(3), rows, cols, device=device, frame_id=0, noise=noise, invalid=invalid)
    dB, cB, _, _ = scene.render(R_ab, t_ab, rows, cols, device=device, frame_id=1, noise=noise, invalid=invalid)
    return dict(depth_a=dA, rgb_a=cA, depth_b=dB, rgb_b=cB, R_ab=R_ab, t_ab=t_ab, intr=intrinsics_for(rows, cols))


def make_sequence(seed, n_frames, rows=480, cols=640, device="cpu", noise=False):
    scene = Scene(seed)
    poses = trajectory(n_frames, seed)
    intr = intrinsics_for(rows, cols)
    depth, rgb = [], []
    for k, (R, t) in enumerate(poses):
        d, c, _, _ = scene.render(R, t, rows, cols, intr=intr, device=device, frame_id=k, noise=noise)
        depth.append(d)
        rgb.append(c)
    return dict(depth=torch.stack(depth), rgb=torch.stack(rgb), poses=poses, intr=intr)


def write_tum_sequence(seq, folder, fps=30.0):
    """Write a sequence in the TUM RGB-D layout the reference's evaluation reader expects
    (tools/evaluation.cpp:164-199): rgb/*.png, depth/*.png (5000 per metre), rgb.txt / depth.txt with three
    header lines, and a match file `t_d depth/... t_rgb rgb/...`."""
    import os
    import cv2
    os.makedirs(os.path.join(folder, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(folder, "depth"), exist_ok=True)
    hdr = "# synthetic sequence\n# file: '%s'\n# timestamp filename\n" % os.path.basename(folder)
    lr, ld, lm = [hdr], [hdr], []
    for k in range(seq["depth"].shape[0]):
        ts = "%.6f" % (1000.0 + k / fps)
        rgb = seq["rgb"][k].cpu().numpy()
        d_mm = seq["depth"][k].cpu().numpy().astype(np.uint32)
        cv2.imwrite(os.path.join(folder, "rgb", ts + ".png"), rgb[:, :, ::-1])
        cv2.imwrite(os.path.join(folder, "depth", ts + ".png"), (d_mm * 5).astype(np.uint16))
        lr.append("%s rgb/%s.png\n" % (ts, ts))
        ld.append("%s depth/%s.png\n" % (ts, ts))
        lm.append("%s depth/%s.png %s rgb/%s.png\n" % (ts, ts, ts, ts))
    open(os.path.join(folder, "rgb.txt"), "w").writelines(lr)
    open(os.path.join(folder, "depth.txt"), "w").writelines(ld)
    open(os.path.join(folder, "matches.txt"), "w").writelines(lm)
