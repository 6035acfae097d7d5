"""CPU restatement (numpy, float32 operation by operation) of the steps either side of the hot path that SURVEY section 8 (f)
ranks next: custom-calibration ingest (f3) and colour fusion / shaded previews (f4).
TEST INFRASTRUCTURE ONLY -- the checker for rgbid_undistort_* / rgbid_register_depthinv / rgbid_integrate_warped_rgb /
rgbid_generate_image; pinned against the reference's own kernels (oracle/ref.py, tests/golden/ref_golden_f34.npz).

Follows src/cuda/undistortion.cu:94-310, src/cuda/warping_registration.cu:148-281, 597-635, 672-712, 720-800 and
src/cuda/image_generator.cu:60-181.  Products and sums are kept un-fused in float32; the compiled kernels contract some
of them into FMAs, so agreement is to a few ulp (and to one texel where a gather position sits on a texel boundary)."""
import numpy as np

F = np.float32
NAN = F(np.nan)


def _f(x):
    return np.asarray(x, dtype=np.float32)


def distort_pixel(uu, vu, intr):
    """distortPixel, undistortion.cu:94-111.  intr: dict fx fy cx cy k1..k5"""
    k1, k2, k3, k4, k5 = (F(intr.get(k, 0.0)) for k in ("k1", "k2", "k3", "k4", "k5"))
    r2 = uu * uu + vu * vu
    r4 = r2 * r2
    r6 = r2 * r4
    factor_r = F(1) + k1 * r2 + k2 * r4 + k5 * r6
    ud = factor_r * uu
    ud = ud + (F(2) * k3 * uu * vu + k4 * (r2 + F(2) * uu * uu))
    vd = factor_r * vu
    vd = vd + (F(2) * k4 * uu * vu + k3 * (r2 + F(2) * vu * vu))
    return ud, vd


def _undistort_source(rows, cols, intr):
    fx, fy, cx, cy = (F(intr[k]) for k in ("fx", "fy", "cx", "cy"))
    xu, yu = np.meshgrid(np.arange(cols, dtype=np.float32), np.arange(rows, dtype=np.float32))
    uu = (xu - cx) * (F(1) / fx)
    vu = (yu - cy) * (F(1) / fy)
    ud, vd = distort_pixel(uu, vu, intr)
    xd = fx * ud + cx + F(0.5)
    yd = fy * vd + cy + F(0.5)
    inside = ~((xd <= 0) | (yd <= 0) | (xd >= cols) | (yd >= rows))
    return xd, yd, inside


def sample_bilinear_q8(img, xt, yt):
    """cudaFilterModeLinear with clamp addressing and the texture unit's 8-bit weights as measured on B200
    (rgbid-slam_b200/csrc/common.cuh sample_bilinear_q8)."""
    rows, cols = img.shape
    xB, yB = xt - F(0.5), yt - F(0.5)
    fxf, fyf = np.floor(xB), np.floor(yB)
    ka = np.floor((xB - fxf) * F(256) + F(0.5)).astype(np.int64)
    kb = np.floor((yB - fyf) * F(256) + F(0.5)).astype(np.int64)
    w11 = (ka * kb + 128) >> 8
    w10, w01 = ka - w11, kb - w11
    w00 = 256 - ka - kb + w11
    i0, j0 = fxf.astype(np.int64), fyf.astype(np.int64)
    i1, j1 = np.clip(i0 + 1, 0, cols - 1), np.clip(j0 + 1, 0, rows - 1)
    i0, j0 = np.clip(i0, 0, cols - 1), np.clip(j0, 0, rows - 1)
    acc = np.zeros(xt.shape, dtype=np.float32)
    for w, jj, ii in ((w00, j0, i0), (w10, j0, i1), (w01, j1, i0), (w11, j1, i1)):
        t = img[jj, ii]
        acc = np.where(w != 0, acc + w.astype(np.float32) * t, acc)  # a zero-weight tap is not blended
    return acc * F(1.0 / 256.0)


def undistort_intensity(src, intr):
    """undistortIntensity, undistortion.cu:143-170, 212-257"""
    src = _f(src)
    rows, cols = src.shape
    xd, yd, inside = _undistort_source(rows, cols, intr)
    with np.errstate(invalid="ignore"):
        val = sample_bilinear_q8(src, np.where(inside, xd, F(1)), np.where(inside, yd, F(1)))
    return np.where(inside, val, NAN).astype(np.float32)


def correct_depthinv(u, v, wm, dp):
    """correctDepthinv, undistortion.cu:113-137.  dp: dict c1 c0 q0[9] q1[9] xshift yshift"""
    q0, q1 = _f(dp["q0"]), _f(dp["q1"])
    wd = F(dp["c1"]) * wm + F(dp["c0"])
    r2 = u * u + v * v
    r4 = r2 * r2
    r6 = r2 * r4
    uv = u * v
    u2v = u * u * v
    uv2 = u * v * v
    D0 = q0[0] + q0[1] * r2 + q0[2] * r4 + q0[3] * r6 + q0[4] * u + q0[5] * v + q0[6] * uv + q0[7] * u2v + q0[8] * uv2
    D1 = q1[0] + q1[1] * r2 + q1[2] * r4 + q1[3] * r6 + q1[4] * u + q1[5] * v + q1[6] * uv + q1[7] * u2v + q1[8] * uv2
    return (F(1) + D1) * wd + D0


def undistort_depthinv(src, intr, dp):
    """undistortDepthInv, undistortion.cu:173-206 (correction) + :143-170 on a point-filtered texture (:260-310)"""
    src = _f(src)
    rows, cols = src.shape
    fx, fy, cx, cy = (F(intr[k]) for k in ("fx", "fy", "cx", "cy"))
    x, y = np.meshgrid(np.arange(cols), np.arange(rows))
    xs, ys = x - int(dp["xshift"]), y - int(dp["yshift"])
    ok = (xs > 0) & (ys > 0)
    u = (x.astype(np.float32) - cx) * (F(1) / fx)
    v = (y.astype(np.float32) - cy) * (F(1) / fy)
    val = src[np.clip(ys, 0, rows - 1), np.clip(xs, 0, cols - 1)]
    with np.errstate(invalid="ignore"):
        corr = np.where(ok, correct_depthinv(u, v, val, dp), NAN).astype(np.float32)
    xd, yd, inside = _undistort_source(rows, cols, intr)
    ix = np.clip(np.floor(xd), 0, cols - 1).astype(np.int64)
    iy = np.clip(np.floor(yd), 0, rows - 1).astype(np.int64)
    return np.where(inside, corr[iy, ix], NAN).astype(np.float32)


def register_depthinv(src, dRc_proj, t_dc_proj, cRd_proj):
    """registerDepthinv, warping_registration.cu:720-800: z-buffer splat with dilation on a 3 rows x 3 cols canvas
    (:148-165, 232-281), then the homography gather (:597-635)."""
    src = _f(src)
    rows, cols = src.shape
    crows, ccols = 3 * rows, 3 * cols
    ox, oy = (ccols - cols) // 2, (crows - rows) // 2
    t = _f(t_dc_proj)
    canvas = np.zeros((crows, ccols), dtype=np.int32)
    yd, xd = np.nonzero(~np.isnan(src))
    wd = src[yd, xd]
    zd = F(1) / wd
    Xx = xd.astype(np.float32) * zd - t[0]
    Xy = yd.astype(np.float32) * zd - t[1]
    Xz = zd - t[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        w_inter = F(1) / Xz
        xc, yc = Xx * w_inter, Xy * w_inter
        keep = w_inter > F(0.01)
        dil = w_inter / wd
        half = F(0.5) * dil
        xmin = np.rint(xc - half).astype(np.int64) + ox
        xmax = np.rint(xc + half).astype(np.int64) + ox
        ymin = np.rint(yc - half).astype(np.int64) + oy
        ymax = np.rint(yc + half).astype(np.int64) + oy
    bits = w_inter.view(np.int32)
    for k in np.nonzero(keep)[0]:
        x0, x1 = max(0, xmin[k]), min(xmax[k] + 1, ccols)
        y0, y1 = max(0, ymin[k]), min(ymax[k] + 1, crows)
        if x0 < x1 and y0 < y1:
            np.maximum(canvas[y0:y1, x0:x1], bits[k], out=canvas[y0:y1, x0:x1])
    inter = np.where(canvas != 0, canvas.view(np.float32), NAN)
    H, Hi = _f(dRc_proj).reshape(3, 3), _f(cRd_proj).reshape(3, 3)
    x, y = np.meshgrid(np.arange(cols, dtype=np.float32), np.arange(rows, dtype=np.float32))
    sx = H[0, 0] * x + H[0, 1] * y + H[0, 2] * F(1)
    sy = H[1, 0] * x + H[1, 1] * y + H[1, 2] * F(1)
    sz = H[2, 0] * x + H[2, 1] * y + H[2, 2] * F(1)
    inv = F(1) / sz
    sx, sy, sz = sx * inv, sy * inv, sz * inv
    x_src, y_src = sx + F(0.5) + F(ox), sy + F(0.5) + F(oy)
    ix, iy = np.floor(x_src).astype(np.int64), np.floor(y_src).astype(np.int64)
    inside = ~((ix < 0) | (iy < 0) | (ix >= ccols) | (iy >= crows))
    w_src = inter[np.clip(iy, 0, crows - 1), np.clip(ix, 0, ccols - 1)]
    dz = Hi[2, 0] * sx + Hi[2, 1] * sy + Hi[2, 2] * sz
    with np.errstate(divide="ignore", invalid="ignore"):
        res = w_src / dz
    return np.where(inside & (res > 0), res, NAN).astype(np.float32)


def integrate_warped_rgb(dw, rw, gw, bw, ww, depth_dst, colors_dst, weight_dst):
    """integrateWarpedRGBKernel, warping_registration.cu:672-712; updates the three destination arrays in place"""
    dw, rw, gw, bw, ww = map(_f, (dw, rw, gw, bw, ww))
    ok = ~(np.isnan(dw) | np.isnan(rw) | np.isnan(gw) | np.isnan(bw))
    new = ok & np.isnan(depth_dst)
    with np.errstate(invalid="ignore"):
        fuse = ok & ~np.isnan(depth_dst) & ((depth_dst - dw) < F(0.0075)) & ((dw - depth_dst) < F(0.0075))
    nw = (weight_dst + ww).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        d_fused = (depth_dst * weight_dst + dw * ww) / nw
        for ch, src in enumerate((rw, gw, bw)):
            c = colors_dst[:, :, ch]
            cf = np.rint((c.astype(np.float32) * weight_dst + src * ww) / nw)
            cn = np.rint(src)
            out = np.where(fuse, cf, np.where(new, cn, c.astype(np.float32)))
            colors_dst[:, :, ch] = np.nan_to_num(out, nan=0.0).astype(np.int64).astype(np.uint8)
    depth_out = np.where(fuse, d_fused, np.where(new, dw, depth_dst)).astype(np.float32)
    weight_out = np.where(fuse, nw, np.where(new, ww, weight_dst)).astype(np.float32)
    depth_dst[...] = depth_out
    weight_dst[...] = weight_out


def generate_image(vmap, nmap, light, rgb=None):
    """generateImageKernel / generateImageRGBKernel, image_generator.cu:60-181 (one light source)"""
    vmap, nmap, light = _f(vmap), _f(nmap), _f(light)
    rows = vmap.shape[0] // 3
    vx, vy, vz = vmap[:rows], vmap[rows:2 * rows], vmap[2 * rows:]
    nx, ny, nz = nmap[:rows], nmap[rows:2 * rows], nmap[2 * rows:]
    valid = ~(np.isnan(vx) | np.isnan(nx))
    with np.errstate(invalid="ignore", divide="ignore"):
        dx, dy, dz = light[0] - vx, light[1] - vy, light[2] - vz
        rn = F(1) / np.sqrt(dx * dx + dy * dy + dz * dz)
        weight = np.abs((dx * rn) * nx + (dy * rn) * ny + (dz * rn) * nz)
        br = np.clip(np.nan_to_num(F(205) * weight, nan=0.0).astype(np.int64) + 50, 0, 255)
    out = np.zeros((rows, vmap.shape[1], 3), dtype=np.uint8)
    if rgb is None:
        for ch in range(3):
            out[:, :, ch] = np.where(valid, br, 0)
    else:
        br_f = br.astype(np.float32) / F(255)
        for ch in range(3):
            out[:, :, ch] = np.where(valid, np.rint(rgb[:, :, ch].astype(np.float32) * br_f), 0)
    return out
