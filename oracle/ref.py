"""ctypes binding of oracle/_ref/libref_oracle.so: the reference's OWN CUDA device layer (src/cuda/*.cu +
ThirdParty/pcl_gpu_containers, compiled verbatim for sm_100a by oracle/Makefile) behind the C shim
oracle/ref_shim.cpp.  TEST INFRASTRUCTURE ONLY; needs a GPU.  Inputs are dense torch CUDA tensors."""
import ctypes as C
import os

import numpy as np
import torch

from . import (REF_LIB_PATH, AlignConfig, FrameStats, IterTrace, Pyramids, SystemParams, _fill_pyramids,  # noqa: F401
               error_geometry, trace_to_dicts)
from . import align as _align_glue

_lib = None


def available():
    return os.path.exists(REF_LIB_PATH) and torch.cuda.is_available()


def lib(device=0):
    global _lib
    if _lib is None:
        l = C.CDLL(REF_LIB_PATH)
        for n in ("ref_pyr_down", "ref_gradient", "ref_bilateral", "ref_warp_invdepth", "ref_warp_intensity",
                  "ref_warp_invdepth_weighted", "ref_integrate_warped_frame", "ref_visibility_ratio", "ref_build_system"):
            getattr(l, n).restype = C.c_float
        rc = l.ref_init(device)
        assert rc == 0, "ref_init failed: %d" % rc
        _lib = l
    return _lib


def _f32(rp):
    return np.ascontiguousarray(np.reshape(rp, -1), dtype=np.float32)


def _np(a):
    return a.ctypes.data_as(C.c_void_p)


def _dense(t):
    assert t.is_cuda and t.is_contiguous()
    torch.cuda.synchronize()
    return C.c_void_p(t.data_ptr())


def _sz(n):
    return C.c_size_t(n)


def convert_depth_to_invdepth(depth_u16, factor_depth=1.0):
    rows, cols = depth_u16.shape
    out = torch.empty(rows, cols, device=depth_u16.device)
    lib().ref_convert_depth_to_invdepth(_dense(depth_u16), _sz(cols * 2), _dense(out), _sz(cols * 4), rows, cols,
                                        C.c_float(factor_depth))
    return out


def compute_intensity(rgb_u8):
    rows, cols, _ = rgb_u8.shape
    out = torch.empty(rows, cols, device=rgb_u8.device)
    lib().ref_compute_intensity(_dense(rgb_u8), _sz(cols * 3), _dense(out), _sz(cols * 4), rows, cols)
    return out


def pyr_down(src):
    rows, cols = src.shape
    out = torch.empty(rows // 2, cols // 2, device=src.device)
    lib().ref_pyr_down(_dense(src), _sz(cols * 4), rows, cols, _dense(out), _sz((cols // 2) * 4))
    return out


def gradient(src):
    rows, cols = src.shape
    gx, gy = torch.empty_like(src), torch.empty_like(src)
    lib().ref_gradient(_dense(src), _sz(cols * 4), rows, cols, _dense(gx), _dense(gy), _sz(cols * 4))
    return gx, gy


def bilateral(src, sigma):
    """bilateralFilter of the reference.  NOTE (reference defect): bilateralKernel declares x, y as unsigned
    (src/cuda/filters.cu:91-92), so `max(y - RADIUS, 0)` wraps for y < 2 (and x < 2) and the window starts at row /
    column -2: the kernel reads out of bounds above the image and the tail of the previous row.  On this B200 pool
    that is an illegal memory access when the image starts an allocation (cudaSafeCall then calls exit(0)).  The
    input is therefore embedded in a NaN-filled buffer (4 rows above, 4 columns of row padding): the stray taps
    hit NaN, which the kernel skips, i.e. exactly the clamped window the code intends and oracle.c restates."""
    rows, cols = src.shape
    pad = torch.full((rows + 8, cols + 4), float("nan"), device=src.device)
    pad[4:rows + 4, :cols] = src
    view = pad[4:rows + 4, :cols]
    out = torch.empty_like(src)
    torch.cuda.synchronize()
    lib().ref_bilateral(C.c_void_p(view.data_ptr()), _sz((cols + 4) * 4), rows, cols, _dense(out), _sz(cols * 4),
                        C.c_float(sigma))
    return out


def _warp(fn, src, prev, Rp, tp):
    rows, cols = prev.shape
    out = torch.empty_like(prev)
    Rp, tp = _f32(Rp), _f32(tp)
    fn(_dense(src), _dense(prev), _dense(out), _sz(cols * 4), rows, cols, _np(Rp), _np(tp))
    return out


def warp_invdepth(src, prev, Rp, tp):
    return _warp(lib().ref_warp_invdepth, src, prev, Rp, tp)


def warp_intensity(src, prev, Rp, tp):
    return _warp(lib().ref_warp_intensity, src, prev, Rp, tp)


def warp_invdepth_weighted(src, prev, weight_inout, Rp, tp):
    rows, cols = prev.shape
    out = torch.empty_like(prev)
    Rp, tp = _f32(Rp), _f32(tp)
    lib().ref_warp_invdepth_weighted(_dense(src), _dense(prev), _dense(out), _dense(weight_inout), _sz(cols * 4), rows,
                                     cols, _np(Rp), _np(tp))
    return out


def integrate_warped_frame(wsrc, wweight, dst_inout, dweight_inout):
    rows, cols = wsrc.shape
    lib().ref_integrate_warped_frame(_dense(wsrc), _dense(wweight), _dense(dst_inout), _dense(dweight_inout),
                                     _sz(cols * 4), rows, cols)


def visibility_ratio(depth_src, depth_dst, Rp, tp, with_mask=False):
    rows, cols = depth_src.shape
    Rp, tp = _f32(Rp), _f32(tp)
    mask = torch.zeros(rows, cols, dtype=torch.uint8, device=depth_src.device) if with_mask else None
    r = lib().ref_visibility_ratio(_dense(depth_src), _dense(depth_dst), _sz(cols * 4), rows, cols, _np(Rp), _np(tp),
                                   _dense(mask) if with_mask else None, _sz(cols))
    return (r, mask) if with_mask else r


def compute_error(im1, im0, nsamples=9999999):
    rows, cols = im0.shape
    kr, kc, _ = error_geometry(rows, cols, nsamples)
    err = torch.empty(kr * kc, device=im0.device)
    lib().ref_compute_error(_dense(im1), _dense(im0), _sz(cols * 4), rows, cols, nsamples, _dense(err))
    return err


def sigma_nu_student(err, bias, sigma, mest=3):
    b, s, nu = C.c_float(bias), C.c_float(sigma), C.c_float(0)
    lib().ref_sigma_nu_student(_dense(err), err.numel(), C.byref(b), C.byref(s), C.byref(nu), mest)
    return b.value, s.value, nu.value


def nu_student(err, bias, sigma):
    nu = C.c_float(0)
    lib().ref_nu_student(_dense(err), err.numel(), C.c_float(bias), C.c_float(sigma), C.byref(nu))
    return nu.value


def sigma_pdf(err, bias, sigma, mest):
    b, s = C.c_float(bias), C.c_float(sigma)
    lib().ref_sigma_pdf(_dense(err), err.numel(), C.byref(b), C.byref(s), mest)
    return b.value, s.value


def chi_square(err_int, err_depth, sigma_int, sigma_depth, mest):
    x, y, z = C.c_float(), C.c_float(), C.c_float()
    lib().ref_chi_square(_dense(err_int), _dense(err_depth), err_int.numel(), C.c_float(sigma_int), C.c_float(sigma_depth),
                         mest, C.byref(x), C.byref(y), C.byref(z))
    return x.value, y.value, z.value


def build_system(W0, I0, gWx, gWy, gIx, gIy, W1, I1, params):
    rows, cols = W0.shape
    A, b = np.zeros(36), np.zeros(6)
    lib().ref_build_system(*[_dense(m) for m in (W0, I0, gWx, gWy, gIx, gIy, W1, I1)], _sz(cols * 4), rows, cols,
                           C.byref(params), _np(A), _np(b))
    return A.reshape(6, 6), b


def vmap(depth_inv, fx, fy, cx, cy):
    rows, cols = depth_inv.shape
    out = torch.full((3 * rows, cols), float("nan"), device=depth_inv.device)
    lib().ref_vmap(_dense(depth_inv), _sz(cols * 4), rows, cols, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                   _dense(out), _sz(cols * 4))
    return out


def nmap_gradients(depth_inv, gx, gy, fx, fy, cx, cy):
    rows, cols = depth_inv.shape
    out = torch.full((3 * rows, cols), float("nan"), device=depth_inv.device)
    lib().ref_nmap_gradients(_dense(depth_inv), _dense(gx), _dense(gy), _sz(cols * 4), rows, cols, C.c_float(fx),
                             C.c_float(fy), C.c_float(cx), C.c_float(cy), _dense(out), _sz(cols * 4))
    return out


def prepare_keyframe(W0, I0, levels, tracker=True):
    """Keyframe pyramids + gradients computed by the reference's own kernels (device tensors)."""
    def pyr(img):
        out = [img.contiguous()]
        for _ in range(1, levels):
            out.append(pyr_down(out[-1]))
        return out
    kf = dict(W=pyr(W0), I=pyr(I0))
    g = [gradient(m) for m in kf["W"]]
    kf["gWx"], kf["gWy"] = [a for a, _ in g], [b for _, b in g]
    g = [gradient(m) for m in kf["I"]]
    kf["gIx"], kf["gIy"] = [a for a, _ in g], [b for _, b in g]
    if tracker:
        Wf, If = pyr(bilateral(kf["W"][0], 2.0 * 0.0025)), pyr(bilateral(kf["I"][0], 3.0))
        g = [gradient(m) for m in Wf]
        kf["cgWx"], kf["cgWy"] = [a for a, _ in g], [b for _, b in g]
        g = [gradient(m) for m in If]
        kf["cgIx"], kf["cgIy"] = [a for a, _ in g], [b for _, b in g]
    return kf


def prepare_current(W, I, levels):
    def pyr(img):
        out = [img.contiguous()]
        for _ in range(1, levels):
            out.append(pyr_down(out[-1]))
        return out
    return dict(W=pyr(W), I=pyr(I))


def align(cfg, kf, cur, R=None, t=None):
    """Reference kernels + restated host loop (ref_align in ref_shim.cpp)."""
    torch.cuda.synchronize()
    l = lib()
    return _align_glue(cfg, kf, cur, R, t, fn=l.ref_align, ptr_of=lambda a: a.data_ptr())


# ---- custom-calibration ingest, colour fusion, shaded previews -------------------------------------
def _intr9(intr):
    return np.array([intr[k] if k in intr else 0.0 for k in ("fx", "fy", "cx", "cy", "k1", "k2", "k3", "k4", "k5")],
                    dtype=np.float32)


def undistort_intensity(src, intr):
    rows, cols = src.shape
    out = torch.empty(rows, cols, device=src.device)
    lib().ref_undistort_intensity(_dense(src), _dense(out), _sz(cols * 4), rows, cols, _np(_intr9(intr)))
    return out


def undistort_depthinv(src, intr, dp):
    rows, cols = src.shape
    out, scratch = torch.empty(rows, cols, device=src.device), torch.empty(rows, cols, device=src.device)
    d = np.array([dp["c1"], dp["c0"]] + list(dp["q0"]) + list(dp["q1"]), dtype=np.float32)
    lib().ref_undistort_depthinv(_dense(src), _dense(scratch), _dense(out), _sz(cols * 4), rows, cols, _np(_intr9(intr)),
                                 _np(d), int(dp["xshift"]), int(dp["yshift"]))
    return out


def register_depthinv(src, dRc_proj, t_dc_proj, cRd_proj):
    rows, cols = src.shape
    out = torch.empty(rows, cols, device=src.device)
    inter = torch.empty(3 * rows, 3 * cols, device=src.device)
    inter_i = torch.empty(3 * rows, 3 * cols, dtype=torch.int32, device=src.device)
    lib().ref_register_depthinv(_dense(src), _dense(inter), _dense(inter_i), _sz(3 * cols * 4), _dense(out), _sz(cols * 4),
                                rows, cols, _np(_f32(dRc_proj)), _np(_f32(t_dc_proj)), _np(_f32(cRd_proj)))
    return out


def integrate_warped_rgb(dw, rw, gw, bw, ww, depth_dst, colors_dst, weight_dst):
    rows, cols = dw.shape
    lib().ref_integrate_warped_rgb(_dense(dw), _dense(rw), _dense(gw), _dense(bw), _dense(ww), _dense(depth_dst),
                                   _dense(colors_dst), _sz(cols * 3), _dense(weight_dst), _sz(cols * 4), rows, cols)


def generate_image(vmap, nmap, light, rgb=None):
    rows, cols = vmap.shape[0] // 3, vmap.shape[1]
    out = torch.empty(rows, cols, 3, dtype=torch.uint8, device=vmap.device)
    lib().ref_generate_image(_dense(vmap), _dense(nmap), _sz(cols * 4), _dense(rgb) if rgb is not None else None,
                             _sz(cols * 3), _np(_f32(light)), _dense(out), _sz(cols * 3), rows, cols)
    return out
