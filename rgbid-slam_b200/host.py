"""Thin torch-tensor front end over the C ABI (capi.py).  PyTorch is plumbing here: it owns device memory and
the CUDA stream; every computation happens inside librgbid_b200.so."""
import ctypes as C

import numpy as np
import torch

from . import capi

_F = C.c_float


def _fp(arr):
    return arr.ctypes.data_as(capi.c_float_p)


def _dp(arr):
    return arr.ctypes.data_as(capi.c_double_p)


def _pitch(t):
    assert t.is_cuda and t.stride(-1) == 1, "expected a CUDA tensor with contiguous rows"
    return t.stride(-2) * t.element_size()


def default_iterations(levels, mode):
    """iters[] prefix of the reference: {10,5,3,0,0,0} tracker (src/visodo.cpp:65), {5,5,3,0,0,0} align
    (src/keyframe_align.cpp:44)."""
    base = [10, 5, 3, 0, 0, 0, 0, 0] if mode == capi.MODE_TRACKER else [5, 5, 3, 0, 0, 0, 0, 0]
    return base[:levels]


class Context:
    """rgbid_ctx bound to a torch stream, so torch events / allocations and our kernels share one stream."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("rgbid_slam_b200 needs a CUDA device (no CPU fallback)")
        self.lib = capi.load()
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(self.device)
        h = C.c_void_p()
        capi.check(self.lib.rgbid_ctx_create(C.byref(h), device, C.c_void_p(self.stream.cuda_stream)), "ctx_create")
        self.h = h
        self._children = []

    def _register(self, child):
        import weakref
        self._children.append(weakref.ref(child))

    def close(self):
        if self.h is not None:
            for ref in self._children:  # handles created on this context must go first
                child = ref()
                if child is not None:
                    child.close()
            self._children = []
            self.lib.rgbid_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        capi.check(self.lib.rgbid_ctx_sync(self.h), "ctx_sync")

    @property
    def launches(self):
        return int(self.lib.rgbid_ctx_launch_count(self.h))

    def _enter(self, *tensors):
        # inputs may have been produced on torch's current stream
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def _leave(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ---- drop-in ops (one per reference bridge function) -------------------------------------------------
    def convert_depth_to_invdepth(self, depth_u16, factor_depth=1.0):
        rows, cols = depth_u16.shape
        out = self.empty(rows, cols)
        self._enter()
        capi.check(self.lib.rgbid_convert_depth_to_invdepth(self.h, depth_u16.data_ptr(), _pitch(depth_u16), out.data_ptr(),
                                                           _pitch(out), rows, cols, factor_depth), "convert_depth")
        self._leave()
        return out

    def compute_intensity(self, rgb_u8):
        rows, cols, _ = rgb_u8.shape
        out = self.empty(rows, cols)
        self._enter()
        capi.check(self.lib.rgbid_compute_intensity(self.h, rgb_u8.data_ptr(), rgb_u8.stride(0), out.data_ptr(), _pitch(out),
                                                    rows, cols), "compute_intensity")
        self._leave()
        return out

    def decompose_rgb(self, rgb_u8):
        rows, cols, _ = rgb_u8.shape
        r, g, b = self.empty(rows, cols), self.empty(rows, cols), self.empty(rows, cols)
        self._enter()
        capi.check(self.lib.rgbid_decompose_rgb(self.h, rgb_u8.data_ptr(), rgb_u8.stride(0), r.data_ptr(), g.data_ptr(),
                                                b.data_ptr(), _pitch(r), rows, cols), "decompose_rgb")
        self._leave()
        return r, g, b

    def pyr_down(self, src):
        rows, cols = src.shape
        out = self.empty(rows // 2, cols // 2)
        self._enter()
        capi.check(self.lib.rgbid_pyr_down(self.h, src.data_ptr(), _pitch(src), rows, cols, out.data_ptr(), _pitch(out)), "pyr_down")
        self._leave()
        return out

    def compute_gradient(self, src):
        rows, cols = src.shape
        gx, gy = self.empty(rows, cols), self.empty(rows, cols)
        self._enter()
        capi.check(self.lib.rgbid_compute_gradient(self.h, src.data_ptr(), _pitch(src), rows, cols, gx.data_ptr(), gy.data_ptr(),
                                                   _pitch(gx)), "compute_gradient")
        self._leave()
        return gx, gy

    def copy_image(self, src, dst=None):
        """copyImage (src/internal.h:260): pitched device-to-device copy; dst may be a row-pitched view."""
        rows, cols = src.shape
        out = self.empty(rows, cols) if dst is None else dst
        self._enter()
        capi.check(self.lib.rgbid_copy_image(self.h, src.data_ptr(), _pitch(src), out.data_ptr(), _pitch(out), rows, cols), "copy_image")
        self._leave()
        return out

    def fill_image(self, dst, value):
        """initialiseDeviceMemory2D<float> / initialiseWeightKeyframe (src/internal.h:271-274)."""
        rows, cols = dst.shape
        self._enter()
        capi.check(self.lib.rgbid_fill_image(self.h, dst.data_ptr(), _pitch(dst), rows, cols, float(value)), "fill_image")
        self._leave()
        return dst

    def bilateral_filter(self, src, sigma):
        rows, cols = src.shape
        out = self.empty(rows, cols)
        self._enter()
        capi.check(self.lib.rgbid_bilateral_filter(self.h, src.data_ptr(), _pitch(src), rows, cols, out.data_ptr(), _pitch(out),
                                                   sigma), "bilateral_filter")
        self._leave()
        return out

    def create_vmap(self, depth_inv, fx, fy, cx, cy):
        rows, cols = depth_inv.shape
        out = torch.full((3 * rows, cols), float("nan"), device=self.device)
        self._enter()
        capi.check(self.lib.rgbid_create_vmap(self.h, depth_inv.data_ptr(), _pitch(depth_inv), rows, cols, fx, fy, cx, cy,
                                              out.data_ptr(), _pitch(out)), "create_vmap")
        self._leave()
        return out

    def create_nmap_gradients(self, depth_inv, gx, gy, fx, fy, cx, cy):
        rows, cols = depth_inv.shape
        out = torch.full((3 * rows, cols), float("nan"), device=self.device)
        self._enter()
        capi.check(self.lib.rgbid_create_nmap_gradients(self.h, depth_inv.data_ptr(), gx.data_ptr(), gy.data_ptr(),
                                                        _pitch(depth_inv), rows, cols, fx, fy, cx, cy, out.data_ptr(),
                                                        _pitch(out)), "create_nmap_gradients")
        self._leave()
        return out

    # ---- custom-calibration ingest, colour fusion, previews -------------------------------------
    @staticmethod
    def _intr(intr):
        return capi.Intr(*[float(intr.get(k, 0.0)) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "k3", "k4", "k5")])

    def undistort_intensity(self, src, intr):
        rows, cols = src.shape
        out = self.empty(rows, cols)
        i = self._intr(intr)
        self._enter()
        capi.check(self.lib.rgbid_undistort_intensity(self.h, src.data_ptr(), _pitch(src), out.data_ptr(), _pitch(out), rows,
                                                      cols, C.byref(i)), "undistort_intensity")
        self._leave()
        return out

    def undistort_depthinv(self, src, intr, dp):
        rows, cols = src.shape
        out = self.empty(rows, cols)
        i = self._intr(intr)
        d = capi.DepthDist(float(dp["c1"]), float(dp["c0"]), (C.c_float * 9)(*dp["q0"]), (C.c_float * 9)(*dp["q1"]),
                           int(dp["xshift"]), int(dp["yshift"]))
        self._enter()
        capi.check(self.lib.rgbid_undistort_depthinv(self.h, src.data_ptr(), _pitch(src), out.data_ptr(), _pitch(out), rows,
                                                     cols, C.byref(i), C.byref(d)), "undistort_depthinv")
        self._leave()
        return out

    def register_depthinv(self, src, dRc_proj, t_dc_proj, cRd_proj):
        rows, cols = src.shape
        out = self.empty(rows, cols)
        a, t, b = (np.ascontiguousarray(np.reshape(m, -1), dtype=np.float32) for m in (dRc_proj, t_dc_proj, cRd_proj))
        self._enter()
        capi.check(self.lib.rgbid_register_depthinv(self.h, src.data_ptr(), _pitch(src), out.data_ptr(), _pitch(out), rows,
                                                    cols, _fp(a), _fp(t), _fp(b)), "register_depthinv")
        self._leave()
        return out

    def integrate_warped_rgb(self, dw, rw, gw, bw, ww, depth_dst, colors_dst, weight_dst):
        rows, cols = dw.shape
        self._enter()
        capi.check(self.lib.rgbid_integrate_warped_rgb(self.h, dw.data_ptr(), rw.data_ptr(), gw.data_ptr(), bw.data_ptr(),
                                                       ww.data_ptr(), depth_dst.data_ptr(), colors_dst.data_ptr(),
                                                       colors_dst.stride(0), weight_dst.data_ptr(), _pitch(dw), rows, cols),
                   "integrate_warped_rgb")
        self._leave()

    def generate_image(self, vmap, nmap, light, rgb=None):
        rows, cols = vmap.shape[0] // 3, vmap.shape[1]
        out = torch.empty(rows, cols, 3, dtype=torch.uint8, device=self.device)
        l = np.ascontiguousarray(light, dtype=np.float32)
        self._enter()
        capi.check(self.lib.rgbid_generate_image(self.h, vmap.data_ptr(), nmap.data_ptr(), _pitch(vmap),
                                                 rgb.data_ptr() if rgb is not None else None,
                                                 rgb.stride(0) if rgb is not None else 0, _fp(l), out.data_ptr(), out.stride(0),
                                                 rows, cols), "generate_image")
        self._leave()
        return out

    def _warp(self, fn, src, prev, Rp, tp, name):
        rows, cols = prev.shape
        out = self.empty(rows, cols)
        Rp = np.ascontiguousarray(Rp, dtype=np.float32).reshape(9)
        tp = np.ascontiguousarray(tp, dtype=np.float32).reshape(3)
        self._enter()
        capi.check(fn(self.h, src.data_ptr(), _pitch(src), prev.data_ptr(), _pitch(prev), out.data_ptr(), _pitch(out), rows,
                      cols, _fp(Rp), _fp(tp)), name)
        self._leave()
        return out

    def warp_invdepth(self, src, prev, Rp, tp):
        return self._warp(self.lib.rgbid_warp_invdepth, src, prev, Rp, tp, "warp_invdepth")

    def warp_intensity(self, src, prev, Rp, tp):
        return self._warp(self.lib.rgbid_warp_intensity, src, prev, Rp, tp, "warp_intensity")

    def warp_invdepth_weighted(self, src, prev, weight_inout, Rp, tp):
        rows, cols = prev.shape
        out = self.empty(rows, cols)
        Rp = np.ascontiguousarray(Rp, dtype=np.float32).reshape(9)
        tp = np.ascontiguousarray(tp, dtype=np.float32).reshape(3)
        self._enter()
        capi.check(self.lib.rgbid_warp_invdepth_weighted(self.h, src.data_ptr(), _pitch(src), prev.data_ptr(), _pitch(prev),
                                                         out.data_ptr(), _pitch(out), weight_inout.data_ptr(),
                                                         _pitch(weight_inout), rows, cols, _fp(Rp), _fp(tp)),
                   "warp_invdepth_weighted")
        self._leave()
        return out

    def integrate_warped_frame(self, wsrc, wweight, dst_inout, dweight_inout):
        rows, cols = wsrc.shape
        self._enter()
        capi.check(self.lib.rgbid_integrate_warped_frame(self.h, wsrc.data_ptr(), _pitch(wsrc), wweight.data_ptr(),
                                                         _pitch(wweight), dst_inout.data_ptr(), _pitch(dst_inout),
                                                         dweight_inout.data_ptr(), _pitch(dweight_inout), rows, cols),
                   "integrate_warped_frame")
        self._leave()

    def visibility_ratio(self, depth_src, depth_dst, Rp, tp, with_mask=False):
        rows, cols = depth_src.shape
        Rp = np.ascontiguousarray(Rp, dtype=np.float32).reshape(9)
        tp = np.ascontiguousarray(tp, dtype=np.float32).reshape(3)
        mask = torch.zeros(rows, cols, dtype=torch.uint8, device=self.device) if with_mask else None
        ratio = _F(0)
        self._enter()
        capi.check(self.lib.rgbid_visibility_ratio(self.h, depth_src.data_ptr(), _pitch(depth_src), depth_dst.data_ptr(),
                                                   _pitch(depth_dst), rows, cols, _fp(Rp), _fp(tp),
                                                   mask.data_ptr() if with_mask else None, cols if with_mask else 0,
                                                   C.byref(ratio)), "visibility_ratio")
        self._leave()
        return (ratio.value, mask) if with_mask else ratio.value

    def error_geometry(self, rows, cols, nsamples):
        kr, kc, s = C.c_int(), C.c_int(), C.c_int()
        capi.check(self.lib.rgbid_error_geometry(rows, cols, nsamples, C.byref(kr), C.byref(kc), C.byref(s)), "error_geometry")
        return kr.value, kc.value, s.value

    def compute_error(self, im1, im0, nsamples=9999999):
        rows, cols = im0.shape
        kr, kc, _ = self.error_geometry(rows, cols, nsamples)
        err = self.empty(kr * kc)
        n = C.c_int()
        self._enter()
        capi.check(self.lib.rgbid_compute_error(self.h, im1.data_ptr(), _pitch(im1), im0.data_ptr(), _pitch(im0), rows, cols,
                                                nsamples, err.data_ptr(), C.byref(n)), "compute_error")
        self._leave()
        return err

    def sigma_nu_student(self, err, bias, sigma, mest=capi.STUDENT):
        b, s, nu = _F(bias), _F(sigma), _F(0)
        self._enter()
        capi.check(self.lib.rgbid_sigma_nu_student(self.h, err.data_ptr(), err.numel(), C.byref(b), C.byref(s), C.byref(nu), mest),
                   "sigma_nu_student")
        return b.value, s.value, nu.value

    def nu_student(self, err, bias, sigma):
        nu = _F(0)
        self._enter()
        capi.check(self.lib.rgbid_nu_student(self.h, err.data_ptr(), err.numel(), bias, sigma, C.byref(nu)), "nu_student")
        return nu.value

    def sigma_pdf(self, err, bias, sigma, mest):
        b, s = _F(bias), _F(sigma)
        self._enter()
        capi.check(self.lib.rgbid_sigma_pdf(self.h, err.data_ptr(), err.numel(), C.byref(b), C.byref(s), mest), "sigma_pdf")
        return b.value, s.value

    def chi_square(self, err_int, err_depth, sigma_int, sigma_depth, mest):
        a, b, c = _F(0), _F(0), _F(0)
        self._enter()
        capi.check(self.lib.rgbid_chi_square(self.h, err_int.data_ptr(), err_depth.data_ptr(), err_int.numel(), sigma_int,
                                             sigma_depth, mest, C.byref(a), C.byref(b), C.byref(c)), "chi_square")
        return a.value, b.value, c.value

    def build_system(self, W0, I0, gWx, gWy, gIx, gIy, W1, I1, params):
        rows, cols = W0.shape
        A = np.zeros(36, dtype=np.float64)
        b = np.zeros(6, dtype=np.float64)
        maps = [W0, I0, gWx, gWy, gIx, gIy, W1, I1]
        pitch = _pitch(W0)
        self._enter()
        if all(_pitch(m) == pitch for m in maps):
            capi.check(self.lib.rgbid_build_system(self.h, *[m.data_ptr() for m in maps], pitch, rows, cols, C.byref(params),
                                                   _dp(A), _dp(b)), "build_system")
        else:  # every map with its own row pitch (PtrStep::step of the reference)
            p8 = (C.c_size_t * 8)(*[_pitch(m) for m in maps])
            capi.check(self.lib.rgbid_build_system_pitched(self.h, *[m.data_ptr() for m in maps], p8, rows, cols,
                                                           C.byref(params), _dp(A), _dp(b)), "build_system_pitched")
        return A.reshape(6, 6), b


def make_align_config(rows, cols, levels, mode, batch=1, fx=525.0, fy=525.0, cx=319.5, cy=239.5, iterations=None,
                      finest_level=0, mestimator=capi.STUDENT, weighting=capi.INDEPENDENT,
                      sigma_estimator=capi.SIGMA_PDF, nsamples=None, factor_depth=1.0, warp_first=0,
                      termination=capi.TERM_ALL_ITERS, conv_eps=0.0):
    cfg = capi.AlignConfig()
    cfg.rows, cfg.cols, cfg.levels, cfg.finest_level = rows, cols, levels, finest_level
    its = iterations if iterations is not None else default_iterations(levels, mode)
    for i in range(capi.MAX_LEVELS):
        cfg.iterations[i] = its[i] if i < len(its) else 0
    cfg.batch, cfg.mode = batch, mode
    cfg.mestimator, cfg.weighting, cfg.sigma_estimator = mestimator, weighting, sigma_estimator
    cfg.nsamples = nsamples if nsamples is not None else (10000 if mode == capi.MODE_TRACKER else 19200)
    cfg.fx, cfg.fy, cfg.cx, cfg.cy = fx, fy, cx, cy
    cfg.factor_depth = factor_depth
    cfg.with_fusion = 0
    cfg.warp_first = int(warp_first)  # tracker: WARP_ORDER = warpFirst (src/visodo.cpp:1078-1105); 0 = pyrFirst
    cfg.termination, cfg.conv_eps = int(termination), float(conv_eps)  # include/rgbid_b200.h RGBID_TERM_*
    return cfg


def trace_to_dicts(trace, n):
    out = []
    for i in range(n):
        T = trace[i]
        out.append(dict(level=T.level, iter=T.iter, sums27=np.array(T.sums27[:]), sigma_int=T.sigma_int,
                        sigma_depthinv=T.sigma_depthinv, bias_int=T.bias_int, bias_depthinv=T.bias_depthinv,
                        nu_int=T.nu_int, nu_depthinv=T.nu_depthinv, irls_iters_int=T.irls_iters_int,
                        irls_iters_depthinv=T.irls_iters_depthinv, x=np.array(T.x[:]),
                        R=np.array(T.R[:]).reshape(3, 3), t=np.array(T.t[:])))
    return out


class Aligner:
    """rgbid_aligner: device-resident coarse-to-fine alignment of `batch` frame pairs."""

    MAP_NAMES = {"W_kf": 0, "I_kf": 1, "gWx": 2, "gWy": 3, "gIx": 4, "gIy": 5, "W_cur": 6, "I_cur": 7,
                 "cgWx": 8, "cgWy": 9, "cgIx": 10, "cgIy": 11}

    def __init__(self, ctx, cfg):
        self.ctx, self.cfg, self.lib = ctx, cfg, ctx.lib
        h = C.c_void_p()
        capi.check(self.lib.rgbid_aligner_create(ctx.h, C.byref(cfg), C.byref(h)), "aligner_create")
        self.h = h
        self.batch = cfg.batch
        self.niters = self.lib.rgbid_aligner_num_iterations(h)
        ctx._register(self)

    def close(self):
        if self.h is not None:
            if self.ctx.h is not None:
                self.lib.rgbid_aligner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _set(self, fn, index, a, b, name):
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a)
            b = np.ascontiguousarray(b)
            pa, pb = a.ctypes.data, b.ctypes.data
            sa, sb = a.strides[0], b.strides[0]
            host = 1
        else:
            self.ctx._enter()
            pa, pb, sa, sb, host = a.data_ptr(), b.data_ptr(), a.stride(0) * a.element_size(), b.stride(0) * b.element_size(), 0
        capi.check(fn(self.h, index, pa, sa, pb, sb, host), name)
        if host:
            self.ctx.sync()  # the host arrays may be released by the caller

    def set_keyframe(self, index, depthinv, intensity):
        self._set(self.lib.rgbid_aligner_set_keyframe, index, depthinv, intensity, "set_keyframe")

    def set_current(self, index, depthinv, intensity):
        self._set(self.lib.rgbid_aligner_set_current, index, depthinv, intensity, "set_current")

    def set_current_rgbd(self, index, depth_u16, rgb_u8):
        self._set(self.lib.rgbid_aligner_set_current_rgbd, index, depth_u16, rgb_u8, "set_current_rgbd")

    def current_to_keyframe(self, index):
        capi.check(self.lib.rgbid_aligner_current_to_keyframe(self.h, index), "current_to_keyframe")

    def run(self, R=None, t=None, want_trace=False):
        B = self.batch
        R = np.tile(np.eye(3), (B, 1, 1)) if R is None else np.array(R, dtype=np.float64).reshape(B, 3, 3)
        t = np.zeros((B, 3)) if t is None else np.array(t, dtype=np.float64).reshape(B, 3)
        R = np.ascontiguousarray(R)
        t = np.ascontiguousarray(t)
        cov = np.zeros((B, 6, 6))
        status = np.zeros(B, dtype=np.int32)
        ntr = (self.niters + 1) * B
        trace = (capi.IterTrace * ntr)() if want_trace else None
        capi.check(self.lib.rgbid_aligner_run(self.h, _dp(R), _dp(t), _dp(cov), status.ctypes.data_as(capi.c_int_p), trace),
                   "aligner_run")
        out = dict(R=R, t=t, cov=cov, status=status)
        if want_trace:
            out["trace"] = [trace_to_dicts(trace[b * (self.niters + 1):(b + 1) * (self.niters + 1)], self.niters + 1)
                            for b in range(B)]
        stats = np.zeros((B, 3), dtype=np.float32)
        self.lib.rgbid_aligner_frame_stats(self.h, _fp(stats))
        out["stats"] = stats
        done = np.zeros((B, capi.MAX_LEVELS), dtype=np.int32)
        capi.check(self.lib.rgbid_aligner_iterations_done(self.h, done.ctypes.data_as(capi.c_int_p)), "iterations_done")
        out["iterations_done"] = done
        if want_trace and self.cfg.termination != capi.TERM_ALL_ITERS:
            # a level that ended early leaves the slots of its remaining iterations untouched: keep the executed ones
            its = [self.cfg.iterations[l] for l in range(self.cfg.levels)]
            for b in range(B):
                keep, k = [], 0
                for l in range(self.cfg.levels - 1, self.cfg.finest_level - 1, -1):
                    keep += list(range(k, k + int(done[b, l])))
                    k += its[l]
                keep.append(self.niters)  # covariance pass
                out["trace"][b] = [out["trace"][b][i] for i in keep]
        return out

    def export_systems(self, out):
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == self.batch * 48
        capi.check(self.lib.rgbid_aligner_export_systems(self.h, out.data_ptr()), "export_systems")
        return out

    def enqueue(self, R, t):
        R = np.ascontiguousarray(R, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        capi.check(self.lib.rgbid_aligner_enqueue(self.h, _dp(R), _dp(t)), "aligner_enqueue")

    def time_build(self, level=0, reps=20, scale=False):
        ms = _F(0)
        fn = self.lib.rgbid_aligner_time_scale if scale else self.lib.rgbid_aligner_time_build
        capi.check(fn(self.h, level, reps, C.byref(ms)), "aligner_time_build")
        return ms.value

    def map(self, name, level, index=0):
        """torch view (rows x cols, row-pitched) of an internal pyramid map."""
        p, pitch = C.c_void_p(), C.c_size_t()
        capi.check(self.lib.rgbid_aligner_map(self.h, self.MAP_NAMES[name], level, index, C.byref(p), C.byref(pitch)), "aligner_map")
        rows, cols = self.cfg.rows >> level, self.cfg.cols >> level
        return _wrap_device(p.value, rows, cols, pitch.value, self.ctx.device)


class _CudaPtr:
    def __init__(self, ptr, shape, strides, typestr):
        self.__cuda_array_interface__ = dict(shape=shape, strides=strides, typestr=typestr, data=(ptr, False), version=3)


def _wrap_device(ptr, rows, cols, pitch, device, dtype=torch.float32):
    """Dense torch copy of a pitched device image owned by the library (test plumbing)."""
    typestr = {torch.float32: "<f4", torch.uint8: "|u1"}[dtype]
    esize = 4 if dtype == torch.float32 else 1
    torch.cuda.synchronize(device)
    view = torch.as_tensor(_CudaPtr(ptr, (rows, cols), (pitch, esize), typestr), device=device)
    return view.clone()


def make_tracker_config(align_cfg, motion_model=capi.CONSTANT_VELOCITY, visratio_odo=0.9, visratio_integr=0.7,
                        image_filtering=capi.NO_FILTERS, delta_t=0.03333):
    cfg = capi.TrackerConfig()
    cfg.align = align_cfg
    cfg.align.mode = capi.MODE_TRACKER
    cfg.motion_model = motion_model
    cfg.visratio_odo, cfg.visratio_integr = visratio_odo, visratio_integr
    cfg.max_odo_kf_count = 9999999
    cfg.max_integr_kf_count = 9999999
    cfg.image_filtering = image_filtering
    cfg.delta_t = delta_t
    return cfg


class Tracker:
    """rgbid_tracker: VisodoTracker::trackNewFrame for `batch` lock-step RGB-D streams."""

    def __init__(self, ctx, cfg):
        self.ctx, self.cfg, self.lib = ctx, cfg, ctx.lib
        h = C.c_void_p()
        capi.check(self.lib.rgbid_tracker_create(ctx.h, C.byref(cfg), C.byref(h)), "tracker_create")
        self.h = h
        self.batch = cfg.align.batch
        self.results = (capi.FrameResult * self.batch)()
        ctx._register(self)

    def close(self):
        if self.h is not None:
            if self.ctx.h is not None:
                self.lib.rgbid_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        capi.check(self.lib.rgbid_tracker_reset(self.h), "tracker_reset")

    def track(self, depth, rgb):
        """depth: [B, rows, cols] uint16, rgb: [B, rows, cols, 3] uint8; torch CUDA tensors (device path) or
        torch CPU / numpy arrays (host path: the H2D copy happens inside the call)."""
        if isinstance(depth, torch.Tensor) and depth.is_cuda:
            self.ctx._enter()
            pd, pc, host = depth.data_ptr(), rgb.data_ptr(), 0
        elif isinstance(depth, torch.Tensor):
            pd, pc, host = depth.data_ptr(), rgb.data_ptr(), 1
        else:
            pd, pc, host = depth.ctypes.data, rgb.ctypes.data, 1
        capi.check(self.lib.rgbid_tracker_track(self.h, pd, pc, host, self.results), "tracker_track")
        return self.results

    def set_custom_calibration(self, rgb_intr, depth_intr, dist, dRc, t_dc):
        """Ingest every frame through the reference's custom-calibration path (prepareImagesCustomCalibration,
        src/visodo.cpp:775-823); dicts as in Context.undistort_*; None for rgb_intr switches it off."""
        if rgb_intr is None:
            capi.check(self.lib.rgbid_tracker_set_custom_calibration(self.h, None), "set_custom_calibration")
            return
        cal = capi.CustomCalibration()
        cal.rgb, cal.depth = Context._intr(rgb_intr), Context._intr(depth_intr)
        cal.dist = capi.DepthDist(float(dist["c1"]), float(dist["c0"]), (C.c_float * 9)(*dist["q0"]), (C.c_float * 9)(*dist["q1"]),
                                  int(dist["xshift"]), int(dist["yshift"]))
        cal.dRc = (C.c_float * 9)(*np.reshape(dRc, -1).tolist())
        cal.t_dc = (C.c_float * 3)(*np.reshape(t_dc, -1).tolist())
        capi.check(self.lib.rgbid_tracker_set_custom_calibration(self.h, C.byref(cal)), "set_custom_calibration")

    def set_keyframe_sink(self, fn):
        """fn(dict) is called inside track() for every outgoing integration keyframe (resetIntegrationKeyframe,
        src/visodo.cpp:1577-1672): indices, global pose, SEQ_KF constraint + covariance, and numpy copies of the
        overlap mask, colours, fused inverse depth and normals.  fn=None removes the sink."""
        if fn is None:
            self._sink = None
            capi.check(self.lib.rgbid_tracker_set_keyframe_sink(self.h, None, None), "set_keyframe_sink")
            return

        def trampoline(_user, kp):
            k = kp.contents
            rows, cols = k.rows, k.cols

            def arr(ptr, shape, dtype):
                n = int(np.prod(shape)) * np.dtype(dtype).itemsize
                return np.frombuffer(C.string_at(ptr, n), dtype=dtype).reshape(shape).copy()

            fn(dict(stream=k.stream, kf_index=k.kf_index, frame_index=k.frame_index, R=np.array(k.R[:]).reshape(3, 3),
                    t=np.array(k.t[:]), rel_R=np.array(k.rel_R[:]).reshape(3, 3), rel_t=np.array(k.rel_t[:]),
                    rel_cov=np.array(k.rel_cov[:]).reshape(6, 6),
                    overlap_mask=arr(k.overlap_mask, (rows, cols), np.uint8), colors=arr(k.colors, (rows, cols, 3), np.uint8),
                    depthinv=arr(k.depthinv, (rows, cols), np.float32), normals=arr(k.normals, (3 * rows, cols), np.float32)))

        self._sink = capi.KEYFRAME_SINK(trampoline)  # keep the callback object alive
        capi.check(self.lib.rgbid_tracker_set_keyframe_sink(self.h, C.cast(self._sink, C.c_void_p), None), "set_keyframe_sink")

    def prefetch(self, depth, rgb):
        """Start uploading the NEXT frame (CPU tensors / numpy arrays, ideally pinned) while the current one is tracked;
        the following track(depth, rgb) with the same buffers uses the uploaded copy."""
        if isinstance(depth, torch.Tensor):
            assert not depth.is_cuda
            pd, pc = depth.data_ptr(), rgb.data_ptr()
        else:
            pd, pc = depth.ctypes.data, rgb.ctypes.data
        capi.check(self.lib.rgbid_tracker_prefetch(self.h, pd, pc), "tracker_prefetch")

    @property
    def aligner_handle(self):
        return C.c_void_p(self.lib.rgbid_tracker_aligner(self.h))

    def time_build(self, level=0, reps=20):
        ms = _F(0)
        capi.check(self.lib.rgbid_aligner_time_build(self.aligner_handle, level, reps, C.byref(ms)), "aligner_time_build")
        return ms.value

    def export_systems(self, out):
        """[batch, 48] float64 CUDA tensor <- (cov 36, R 9, t 3) of every stream's last alignment, written on the
        context's stream by a kernel (no host copy): the payload of the multi-GPU all-gather."""
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == self.batch * 48
        capi.check(self.lib.rgbid_aligner_export_systems(self.aligner_handle, out.data_ptr()), "export_systems")
        return out

    def overlap_mask(self, index=0):
        """Overlap mask of the integration keyframe of stream `index` (uint8, 1 = seen by the previous keyframe)."""
        p, pitch = C.c_void_p(), C.c_size_t()
        capi.check(self.lib.rgbid_tracker_overlap_mask(self.h, index, C.byref(p), C.byref(pitch)), "overlap_mask")
        return _wrap_device(p.value, self.cfg.align.rows, self.cfg.align.cols, pitch.value, self.ctx.device, dtype=torch.uint8)

    def keyframe_map(self, which, index=0):
        p, pitch = C.c_void_p(), C.c_size_t()
        capi.check(self.lib.rgbid_tracker_keyframe_map(self.h, which, index, C.byref(p), C.byref(pitch)), "keyframe_map")
        rows, cols = self.cfg.align.rows, self.cfg.align.cols
        if which in (3, 4):
            rows *= 3
        return _wrap_device(p.value, rows, cols, pitch.value, self.ctx.device)
