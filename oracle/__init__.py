"""CPU oracle (numpy-facing ctypes binding of oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this package.
The product (rgbid-slam_b200/) never does.  oracle.c restates the reference's algorithm function by function
(file:line citations there); oracle/_ref/libref_oracle.so (see ref.py) is the reference's own device layer
compiled verbatim and is what pins this restatement (tests/golden/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libref_oracle.so")
REFERENCE_ROOT = "/root/reference"

LSQ, HUBER, TUKEY, STUDENT = 0, 1, 2, 3
SIGMA_MAD, SIGMA_PDF, SIGMA_CONS = 0, 1, 2
INDEPENDENT, MIN_WEIGHT, GEOM_ONLY, PHOT_ONLY = 0, 1, 2, 3
MODE_TRACKER, MODE_ALIGN = 0, 1
TEX_FRAC_ROUND, TEX_FRAC_TRUNC, TEX_FRAC_EXACT = 0, 1, 2
MAX_LEVELS = 8

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)


class SystemParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("mestimator", C.c_int), ("weighting", C.c_int), ("student_nu", C.c_int),
                ("sigma_depthinv", C.c_float), ("sigma_int", C.c_float), ("bias_depthinv", C.c_float),
                ("bias_int", C.c_float), ("nu_depthinv", C.c_float), ("nu_int", C.c_float)]


class AlignConfig(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("levels", C.c_int), ("finest_level", C.c_int),
                ("iterations", C.c_int * MAX_LEVELS), ("mode", C.c_int), ("mestimator", C.c_int),
                ("weighting", C.c_int), ("sigma_estimator", C.c_int), ("nsamples", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("warp_first", C.c_int),
                ("termination", C.c_int), ("conv_eps", C.c_float)]


class Pyramids(C.Structure):
    _fields_ = [(n, C.c_void_p * MAX_LEVELS) for n in
                ("W_kf", "I_kf", "gWx_kf", "gWy_kf", "gIx_kf", "gIy_kf", "gWx_cov", "gWy_cov", "gIx_cov", "gIy_cov",
                 "W_cur", "I_cur")]


class IterTrace(C.Structure):
    _fields_ = [("level", C.c_int), ("iter", C.c_int), ("sums27", C.c_double * 27),
                ("sigma_int", C.c_float), ("sigma_depthinv", C.c_float), ("bias_int", C.c_float),
                ("bias_depthinv", C.c_float), ("nu_int", C.c_float), ("nu_depthinv", C.c_float),
                ("irls_iters_int", C.c_int), ("irls_iters_depthinv", C.c_int),
                ("x", C.c_double * 6), ("R", C.c_double * 9), ("t", C.c_double * 3)]


class FrameStats(C.Structure):
    _fields_ = [("cov_sums27", C.c_double * 27), ("chi_square", C.c_float), ("chi_test", C.c_float),
                ("ndof", C.c_float)]


def build(with_ref=None):
    """Compile liboracle.so (gcc) and, when the reference sources are present, oracle/_ref/libref_oracle.so."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    if with_ref is None:
        with_ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "cuda"))
    if with_ref:
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref", "REF=" + REFERENCE_ROOT], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_ref=False)
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_digamma.restype = C.c_double
        _lib.orc_digamma.argtypes = [C.c_double]
        _lib.orc_visibility_ratio.restype = C.c_float
    return _lib


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_tex_frac_mode(mode):
    lib().orc_set_tex_frac_mode(C.c_int(mode))


def depth_to_invdepth(depth_u16, factor_depth=1.0):
    d = np.ascontiguousarray(depth_u16, dtype=np.uint16)
    out = np.empty(d.shape, dtype=np.float32)
    lib().orc_depth_to_invdepth(_p(d), _p(out), d.shape[0], d.shape[1], C.c_float(factor_depth))
    return out


def intensity(rgb_u8):
    c = np.ascontiguousarray(rgb_u8, dtype=np.uint8)
    out = np.empty(c.shape[:2], dtype=np.float32)
    lib().orc_intensity(_p(c), _p(out), c.shape[0], c.shape[1])
    return out


def pyr_down(src):
    s = f32(src)
    out = np.empty((s.shape[0] // 2, s.shape[1] // 2), dtype=np.float32)
    lib().orc_pyr_down(_p(s), s.shape[0], s.shape[1], _p(out))
    return out


def gradient(src):
    s = f32(src)
    gx, gy = np.empty_like(s), np.empty_like(s)
    lib().orc_gradient(_p(s), s.shape[0], s.shape[1], _p(gx), _p(gy))
    return gx, gy


def bilateral(src, sigma):
    s = f32(src)
    out = np.empty_like(s)
    lib().orc_bilateral(_p(s), s.shape[0], s.shape[1], _p(out), C.c_float(sigma))
    return out


def _warp(fn, src, prev, Rp, tp):
    s, p = f32(src), f32(prev)
    Rp, tp = f32(np.reshape(Rp, 9)), f32(np.reshape(tp, 3))
    out = np.empty_like(p)
    fn(_p(s), _p(p), _p(out), p.shape[0], p.shape[1], _p(Rp), _p(tp))
    return out


def warp_invdepth(src, prev, Rp, tp):
    return _warp(lib().orc_warp_invdepth, src, prev, Rp, tp)


def warp_intensity(src, prev, Rp, tp):
    return _warp(lib().orc_warp_intensity, src, prev, Rp, tp)


def warp_invdepth_weighted(src, prev, weight_inout, Rp, tp):
    s, p = f32(src), f32(prev)
    Rp, tp = f32(np.reshape(Rp, 9)), f32(np.reshape(tp, 3))
    assert weight_inout.dtype == np.float32 and weight_inout.flags.c_contiguous
    out = np.empty_like(p)
    lib().orc_warp_invdepth_weighted(_p(s), _p(p), _p(out), _p(weight_inout), p.shape[0], p.shape[1], _p(Rp), _p(tp))
    return out


def integrate_warped_frame(wsrc, wweight, dst_inout, dweight_inout):
    a, b = f32(wsrc), f32(wweight)
    assert dst_inout.dtype == np.float32 and dweight_inout.dtype == np.float32
    lib().orc_integrate_warped_frame(_p(a), _p(b), _p(dst_inout), _p(dweight_inout), a.shape[0], a.shape[1])


def visibility_ratio(depth_src, depth_dst, Rp, tp, with_mask=False):
    a, b = f32(depth_src), f32(depth_dst)
    Rp, tp = f32(np.reshape(Rp, 9)), f32(np.reshape(tp, 3))
    mask = np.zeros(a.shape, dtype=np.uint8) if with_mask else None
    nv, nn = C.c_double(), C.c_double()
    r = lib().orc_visibility_ratio(_p(a), _p(b), a.shape[0], a.shape[1], _p(Rp), _p(tp),
                                   _p(mask) if with_mask else None, C.byref(nv), C.byref(nn))
    return (r, mask) if with_mask else r


def error_geometry(rows, cols, nsamples):
    kr, kc, s = C.c_int(), C.c_int(), C.c_int()
    lib().orc_error_geometry(rows, cols, nsamples, C.byref(kr), C.byref(kc), C.byref(s))
    return kr.value, kc.value, s.value


def compute_error(im1, im0, nsamples=9999999):
    a, b = f32(im1), f32(im0)
    kr, kc, _ = error_geometry(a.shape[0], a.shape[1], nsamples)
    err = np.empty(kr * kc, dtype=np.float32)
    lib().orc_compute_error(_p(a), _p(b), a.shape[0], a.shape[1], nsamples, _p(err))
    return err


def digamma(x):
    return lib().orc_digamma(float(x))


def sigma_nu_student(err, bias, sigma, mest=STUDENT):
    e = f32(err)
    b, s, nu = C.c_float(bias), C.c_float(sigma), C.c_float(0)
    iters = lib().orc_sigma_nu_student(_p(e), e.size, C.byref(b), C.byref(s), C.byref(nu), mest)
    return b.value, s.value, nu.value, iters


def nu_student(err, bias, sigma):
    e = f32(err)
    nu = C.c_float(0)
    lib().orc_nu_student(_p(e), e.size, C.c_float(bias), C.c_float(sigma), C.byref(nu))
    return nu.value


def sigma_pdf(err, bias, sigma, mest):
    e = f32(err)
    b, s = C.c_float(bias), C.c_float(sigma)
    lib().orc_sigma_pdf(_p(e), e.size, C.byref(b), C.byref(s), mest)
    return b.value, s.value


def chi_square(err_int, err_depth, sigma_int, sigma_depth, mest):
    a, b = f32(err_int), f32(err_depth)
    x, y, z = C.c_float(), C.c_float(), C.c_float()
    lib().orc_chi_square(_p(a), _p(b), a.size, C.c_float(sigma_int), C.c_float(sigma_depth), mest, C.byref(x),
                         C.byref(y), C.byref(z))
    return x.value, y.value, z.value


def system_params(fx, fy, cx, cy, mestimator=STUDENT, weighting=INDEPENDENT, student_nu=1, sigma_depthinv=0.0025,
                  sigma_int=5.0, bias_depthinv=0.0, bias_int=0.0, nu_depthinv=5.0, nu_int=5.0):
    return SystemParams(fx, fy, cx, cy, mestimator, weighting, student_nu, sigma_depthinv, sigma_int, bias_depthinv,
                        bias_int, nu_depthinv, nu_int)


def build_system(W0, I0, gWx, gWy, gIx, gIy, W1, I1, params):
    maps = [f32(m) for m in (W0, I0, gWx, gWy, gIx, gIy, W1, I1)]
    rows, cols = maps[0].shape
    sums, A, b = np.zeros(27), np.zeros(36), np.zeros(6)
    lib().orc_build_system(*[_p(m) for m in maps], rows, cols, C.byref(params), _p(sums), _p(A), _p(b))
    return A.reshape(6, 6), b, sums


def vmap(depth_inv, fx, fy, cx, cy):
    d = f32(depth_inv)
    out = np.full((3 * d.shape[0], d.shape[1]), np.nan, dtype=np.float32)
    lib().orc_vmap(_p(d), d.shape[0], d.shape[1], C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(out))
    return out


def nmap_gradients(depth_inv, gx, gy, fx, fy, cx, cy):
    d, a, b = f32(depth_inv), f32(gx), f32(gy)
    out = np.full((3 * d.shape[0], d.shape[1]), np.nan, dtype=np.float32)
    lib().orc_nmap_gradients(_p(d), _p(a), _p(b), d.shape[0], d.shape[1], C.c_float(fx), C.c_float(fy), C.c_float(cx),
                             C.c_float(cy), _p(out))
    return out


# ---- host algebra -------------------------------------------------------------------------------------
def _d(a, n):
    return np.ascontiguousarray(np.reshape(a, n), dtype=np.float64)


def exp_map_rot(omega):
    R = np.zeros(9)
    lib().orc_exp_map_rot(_p(_d(omega, 3)), _p(R))
    return R.reshape(3, 3)


def exp_map(omega, v):
    R, t = np.zeros(9), np.zeros(3)
    lib().orc_exp_map(_p(_d(omega, 3)), _p(_d(v, 3)), _p(R), _p(t))
    return R.reshape(3, 3), t


def log_map(R, t):
    tw = np.zeros(6)
    lib().orc_log_map(_p(_d(R, 9)), _p(_d(t, 3)), _p(tw))
    return tw


def force_orthogonal(M):
    R = np.zeros(9)
    lib().orc_force_orthogonal(_p(_d(M, 9)), _p(R))
    return R.reshape(3, 3)


def llt_solve6(A, b):
    x = np.zeros(6)
    lib().orc_llt_solve6(_p(_d(A, 36)), _p(_d(b, 6)), _p(x))
    return x


def inverse6(A):
    Ai = np.zeros(36)
    lib().orc_inverse6(_p(_d(A, 36)), _p(Ai))
    return Ai.reshape(6, 6)


def projective_pose(R, t, fx, fy, cx, cy, inverse=False):
    Rp, tp = np.zeros(9, dtype=np.float32), np.zeros(3, dtype=np.float32)
    fn = lib().orc_projective_inverse_pose if inverse else lib().orc_projective_pose
    fn(_p(_d(R, 9)), _p(_d(t, 3)), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(Rp), _p(tp))
    return Rp.reshape(3, 3), tp


def gn_update(A, b, R, t):
    R, t, x = _d(R, 9).copy(), _d(t, 3).copy(), np.zeros(6)
    bad = lib().orc_gn_update(_p(_d(A, 36)), _p(_d(b, 6)), _p(R), _p(t), _p(x))
    return R.reshape(3, 3), t, x, bad


# ---- pyramids + drivers -----------------------------------------------------------------------------------
def build_pyramid(img, levels):
    out = [f32(img)]
    for _ in range(1, levels):
        out.append(pyr_down(out[-1]))
    return out


def prepare_keyframe(W0, I0, levels, tracker=True):
    """saveCurrentImagesAsOdoKeyframes (src/visodo.cpp:826-878) / keyframe_align.cpp:157-176."""
    kf = dict(W=build_pyramid(W0, levels), I=build_pyramid(I0, levels))
    g = [gradient(m) for m in kf["W"]]
    kf["gWx"], kf["gWy"] = [a for a, _ in g], [b for _, b in g]
    g = [gradient(m) for m in kf["I"]]
    kf["gIx"], kf["gIy"] = [a for a, _ in g], [b for _, b in g]
    if tracker:
        Wf = build_pyramid(bilateral(kf["W"][0], 2.0 * 0.0025), levels)
        If = build_pyramid(bilateral(kf["I"][0], 3.0), levels)
        g = [gradient(m) for m in Wf]
        kf["cgWx"], kf["cgWy"] = [a for a, _ in g], [b for _, b in g]
        g = [gradient(m) for m in If]
        kf["cgIx"], kf["cgIy"] = [a for a, _ in g], [b for _, b in g]
    return kf


def prepare_current(W, I, levels):
    return dict(W=build_pyramid(W, levels), I=build_pyramid(I, levels))


def make_config(rows, cols, levels, mode, iterations, fx, fy, cx, cy, finest_level=0, mestimator=STUDENT,
                weighting=INDEPENDENT, sigma_estimator=SIGMA_PDF, nsamples=None, warp_first=0,
                termination=0, conv_eps=0.0):
    c = AlignConfig()
    c.rows, c.cols, c.levels, c.finest_level = rows, cols, levels, finest_level
    for i in range(MAX_LEVELS):
        c.iterations[i] = iterations[i] if i < len(iterations) else 0
    c.mode, c.mestimator, c.weighting, c.sigma_estimator = mode, mestimator, weighting, sigma_estimator
    c.nsamples = nsamples if nsamples is not None else (10000 if mode == MODE_TRACKER else 19200)
    c.fx, c.fy, c.cx, c.cy = fx, fy, cx, cy
    c.warp_first = int(warp_first)
    c.termination, c.conv_eps = int(termination), float(conv_eps)  # ORC_TERM_*: 0 all iterations, 1 CHI_SQUARED, 2 convergence
    return c


def _fill_pyramids(kf, cur, levels, ptr_of):
    P = Pyramids()
    names = dict(W_kf=kf["W"], I_kf=kf["I"], gWx_kf=kf["gWx"], gWy_kf=kf["gWy"], gIx_kf=kf["gIx"], gIy_kf=kf["gIy"],
                 W_cur=cur["W"], I_cur=cur["I"])
    if "cgWx" in kf:
        names.update(gWx_cov=kf["cgWx"], gWy_cov=kf["cgWy"], gIx_cov=kf["cgIx"], gIy_cov=kf["cgIy"])
    for n, lst in names.items():
        arr = getattr(P, n)
        for l in range(levels):
            arr[l] = ptr_of(lst[l])
    return P


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


def align(cfg, kf, cur, R=None, t=None, fn=None, ptr_of=None):
    """Coarse-to-fine driver (orc_align).  fn / ptr_of let ref.py reuse this glue for the reference library."""
    R = np.eye(3) if R is None else np.array(R, dtype=np.float64)
    t = np.zeros(3) if t is None else np.array(t, dtype=np.float64)
    R = np.ascontiguousarray(R).reshape(9).copy()
    t = np.ascontiguousarray(t).reshape(3).copy()
    cov = np.zeros(36)
    niters = sum(cfg.iterations[l] for l in range(cfg.finest_level, cfg.levels))
    trace = (IterTrace * (niters + 1))()
    nt = C.c_int()
    stats = FrameStats()
    P = _fill_pyramids(kf, cur, cfg.levels, ptr_of or (lambda a: a.ctypes.data))
    status = (fn or lib().orc_align)(C.byref(cfg), C.byref(P), _p(R), _p(t), _p(cov), trace, niters + 1, C.byref(nt),
                                      C.byref(stats))
    return dict(R=R.reshape(3, 3), t=t, cov=cov.reshape(6, 6), status=status, trace=trace_to_dicts(trace, nt.value),
                cov_sums27=np.array(stats.cov_sums27[:]), chi_square=stats.chi_square, chi_test=stats.chi_test,
                ndof=stats.ndof)
