"""ctypes binding of include/rgbid_b200.h (the C ABI of librgbid_b200.so).

This is the reference-side binding a maintainer would write for a Python caller; the tests and bench.py go
through it so that every measured or checked call crosses the same boundary a C++ caller would.
There is NO fallback: if the CUDA library is missing, loading fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RGBID_LIB selects another build of the same library (a kernel variant built with RGBID_BUILD_TAG, see build.py)
LIB_PATH = os.environ.get("RGBID_LIB") or os.path.join(_HERE, "lib", "librgbid_b200.so")

MAX_LEVELS = 8
OK, ERR_NAN, ERR_ARG, ERR_NOMEM, ERR_STATE, ERR_TIMEOUT = 0, -1, -2, -3, -4, -5
LSQ, HUBER, TUKEY, STUDENT = 0, 1, 2, 3
NO_MM, CONSTANT_VELOCITY = 0, 1
SIGMA_MAD, SIGMA_PDF, SIGMA_CONS = 0, 1, 2
INDEPENDENT, MIN_WEIGHT, GEOM_ONLY, PHOT_ONLY = 0, 1, 2, 3
NO_FILTERS, FILTER_GRADS = 0, 1
MODE_TRACKER, MODE_ALIGN = 0, 1
TERM_ALL_ITERS, TERM_CHI_SQUARED, TERM_CONVERGENCE = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class SystemParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("mestimator", C.c_int), ("weighting", C.c_int), ("student_nu", C.c_int),
                ("sigma_depthinv", C.c_float), ("sigma_int", C.c_float), ("bias_depthinv", C.c_float),
                ("bias_int", C.c_float), ("nu_depthinv", C.c_float), ("nu_int", C.c_float)]


class Intr(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "k1", "k2", "k3", "k4", "k5")]


class DepthDist(C.Structure):
    _fields_ = [("c1", C.c_float), ("c0", C.c_float), ("q0", C.c_float * 9), ("q1", C.c_float * 9), ("xshift", C.c_int),
                ("yshift", C.c_int)]


class CustomCalibration(C.Structure):
    _fields_ = [("rgb", Intr), ("depth", Intr), ("dist", DepthDist), ("dRc", C.c_float * 9), ("t_dc", C.c_float * 3)]


class AlignConfig(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("levels", C.c_int), ("finest_level", C.c_int),
                ("iterations", C.c_int * MAX_LEVELS), ("batch", C.c_int), ("mode", C.c_int),
                ("mestimator", C.c_int), ("weighting", C.c_int), ("sigma_estimator", C.c_int),
                ("nsamples", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("factor_depth", C.c_float), ("with_fusion", C.c_int), ("warp_first", C.c_int),
                ("termination", C.c_int), ("conv_eps", C.c_float)]


class IterTrace(C.Structure):
    _fields_ = [("level", C.c_int), ("iter", C.c_int), ("sums27", C.c_double * 27),
                ("sigma_int", C.c_float), ("sigma_depthinv", C.c_float), ("bias_int", C.c_float),
                ("bias_depthinv", C.c_float), ("nu_int", C.c_float), ("nu_depthinv", C.c_float),
                ("irls_iters_int", C.c_int), ("irls_iters_depthinv", C.c_int),
                ("x", C.c_double * 6), ("R", C.c_double * 9), ("t", C.c_double * 3)]


class TrackerConfig(C.Structure):
    _fields_ = [("align", AlignConfig), ("motion_model", C.c_int), ("visratio_odo", C.c_float),
                ("visratio_integr", C.c_float), ("max_odo_kf_count", C.c_int),
                ("max_integr_kf_count", C.c_int), ("image_filtering", C.c_int), ("delta_t", C.c_float)]


class FrameResult(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("dR", C.c_double * 9), ("dt", C.c_double * 3),
                ("cov", C.c_double * 36), ("visibility_odo", C.c_float), ("visibility_integr", C.c_float),
                ("chi_square", C.c_float), ("chi_test", C.c_float), ("ndof", C.c_float), ("status", C.c_int),
                ("new_odo_keyframe", C.c_int), ("new_integr_keyframe", C.c_int), ("frame_index", C.c_int),
                ("seq_R", C.c_double * 9), ("seq_t", C.c_double * 3), ("seq_cov", C.c_double * 36),
                ("lost_again", C.c_int)]


class KeyframeHandoff(C.Structure):
    _fields_ = [("stream", C.c_int), ("kf_index", C.c_int), ("frame_index", C.c_int), ("rows", C.c_int), ("cols", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("R", C.c_double * 9), ("t", C.c_double * 3), ("rel_R", C.c_double * 9), ("rel_t", C.c_double * 3),
                ("rel_cov", C.c_double * 36),
                ("overlap_mask", C.c_void_p), ("overlap_mask_pitch", C.c_size_t), ("colors", C.c_void_p),
                ("depthinv", C.c_void_p), ("depthinv_pitch", C.c_size_t), ("normals", C.c_void_p),
                ("normals_pitch", C.c_size_t)]


KEYFRAME_SINK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(KeyframeHandoff))


P, SZ, I, F = C.c_void_p, C.c_size_t, C.c_int, C.c_float

# name -> (restype, argtypes); mirrors include/rgbid_b200.h one to one
PROTOTYPES = {
    "rgbid_ctx_create": (I, [C.POINTER(P), I, P]),
    "rgbid_ctx_destroy": (I, [P]),
    "rgbid_ctx_sync": (I, [P]),
    "rgbid_ctx_stream": (P, [P]),
    "rgbid_version": (I, []),
    "rgbid_status_string": (C.c_char_p, [I]),
    "rgbid_ctx_launch_count": (C.c_longlong, [P]),
    "rgbid_convert_depth_to_invdepth": (I, [P, P, SZ, P, SZ, I, I, F]),
    "rgbid_compute_intensity": (I, [P, P, SZ, P, SZ, I, I]),
    "rgbid_decompose_rgb": (I, [P, P, SZ, P, P, P, SZ, I, I]),
    "rgbid_pyr_down": (I, [P, P, SZ, I, I, P, SZ]),
    "rgbid_compute_gradient": (I, [P, P, SZ, I, I, P, P, SZ]),
    "rgbid_bilateral_filter": (I, [P, P, SZ, I, I, P, SZ, F]),
    "rgbid_copy_image": (I, [P, P, SZ, P, SZ, I, I]),
    "rgbid_fill_image": (I, [P, P, SZ, I, I, F]),
    "rgbid_create_vmap": (I, [P, P, SZ, I, I, F, F, F, F, P, SZ]),
    "rgbid_create_nmap_gradients": (I, [P, P, P, P, SZ, I, I, F, F, F, F, P, SZ]),
    "rgbid_warp_invdepth": (I, [P, P, SZ, P, SZ, P, SZ, I, I, c_float_p, c_float_p]),
    "rgbid_warp_intensity": (I, [P, P, SZ, P, SZ, P, SZ, I, I, c_float_p, c_float_p]),
    "rgbid_warp_invdepth_weighted": (I, [P, P, SZ, P, SZ, P, SZ, P, SZ, I, I, c_float_p, c_float_p]),
    "rgbid_integrate_warped_frame": (I, [P, P, SZ, P, SZ, P, SZ, P, SZ, I, I]),
    "rgbid_visibility_ratio": (I, [P, P, SZ, P, SZ, I, I, c_float_p, c_float_p, P, SZ, c_float_p]),
    "rgbid_undistort_intensity": (I, [P, P, SZ, P, SZ, I, I, C.POINTER(Intr)]),
    "rgbid_undistort_depthinv": (I, [P, P, SZ, P, SZ, I, I, C.POINTER(Intr), C.POINTER(DepthDist)]),
    "rgbid_register_depthinv": (I, [P, P, SZ, P, SZ, I, I, c_float_p, c_float_p, c_float_p]),
    "rgbid_integrate_warped_rgb": (I, [P, P, P, P, P, P, P, P, SZ, P, SZ, I, I]),
    "rgbid_generate_image": (I, [P, P, P, SZ, P, SZ, c_float_p, P, SZ, I, I]),
    "rgbid_error_geometry": (I, [I, I, I, c_int_p, c_int_p, c_int_p]),
    "rgbid_compute_error": (I, [P, P, SZ, P, SZ, I, I, I, P, c_int_p]),
    "rgbid_sigma_nu_student": (I, [P, P, I, c_float_p, c_float_p, c_float_p, I]),
    "rgbid_nu_student": (I, [P, P, I, F, F, c_float_p]),
    "rgbid_sigma_pdf": (I, [P, P, I, c_float_p, c_float_p, I]),
    "rgbid_chi_square": (I, [P, P, P, I, F, F, I, c_float_p, c_float_p, c_float_p]),
    "rgbid_build_system": (I, [P, P, P, P, P, P, P, P, P, SZ, I, I, C.POINTER(SystemParams), c_double_p, c_double_p]),
    "rgbid_build_system_pitched": (I, [P, P, P, P, P, P, P, P, P, C.POINTER(C.c_size_t), I, I, C.POINTER(SystemParams),
                                       c_double_p, c_double_p]),
    "rgbid_aligner_create": (I, [P, C.POINTER(AlignConfig), C.POINTER(P)]),
    "rgbid_aligner_destroy": (I, [P]),
    "rgbid_aligner_num_iterations": (I, [P]),
    "rgbid_aligner_iterations_done": (I, [P, c_int_p]),
    "rgbid_aligner_set_keyframe": (I, [P, I, P, SZ, P, SZ, I]),
    "rgbid_aligner_set_current": (I, [P, I, P, SZ, P, SZ, I]),
    "rgbid_aligner_set_current_rgbd": (I, [P, I, P, SZ, P, SZ, I]),
    "rgbid_aligner_current_to_keyframe": (I, [P, I]),
    "rgbid_aligner_set_trace": (I, [P, I]),
    "rgbid_aligner_run": (I, [P, c_double_p, c_double_p, c_double_p, c_int_p, C.POINTER(IterTrace)]),
    "rgbid_aligner_enqueue": (I, [P, c_double_p, c_double_p]),
    "rgbid_aligner_fetch": (I, [P, c_double_p, c_double_p, c_double_p, c_int_p, C.POINTER(IterTrace)]),
    "rgbid_aligner_frame_stats": (I, [P, c_float_p]),
    "rgbid_aligner_export_systems": (I, [P, P]),
    "rgbid_aligner_state_bytes": (SZ, []),
    "rgbid_aligner_time_build": (I, [P, I, I, c_float_p]),
    "rgbid_aligner_time_scale": (I, [P, I, I, c_float_p]),
    "rgbid_aligner_map": (I, [P, I, I, I, C.POINTER(P), C.POINTER(SZ)]),
    "rgbid_tracker_create": (I, [P, C.POINTER(TrackerConfig), C.POINTER(P)]),
    "rgbid_tracker_destroy": (I, [P]),
    "rgbid_tracker_reset": (I, [P]),
    "rgbid_tracker_track": (I, [P, P, P, I, C.POINTER(FrameResult)]),
    "rgbid_tracker_prefetch": (I, [P, P, P]),
    "rgbid_tracker_set_keyframe_sink": (I, [P, P, P]),
    "rgbid_tracker_set_custom_calibration": (I, [P, C.POINTER(CustomCalibration)]),
    "rgbid_tracker_track_device": (I, [P, P, SZ, SZ, P, SZ, SZ, C.POINTER(FrameResult)]),
    "rgbid_tracker_keyframe_map": (I, [P, I, I, C.POINTER(P), C.POINTER(SZ)]),
    "rgbid_tracker_overlap_mask": (I, [P, I, C.POINTER(P), C.POINTER(SZ)]),
    "rgbid_tracker_aligner": (P, [P]),
}

_lib = None


class RgbidError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = load().rgbid_status_string(status).decode() if _lib is not None else str(status)
        super().__init__("%s failed: status %d (%s)" % (where, status, msg))


def load():
    """Load librgbid_b200.so and bind every symbol the header declares.  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "librgbid_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, where=""):
    if status != OK:
        raise RgbidError(status, where)
