"""Pins the CPU oracle against outputs of the reference's own CUDA kernels (tests/golden/ref_golden_r1.npz)."""
import numpy as np

import golden_checks as gc
import oracle as orc
from oracle.tracker import OracleTracker


class OracleBackend:
    depth_to_invdepth = staticmethod(orc.depth_to_invdepth)
    intensity = staticmethod(orc.intensity)
    pyr_down = staticmethod(orc.pyr_down)
    gradient = staticmethod(orc.gradient)
    bilateral = staticmethod(orc.bilateral)
    warp_invdepth = staticmethod(orc.warp_invdepth)
    warp_intensity = staticmethod(orc.warp_intensity)
    warp_invdepth_weighted = staticmethod(orc.warp_invdepth_weighted)
    compute_error = staticmethod(orc.compute_error)
    sigma_nu_student = staticmethod(orc.sigma_nu_student)
    nu_student = staticmethod(orc.nu_student)
    sigma_pdf = staticmethod(orc.sigma_pdf)
    chi_square = staticmethod(orc.chi_square)

    @staticmethod
    def vmap(W, fx, fy, cx, cy):
        return orc.vmap(W, fx, fy, cx, cy)

    @staticmethod
    def nmap_gradients(W, gx, gy, fx, fy, cx, cy):
        return orc.nmap_gradients(W, gx, gy, fx, fy, cx, cy)

    @staticmethod
    def integrate_warped_frame(ws, ww, kf, kfw):
        orc.integrate_warped_frame(ws, ww, kf, kfw)
        return kf, kfw

    @staticmethod
    def visibility_ratio(src, dst, Rp, tp, with_mask):
        return orc.visibility_ratio(src, dst, Rp, tp, with_mask=with_mask)

    @staticmethod
    def build_system(W0, I0, gWx, gWy, gIx, gIy, W1, I1, kw):
        p = orc.system_params(kw.pop("fx"), kw.pop("fy"), kw.pop("cx"), kw.pop("cy"), **kw)
        return orc.build_system(W0, I0, gWx, gWy, gIx, gIy, W1, I1, p)[2]

    @staticmethod
    def align(WA, IA, depth_b, rgb_b, mode, its, intr, nsamples):
        WB, IB = orc.depth_to_invdepth(depth_b), orc.intensity(rgb_b)
        cfg = orc.make_config(gc.ROWS, gc.COLS, gc.LEVELS, mode, its, *intr, nsamples=nsamples)
        out = orc.align(cfg, orc.prepare_keyframe(WA, IA, gc.LEVELS, mode == orc.MODE_TRACKER),
                        orc.prepare_current(WB, IB, gc.LEVELS))
        tr = out["trace"]
        return dict(R=out["R"], t=out["t"], cov=out["cov"], sums27=[t["sums27"] for t in tr],
                    scale=[np.array([t["sigma_int"], t["sigma_depthinv"], t["bias_int"], t["bias_depthinv"], t["nu_int"],
                                     t["nu_depthinv"]], dtype=np.float32) for t in tr],
                    trace_t=[t["t"] for t in tr], cov_sums27=out["cov_sums27"],
                    chi=(out["chi_square"], out["chi_test"], out["ndof"]))

    @staticmethod
    def track_sequence(depth, rgb, intr, nsamples):
        i = dict(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3])
        ot = OracleTracker(gc.ROWS, gc.COLS, i, levels=gc.LEVELS, iterations=(10, 5, 3), kind="cpu", nsamples=nsamples)
        poses, flags = [], []
        for k in range(depth.shape[0]):
            o = ot.track(depth[k], rgb[k])
            poses.append(np.concatenate([o["R"].reshape(9), o["t"]]))
            flags.append([o["new_odo_keyframe"], o["new_integr_keyframe"], o["status"]])
        return poses, flags, ot.intW


def test_oracle_image_ops_match_reference_kernels():
    gc.check_image_ops(OracleBackend, gc.load())


def test_oracle_warps_fusion_visibility_match_reference_kernels():
    gc.check_warps(OracleBackend, gc.load())


def test_oracle_scale_estimation_matches_reference_kernels():
    gc.check_scale(OracleBackend, gc.load())


def test_oracle_normal_equations_match_reference_kernels():
    gc.check_systems(OracleBackend, gc.load())


def test_oracle_alignment_matches_reference_pipeline():
    gc.check_align(OracleBackend, gc.load())


def test_oracle_tracker_sequence_matches_reference_pipeline():
    gc.check_sequence(OracleBackend, gc.load())
