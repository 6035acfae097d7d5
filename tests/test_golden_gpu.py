"""The new CUDA path (through the C ABI) against the committed outputs of the reference's own CUDA kernels."""
import numpy as np
import pytest
import torch

import golden_checks as gc
from rgbid_slam_b200 import capi, host

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


class CudaBackend:
    def __init__(self, ctx):
        self.c = ctx

    def depth_to_invdepth(self, d, f):
        return npy(self.c.convert_depth_to_invdepth(cu(d), f))

    def intensity(self, c):
        return npy(self.c.compute_intensity(cu(c)))

    def pyr_down(self, a):
        return npy(self.c.pyr_down(cu(a)))

    def gradient(self, a):
        gx, gy = self.c.compute_gradient(cu(a))
        return npy(gx), npy(gy)

    def bilateral(self, a, s):
        return npy(self.c.bilateral_filter(cu(a), s))

    def vmap(self, W, fx, fy, cx, cy):
        return npy(self.c.create_vmap(cu(W), fx, fy, cx, cy))

    def nmap_gradients(self, W, gx, gy, fx, fy, cx, cy):
        return npy(self.c.create_nmap_gradients(cu(W), cu(gx), cu(gy), fx, fy, cx, cy))

    def warp_invdepth(self, src, prev, Rp, tp):
        return npy(self.c.warp_invdepth(cu(src), cu(prev), Rp, tp))

    def warp_intensity(self, src, prev, Rp, tp):
        return npy(self.c.warp_intensity(cu(src), cu(prev), Rp, tp))

    def warp_invdepth_weighted(self, src, prev, w_inout, Rp, tp):
        wg = cu(w_inout)
        out = npy(self.c.warp_invdepth_weighted(cu(src), cu(prev), wg, Rp, tp))
        w_inout[...] = npy(wg)
        return out

    def integrate_warped_frame(self, ws, ww, kf, kfw):
        a, b = cu(kf), cu(kfw)
        self.c.integrate_warped_frame(cu(ws), cu(ww), a, b)
        return npy(a), npy(b)

    def visibility_ratio(self, src, dst, Rp, tp, with_mask):
        if with_mask:
            r, m = self.c.visibility_ratio(cu(src), cu(dst), Rp, tp, with_mask=True)
            return r, npy(m)
        return self.c.visibility_ratio(cu(src), cu(dst), Rp, tp)

    def compute_error(self, a, b, n):
        return npy(self.c.compute_error(cu(a), cu(b), n))

    def sigma_nu_student(self, e, b, s):
        return self.c.sigma_nu_student(cu(e), b, s)

    def nu_student(self, e, b, s):
        return self.c.nu_student(cu(e), b, s)

    def sigma_pdf(self, e, b, s, m):
        return self.c.sigma_pdf(cu(e), b, s, m)

    def chi_square(self, a, b, si, sd, m):
        return self.c.chi_square(cu(a), cu(b), si, sd, m)

    def build_system(self, W0, I0, gWx, gWy, gIx, gIy, W1, I1, kw):
        p = capi.SystemParams(kw["fx"], kw["fy"], kw["cx"], kw["cy"], kw["mestimator"], kw["weighting"], kw["student_nu"],
                              kw["sigma_depthinv"], kw["sigma_int"], kw["bias_depthinv"], kw["bias_int"], kw["nu_depthinv"],
                              kw["nu_int"])
        A, b = self.c.build_system(*[cu(m) for m in (W0, I0, gWx, gWy, gIx, gIy, W1, I1)], p)
        return np.concatenate([np.concatenate([A[r, r:], [b[r]]]) for r in range(6)])

    def align(self, WA, IA, depth_b, rgb_b, mode, its, intr, nsamples):
        cfg = host.make_align_config(gc.ROWS, gc.COLS, gc.LEVELS, mode, iterations=its, fx=intr[0], fy=intr[1], cx=intr[2],
                                     cy=intr[3], nsamples=nsamples)
        al = host.Aligner(self.c, cfg)
        al.set_keyframe(0, cu(WA), cu(IA))
        al.set_current_rgbd(0, cu(depth_b), cu(rgb_b))
        out = al.run(want_trace=True)
        tr = out["trace"][0]
        n = al.niters
        res = dict(R=out["R"][0], t=out["t"][0], cov=out["cov"][0], sums27=[t["sums27"] for t in tr[:n]],
                   scale=[np.array([t["sigma_int"], t["sigma_depthinv"], t["bias_int"], t["bias_depthinv"], t["nu_int"],
                                    t["nu_depthinv"]], dtype=np.float32) for t in tr[:n]],
                   trace_t=[t["t"] for t in tr[:n]], cov_sums27=tr[n]["sums27"], chi=tuple(out["stats"][0]))
        al.close()
        return res

    def track_sequence(self, depth, rgb, intr, nsamples):
        acfg = host.make_align_config(gc.ROWS, gc.COLS, gc.LEVELS, capi.MODE_TRACKER, iterations=[10, 5, 3], fx=intr[0],
                                      fy=intr[1], cx=intr[2], cy=intr[3], nsamples=nsamples)
        trk = host.Tracker(self.c, host.make_tracker_config(acfg))
        poses, flags = [], []
        for k in range(depth.shape[0]):
            r = trk.track(np.ascontiguousarray(depth[k][None]), np.ascontiguousarray(rgb[k][None]))[0]
            poses.append(np.concatenate([np.array(r.R[:]), np.array(r.t[:])]))
            flags.append([r.new_odo_keyframe, r.new_integr_keyframe, r.status])
        fused = npy(trk.keyframe_map(0, 0))
        trk.close()
        return poses, flags, fused


@pytest.fixture(scope="module")
def B(ctx):
    return CudaBackend(ctx)


def test_image_ops_match_reference_kernels(B):
    gc.check_image_ops(B, gc.load())


def test_warps_fusion_visibility_match_reference_kernels(B):
    gc.check_warps(B, gc.load())


def test_scale_estimation_matches_reference_kernels(B):
    gc.check_scale(B, gc.load())


def test_normal_equations_match_reference_kernels(B):
    gc.check_systems(B, gc.load())


def test_alignment_matches_reference_pipeline(B):
    gc.check_align(B, gc.load())


def test_tracker_sequence_matches_reference_pipeline(B):
    gc.check_sequence(B, gc.load())
