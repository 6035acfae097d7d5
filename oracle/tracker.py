"""Restatement of VisodoTracker::trackNewFrame (src/visodo.cpp:1967-2247) on top of a kernel backend:
`oracle` itself (CPU, numpy arrays) or `oracle.ref` (the reference's own CUDA kernels, torch CUDA tensors).
TEST INFRASTRUCTURE ONLY -- this is the checker for the tracker and the reference arm of bench.py."""
import numpy as np

import oracle as orc


class _CpuBackend:
    name = "cpu-oracle"
    depth_to_invdepth = staticmethod(orc.depth_to_invdepth)
    intensity = staticmethod(orc.intensity)
    prepare_keyframe = staticmethod(orc.prepare_keyframe)
    prepare_current = staticmethod(orc.prepare_current)
    align = staticmethod(orc.align)
    visibility_ratio = staticmethod(orc.visibility_ratio)
    gradient = staticmethod(orc.gradient)
    vmap = staticmethod(orc.vmap)
    nmap_gradients = staticmethod(orc.nmap_gradients)

    @staticmethod
    def copy(a):
        return a.copy()

    @staticmethod
    def ones_like(a):
        return np.ones_like(a)

    @staticmethod
    def zeros_like(a):
        return np.zeros_like(a)

    @staticmethod
    def fuse(cur_W, kf_W, kf_weight, wstate, Rp, tp):
        warped = orc.warp_invdepth_weighted(cur_W, kf_W, wstate, Rp, tp)
        orc.integrate_warped_frame(warped, wstate, kf_W, kf_weight)


class _RefBackend:
    name = "reference-cuda"

    def __init__(self):
        from oracle import ref
        self.r = ref
        self.depth_to_invdepth = ref.convert_depth_to_invdepth
        self.intensity = ref.compute_intensity
        self.prepare_keyframe = ref.prepare_keyframe
        self.prepare_current = ref.prepare_current
        self.align = ref.align
        self.visibility_ratio = ref.visibility_ratio
        self.gradient = ref.gradient
        self.vmap = ref.vmap
        self.nmap_gradients = ref.nmap_gradients

    @staticmethod
    def copy(a):
        return a.clone()

    @staticmethod
    def ones_like(a):
        import torch
        return torch.ones_like(a)

    @staticmethod
    def zeros_like(a):
        import torch
        return torch.zeros_like(a)

    def fuse(self, cur_W, kf_W, kf_weight, wstate, Rp, tp):
        warped = self.r.warp_invdepth_weighted(cur_W, kf_W, wstate, Rp, tp)
        self.r.integrate_warped_frame(warped, wstate, kf_W, kf_weight)


def backend(kind):
    return _CpuBackend() if kind == "cpu" else _RefBackend()


def _skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def relative_constraint(R_last, t_last, cov_last, R_new, t_new, cov_new):
    """Relative transform between two poses given in the same frame and its first-order covariance: the SEQ_ODO
    constraint of src/visodo.cpp:2126-2146 and the SEQ_KF constraint of :1610-1629 (same formulas)."""
    Rt = R_last.T
    R, t = Rt @ R_new, Rt @ (t_new - t_last)
    Jn, Jl = np.zeros((6, 6)), np.zeros((6, 6))
    Jn[:3, :3], Jn[3:, 3:] = Rt, Rt
    Jl[:3, :3], Jl[3:, 3:] = -Rt, -Rt
    Jl[:3, 3:] = _skew(t) @ Rt
    return R, t, Jl @ cov_last @ Jl.T + Jn @ cov_new @ Jn.T


class OracleTracker:
    def __init__(self, rows, cols, intr, levels=3, iterations=(10, 5, 3), kind="cpu", motion_model=True,
                 visratio_odo=0.9, visratio_integr=0.7, delta_t=0.03333, mestimator=orc.STUDENT,
                 sigma_estimator=orc.SIGMA_PDF, nsamples=10000, factor_depth=1.0, warp_first=0, termination=0,
                 conv_eps=0.0):
        self.B = backend(kind)
        self.rows, self.cols, self.intr, self.levels = rows, cols, intr, levels
        self.cfg = orc.make_config(rows, cols, levels, orc.MODE_TRACKER, list(iterations), intr["fx"], intr["fy"],
                                   intr["cx"], intr["cy"], mestimator=mestimator, sigma_estimator=sigma_estimator,
                                   nsamples=nsamples, warp_first=warp_first, termination=termination,
                                   conv_eps=conv_eps)
        self.motion_model, self.vo, self.vi = motion_model, visratio_odo, visratio_integr
        self.dt = np.float32(delta_t)
        self.factor_depth = factor_depth
        self.reset()

    def reset(self):
        self.global_time = 0
        self.lost = False
        self.R_odoKF, self.t_odoKF = np.eye(3), np.zeros(3)
        self.R_est, self.t_est = np.eye(3), np.zeros(3)
        self.dR, self.dt_, self.dcov = np.eye(3), np.zeros(3), np.zeros((6, 6))
        self.R_intKF, self.t_intKF = np.eye(3), np.zeros(3)
        self.vel, self.omega = np.zeros(3), np.zeros(3)
        self.kf = None
        self.wstate = None
        # delta_*_odo2integr_{last,next}_ and last_integrKF_index_ (src/visodo.cpp:1553-1670)
        self.o2i_next = [np.eye(3), np.zeros(3), np.zeros((6, 6))]
        self.o2i_last = [np.eye(3), np.zeros(3), np.zeros((6, 6))]
        self.last_integrKF_index = 0

    def _chain_compose(self):
        """T_new = T_old * T_upd on the odometry-keyframe -> integration-keyframe chain: the common part of
        resetOdometryKeyframe (:1553-1566) and resetIntegrationKeyframe (:1592-1605)."""
        Rn, tn, Cn = self.o2i_next
        J = np.zeros((6, 6))
        J[:3, :3], J[3:, 3:] = Rn, Rn
        t_new = Rn @ self.dt_
        J[:3, 3:] = _skew(t_new)
        self.o2i_next = [Rn @ self.dR, t_new + tn, Cn + J @ self.dcov @ J.T]

    def _proj(self, R, t, inverse):
        i = self.intr
        return orc.projective_pose(R, t, i["fx"], i["fy"], i["cx"], i["cy"], inverse=inverse)

    def _covisibility(self, R_ab, t_ab, W_a, W_b):
        """computeCovisibility, src/visodo.cpp:1481-1514 (A = keyframe, B = current)."""
        Rf, tf = self._proj(R_ab, t_ab, False)
        Ri, ti = self._proj(R_ab, t_ab, True)
        r_ba = self.B.visibility_ratio(W_b, W_a, Rf, tf)
        r_ab = self.B.visibility_ratio(W_a, W_b, Ri, ti)
        return min(r_ab, r_ba)

    def _save_integration_kf(self, cur):
        B = self.B
        self.intW, self.intWraw = B.copy(cur["W"][0]), B.copy(cur["W"][0])
        self.intWeight = B.ones_like(cur["W"][0])
        if self.wstate is None:
            self.wstate = B.zeros_like(cur["W"][0])
        self._refresh_maps()

    def _refresh_maps(self):
        i = self.intr
        self.vmap = self.B.vmap(self.intW, i["fx"], i["fy"], i["cx"], i["cy"])
        gx, gy = self.B.gradient(self.intW)
        self.nmap = self.B.nmap_gradients(self.intW, gx, gy, i["fx"], i["fy"], i["cx"], i["cy"])

    def track(self, depth_u16, rgb_u8):
        B = self.B
        W, I = B.depth_to_invdepth(depth_u16, self.factor_depth), B.intensity(rgb_u8)
        cur = B.prepare_current(W, I, self.levels)
        res = dict(frame_index=self.global_time, status=0, new_odo_keyframe=0, new_integr_keyframe=0,
                   visibility_odo=1.0, visibility_integr=1.0, lost_again=0)
        if self.global_time == 0:
            self.kf = B.prepare_keyframe(W, I, self.levels, tracker=True)
            self._save_integration_kf(cur)
            self.global_time = 1
            res.update(R=np.eye(3), t=np.zeros(3), dR=np.eye(3), dt=np.zeros(3), cov=np.zeros((6, 6)),
                       new_odo_keyframe=1, new_integr_keyframe=1)
            return res
        prevR, prevt, prevcov = self.dR.copy(), self.dt_.copy(), self.dcov.copy()
        was_lost = self.lost
        if self.global_time > 1 and self.motion_model and not self.lost:
            dRp, dtp = orc.exp_map(self.omega * float(self.dt), self.vel * float(self.dt))
            Ri, ti = prevR @ dRp, prevR @ dtp + prevt
        else:
            Ri, ti = prevR, prevt
        out = B.align(self.cfg, self.kf, cur, Ri, ti)
        ok = (out["status"] == 0)
        res["status"] = 0 if ok else -1
        res["align"] = out
        if ok:
            self.dR, self.dt_, self.dcov = out["R"], out["t"], out["cov"]
            tw = orc.log_map(prevR.T @ self.dR, prevR.T @ (self.dt_ - prevt))
            inv_dt = float(np.float32(1.0) / self.dt)
            self.vel, self.omega = tw[:3] * inv_dt, tw[3:] * inv_dt
            self.lost = False
        else:
            self.dcov = 100.0 * np.eye(6)
        self.t_est = self.t_odoKF + self.R_odoKF @ self.dt_
        self.R_est = self.R_odoKF @ self.dR
        res.update(R=self.R_est.copy(), t=self.t_est.copy(), dR=self.dR.copy(), dt=self.dt_.copy(), cov=self.dcov.copy())
        dRi = np.linalg.inv(self.R_intKF) @ self.R_est
        dti = np.linalg.inv(self.R_intKF) @ (self.t_est - self.t_intKF)
        if not ok and was_lost:
            # failed again while lost (:2111-2116): re-save both keyframes from this frame, nothing else -- no reset of
            # the keyframe poses, no keyframe or constraint for the back end, global_time_ stands still
            self.kf = B.prepare_keyframe(W, I, self.levels, tracker=True)
            self._save_integration_kf(cur)
            res.update(new_odo_keyframe=1, new_integr_keyframe=1, lost_again=1, kf_handoff=None,
                       seq=(np.eye(3), np.zeros(3), 100.0 * np.eye(6)))
            return res
        if not ok:
            self.lost = True
            new_odo, new_int = True, True
        else:
            vis_odo = self._covisibility(self.dR, self.dt_, self.kf["W"][0], cur["W"][0])
            vis_int = self._covisibility(dRi, dti, self.intWraw, cur["W"][0])
            res["visibility_odo"], res["visibility_integr"] = vis_odo, vis_int
            new_odo, new_int = vis_odo < self.vo, vis_int < self.vi
        res["new_odo_keyframe"], res["new_integr_keyframe"] = int(new_odo), int(new_int)
        # sequential odometry constraint (:2126-2156) or the dummy one when lost (:2068-2071)
        res["seq"] = relative_constraint(prevR, prevt, prevcov, self.dR, self.dt_, self.dcov) if ok else \
            (np.eye(3), np.zeros(3), 100.0 * np.eye(6))
        res["kf_handoff"] = None
        if new_odo:
            self._chain_compose()
            self.R_odoKF, self.t_odoKF = self.R_est.copy(), self.t_est.copy()
            self.dR, self.dt_, self.dcov = np.eye(3), np.zeros(3), np.zeros((6, 6))
            self.kf = B.prepare_keyframe(W, I, self.levels, tracker=True)
        if new_int:
            # resetIntegrationKeyframe (:1577-1672): the outgoing keyframe with its SEQ_KF constraint
            self._chain_compose()
            kR, kt, kcov = relative_constraint(*self.o2i_last, *self.o2i_next)
            res["kf_handoff"] = dict(kf_index=self.last_integrKF_index, frame_index=self.global_time, R=self.R_intKF.copy(),
                                     t=self.t_intKF.copy(), rel_R=kR, rel_t=kt, rel_cov=kcov, depthinv=B.copy(self.intW),
                                     normals=B.copy(self.nmap))
            self.last_integrKF_index = self.global_time
            self.o2i_last = [self.dR.copy(), self.dt_.copy(), self.dcov.copy()]
            self.o2i_next = [np.eye(3), np.zeros(3), np.zeros((6, 6))]
            self.R_intKF, self.t_intKF = self.R_est.copy(), self.t_est.copy()
            self._save_integration_kf(cur)
        else:
            Rp, tp = self._proj(dRi, dti, True)
            B.fuse(cur["W"][0], self.intW, self.intWeight, self.wstate, Rp, tp)
            self._refresh_maps()
        self.global_time += 1
        return res
