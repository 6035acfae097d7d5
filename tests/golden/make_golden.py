"""Generates tests/golden/ref_golden_r1.npz from the REFERENCE'S OWN CUDA kernels.

The reference ships no tests or golden vectors, so the fixtures are produced by running the
reference's device layer itself -- src/cuda/*.cu + ThirdParty/pcl_gpu_containers compiled verbatim for sm_100a
by oracle/Makefile into oracle/_ref/libref_oracle.so -- on a B200, driven through oracle/ref_shim.cpp.

    make -C oracle ref                      # here (needs /root/reference)
    gpurun -- python tests/golden/make_golden.py gpurun_out/ref_golden_r1.npz
    cp gpurun_out/ref_golden_r1.npz tests/golden/

Inputs are stored in the file, so the consumers (tests/test_golden_cpu.py: CPU oracle; tests/test_golden_gpu.py:
the new CUDA path) do not depend on the synthetic generator.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
from oracle import ref as refk  # noqa: E402
from oracle.tracker import OracleTracker  # noqa: E402
import rgbid_slam_b200  # noqa: E402,F401
from rgbid_slam_b200 import synth  # noqa: E402

ROWS, COLS, LEVELS = 96, 128, 3


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def np_(t):
    return t.detach().cpu().numpy()


def trace_arrays(tr):
    return dict(sums27=np.stack([t["sums27"] for t in tr]), x=np.stack([t["x"] for t in tr]),
                R=np.stack([t["R"] for t in tr]), t=np.stack([t["t"] for t in tr]),
                scale=np.array([[t["sigma_int"], t["sigma_depthinv"], t["bias_int"], t["bias_depthinv"], t["nu_int"],
                                 t["nu_depthinv"]] for t in tr], dtype=np.float32))


def main(out_path):
    assert refk.available(), "needs oracle/_ref/libref_oracle.so and a GPU"
    G = {}
    seq = synth.make_sequence(seed=777, n_frames=6, rows=ROWS, cols=COLS, noise=True)
    intr = seq["intr"]
    G["intr"] = np.array([intr["fx"], intr["fy"], intr["cx"], intr["cy"]], dtype=np.float64)
    depth = seq["depth"].numpy().astype(np.uint16)
    rgb = seq["rgb"].numpy()
    G["depth"], G["rgb"] = depth, rgb
    gtR, gtt = synth.relative_pose(seq["poses"][0], seq["poses"][2])
    G["gt_R_02"], G["gt_t_02"] = gtR, gtt

    dA, cA, dB, cB = cu(depth[0]), cu(rgb[0]), cu(depth[2]), cu(rgb[2])
    WA, IA = refk.convert_depth_to_invdepth(dA), refk.compute_intensity(cA)
    WB, IB = refk.convert_depth_to_invdepth(dB), refk.compute_intensity(cB)
    G["WA"], G["IA"] = np_(WA), np_(IA)
    G["invdepth_factor5"] = np_(refk.convert_depth_to_invdepth(dA, 5.0))
    G["pyr1_W"], G["pyr1_I"] = np_(refk.pyr_down(WA)), np_(refk.pyr_down(IA))
    G["pyr2_W"] = np_(refk.pyr_down(refk.pyr_down(WA)))
    gx, gy = refk.gradient(IA)
    G["gradI_x"], G["gradI_y"] = np_(gx), np_(gy)
    gwx, gwy = refk.gradient(WA)
    G["gradW_x"], G["gradW_y"] = np_(gwx), np_(gwy)
    G["bilateral_W"], G["bilateral_I"] = np_(refk.bilateral(WA, 2 * 0.0025)), np_(refk.bilateral(IA, 3.0))

    Rp, tp = orc.projective_pose(gtR, gtt, intr["fx"], intr["fy"], intr["cx"], intr["cy"], inverse=True)
    Rf, tf = orc.projective_pose(gtR, gtt, intr["fx"], intr["fy"], intr["cx"], intr["cy"], inverse=False)
    G["Rp"], G["tp"], G["Rf"], G["tf"] = Rp, tp, Rf, tf
    W1 = refk.warp_invdepth(WB, WA, Rp, tp)
    I1_kfgeom = refk.warp_intensity(IB, WA, Rp, tp)
    I1 = refk.warp_intensity(IB, W1, Rp, tp)
    G["warp_W"], G["warp_I_kfgeom"], G["warp_I"] = np_(W1), np_(I1_kfgeom), np_(I1)
    wstate = torch.full((ROWS, COLS), 0.5, device="cuda")
    Ww = refk.warp_invdepth_weighted(WB, WA, wstate, Rp, tp)
    G["warp_weighted_W"], G["warp_weighted_weight"] = np_(Ww), np_(wstate)
    kf, kfw = WA.clone(), torch.ones_like(WA)
    kf[10:20, 30:50] = float("nan")
    G["fusion_kf_in"] = np_(kf)
    refk.integrate_warped_frame(Ww, wstate, kf, kfw)
    G["fusion_kf_out"], G["fusion_weight_out"] = np_(kf), np_(kfw)
    r, mask = refk.visibility_ratio(WB, WA, Rf, tf, with_mask=True)
    G["visibility_ratio"], G["overlap_mask"] = np.float32(r), np.packbits(np_(mask))
    G["visibility_ratio_inv"] = np.float32(refk.visibility_ratio(WA, WB, Rp, tp))

    eI, eW = refk.compute_error(I1, IA, 3000), refk.compute_error(W1, WA, 3000)
    G["err_I"], G["err_W"] = np_(eI), np_(eW)
    G["sigma_nu_I"] = np.array(refk.sigma_nu_student(eI, 0.0, 5.0), dtype=np.float32)
    G["sigma_nu_W"] = np.array(refk.sigma_nu_student(eW, 0.0, 0.0025), dtype=np.float32)
    G["nu_only_I"] = np.float32(refk.nu_student(eI, 0.0, 5.0))
    G["nu_only_W"] = np.float32(refk.nu_student(eW, 0.0, 0.0025))
    for name, m in (("lsq", orc.LSQ), ("huber", orc.HUBER), ("tukey", orc.TUKEY), ("student", orc.STUDENT)):
        G["sigma_pdf_I_" + name] = np.array(refk.sigma_pdf(eI, 0.0, 5.0, m), dtype=np.float32)
        G["chi_" + name] = np.array(refk.chi_square(eI, eW, 5.0, 0.0025, m), dtype=np.float32)
    rng = np.random.default_rng(5)
    heavy = (rng.standard_t(3.0, 4096) * 2.0 + 0.1).astype(np.float32)
    heavy[::53] = np.nan
    G["heavy_err"] = heavy
    G["heavy_sigma_nu"] = np.array(refk.sigma_nu_student(cu(heavy), 0.0, 5.0), dtype=np.float32)

    sys_cfgs = [dict(student_nu=1, mestimator=orc.STUDENT, weighting=orc.INDEPENDENT),
                dict(student_nu=0, mestimator=orc.HUBER, weighting=orc.INDEPENDENT),
                dict(student_nu=0, mestimator=orc.TUKEY, weighting=orc.MIN_WEIGHT),
                dict(student_nu=0, mestimator=orc.LSQ, weighting=orc.GEOM_ONLY)]
    sums = []
    for c in sys_cfgs:
        p = orc.system_params(intr["fx"], intr["fy"], intr["cx"], intr["cy"], sigma_depthinv=0.0012, sigma_int=3.5,
                              bias_depthinv=1e-5, bias_int=0.2, nu_depthinv=4.25, nu_int=6.5, **c)
        A, b = refk.build_system(WA, IA, gwx, gwy, gx, gy, W1, I1, p)
        sums.append(np.concatenate([np.concatenate([A[r_, r_:], [b[r_]]]) for r_ in range(6)]))
    G["system_sums"] = np.stack(sums)
    G["system_cfgs"] = np.array([[c["student_nu"], c["mestimator"], c["weighting"]] for c in sys_cfgs], dtype=np.int32)

    G["vmap"] = np_(refk.vmap(WA, intr["fx"], intr["fy"], intr["cx"], intr["cy"]))
    G["nmap"] = np_(refk.nmap_gradients(WA, gwx, gwy, intr["fx"], intr["fy"], intr["cx"], intr["cy"]))

    for mode, name, its in ((orc.MODE_ALIGN, "align", [5, 5, 3]), (orc.MODE_TRACKER, "tracker", [10, 5, 3])):
        cfg = orc.make_config(ROWS, COLS, LEVELS, mode, its, intr["fx"], intr["fy"], intr["cx"], intr["cy"],
                              nsamples=3000)
        out = refk.align(cfg, refk.prepare_keyframe(WA, IA, LEVELS, mode == orc.MODE_TRACKER),
                         refk.prepare_current(WB, IB, LEVELS))
        assert out["status"] == 0
        for k, v in trace_arrays(out["trace"]).items():
            G["%s_trace_%s" % (name, k)] = v
        G[name + "_R"], G[name + "_t"], G[name + "_cov"] = out["R"], out["t"], out["cov"]
        if mode == orc.MODE_TRACKER:
            G["tracker_cov_sums27"] = out["cov_sums27"]
            G["tracker_chi"] = np.array([out["chi_square"], out["chi_test"], out["ndof"]], dtype=np.float32)

    ot = OracleTracker(ROWS, COLS, intr, levels=LEVELS, iterations=(10, 5, 3), kind="ref", nsamples=3000)
    poses, flags = [], []
    for k in range(depth.shape[0]):
        o = ot.track(cu(depth[k]), cu(rgb[k]))
        poses.append(np.concatenate([o["R"].reshape(9), o["t"]]))
        flags.append([o["new_odo_keyframe"], o["new_integr_keyframe"], o["status"]])
    G["seq_poses"], G["seq_flags"] = np.array(poses), np.array(flags, dtype=np.int32)
    G["seq_fused_kf"] = np_(ot.intW)
    np.savez_compressed(out_path, **G)
    print("wrote %s: %d arrays, %.1f KiB" % (out_path, len(G), os.path.getsize(out_path) / 1024))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_golden_r1.npz"))
