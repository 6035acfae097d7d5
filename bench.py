#!/usr/bin/env python
"""bench.py -- RGB-D frames/s of the dense frame-to-keyframe tracking path (640x480, 4-level pyramid).

One "step" = one pass of the hot path over one batch of synthetic input: every one of the `streams_per_gpu`
independent TUM-format synthetic RGB-D streams resident on a GPU advances by one frame, i.e. everything
VisodoTracker::trackNewFrame does per frame: ingest (uint16 depth + RGB8), pyramid, the
full coarse-to-fine Gauss-Newton schedule with sigma / nu estimation, the covariance pass with the
end-of-frame chi^2, both covisibility tests, keyframe switching and inverse-depth fusion.

  value   : frames/s with the frames already resident in HBM when the timed region starts
  e2e     : the same through the C-ABI call with HOST buffers (H2D of every frame + D2H of the results inside
            the timed region)
  roofline: the fused warp+residual+J^T J kernel at level 0, algorithmic bytes = 32 B per keyframe pixel
  cpu_baseline / --impl reference: see DESIGN.md ("Measurement")

Multi-GPU: one process per GPU (torchrun), streams sharded across ranks (weak scaling, no data-path
collective); the per-stream 6x6 systems / poses are all-gathered with NCCL once per step on a side stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RGB-D frames/s (640x480, 4-level pyr)"
UNIT = "frames/s"
ALGO_BYTES_PER_PX = 32  # 6 keyframe maps + 2 current-frame maps, fp32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("RGBID_BENCH_STREAMS", "32")),
                    help="independent RGB-D streams per GPU (batch of one step)")
    ap.add_argument("--rows", type=int, default=480)
    ap.add_argument("--cols", type=int, default=640)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warp-order", default="pyrFirst", choices=["pyrFirst", "warpFirst"],
                    help="WARP_ORDER of the tracker; the headline metric is quoted on the reference's shipped pyrFirst")
    ap.add_argument("--ref-streams", type=int, default=2, help="streams per step for --impl reference")
    ap.add_argument("--mode", default="tracker", choices=["tracker", "align"],
                    help="tracker: the headline (trackNewFrame per frame); align: KeyframeAlign::alignKeyframes on resident "
                         "pyramids ({5,5,3,0}, 19 200 samples), pairs/s as the value")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 latency and KeyframeAlign side measurements")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.max_mhz = gpu_index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        import datetime
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        # a rank that leaves the collective sequence must fail the run in two minutes, not hold the box for ten
        dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    return world, rank, local


def shard_streams(total_streams, world, rank):
    """Stream s lives on rank s mod world; returns the global ids owned by `rank`."""
    return [s for s in range(total_streams) if s % world == rank]


def gather_systems(local_sys, world, out=None):
    """All-gather of the per-stream results ([streams, 48] float64: 6x6 system/covariance, R, t) across ranks --
    the one exchange step of the batched mode.  NCCL on the GPU, gloo in the CPU tests."""
    import torch.distributed as dist
    if out is None:
        out = torch.empty((world * local_sys.shape[0],) + tuple(local_sys.shape[1:]), dtype=local_sys.dtype,
                          device=local_sys.device)
    if dist.get_backend() == "gloo":
        parts = [torch.empty_like(local_sys) for _ in range(world)]
        dist.all_gather(parts, local_sys)
        out.copy_(torch.cat(parts, 0))
    else:
        dist.all_gather_into_tensor(out, local_sys)
    return out


def make_frames(args, stream_ids, n_frames, device):
    """[n_frames, S, rows, cols] uint16 depth and [n_frames, S, rows, cols, 3] uint8 RGB, rendered on `device`."""
    from rgbid_slam_b200 import synth
    intr = synth.intrinsics_for(args.rows, args.cols)
    S = len(stream_ids)
    depth = torch.empty(n_frames, S, args.rows, args.cols, dtype=torch.uint16, device=device)
    rgb = torch.empty(n_frames, S, args.rows, args.cols, 3, dtype=torch.uint8, device=device)
    for j, sid in enumerate(stream_ids):
        scene = synth.Scene(20261017 + 1 + sid)            # seed = 20261017 + config index (+ stream)
        poses = synth.trajectory(n_frames, seed=sid)
        for k, (R, t) in enumerate(poses):
            d, c, _, _ = scene.render(R, t, args.rows, args.cols, intr=intr, device=device, frame_id=k, noise=True)
            depth[k, j], rgb[k, j] = d, c
    return depth, rgb, intr


def run_b200(args, world, rank, local):
    from rgbid_slam_b200 import capi, host
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    S, K, W = args.streams, args.steps, args.warmup
    stream_ids = shard_streams(S * world, world, rank)
    if args.mode == "align":
        return run_align(args, world, rank, local, stream_ids)
    n_frames = 1 + W + K + 1  # + 1: the end-to-end arm uploads frame k + 1 while it tracks frame k, also in the last step
    depth, rgb, intr = make_frames(args, stream_ids, n_frames, device)
    ctx = host.Context(local)
    its = host.default_iterations(args.levels, capi.MODE_TRACKER)
    acfg = host.make_align_config(args.rows, args.cols, args.levels, capi.MODE_TRACKER, batch=S, iterations=its,
                                  warp_first=int(args.warp_order == "warpFirst"), **intr)
    trk = host.Tracker(ctx, host.make_tracker_config(acfg))
    state_bytes = int(ctx.lib.rgbid_aligner_state_bytes())

    gather = None
    side = None
    if world > 1:
        # The one exchange step of the batched mode: all-gather of the per-stream results.  The
        # payload never leaves the device: a kernel on the tracker's stream packs [S, 48] doubles (6x6 covariance, R, t)
        # from the solver state, an event orders the NCCL all-gather behind it on a side stream, and two buffers
        # alternate so that the next step never waits for the collective.  The side stream is joined before the
        # closing event, so the collective is inside `value`.
        import torch.distributed as dist
        side = torch.cuda.Stream(device)
        local_sys = [torch.zeros(S, 48, dtype=torch.float64, device=device) for _ in range(2)]
        all_sys = [torch.zeros(world * S, 48, dtype=torch.float64, device=device) for _ in range(2)]
        packed = [torch.cuda.Event() for _ in range(2)]
        gathered = [torch.cuda.Event() for _ in range(2)]
        state = {"k": 0}

        def gather(_results):
            slot = state["k"] & 1
            state["k"] += 1
            ctx.stream.wait_event(gathered[slot])        # the collective of two steps ago has read this buffer
            trk.export_systems(local_sys[slot])          # kernel on ctx.stream, device -> device
            packed[slot].record(ctx.stream)
            side.wait_event(packed[slot])
            with torch.cuda.stream(side):
                gather_systems(local_sys[slot], world, all_sys[slot])
                gathered[slot].record(side)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(frames_d, frames_c, host_path, tracker=None, steps=None):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the context's stream (one event per
        step boundary: the spread of the steps is reported, the value is total / K)."""
        tr = tracker or trk
        n = steps or K
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        poses0 = []
        # the side measurements (tracker != None) run on one rank only: no collective there
        sync_all = barrier if tracker is None else (lambda: torch.cuda.synchronize(device))
        sync_all()
        launches0 = ctx.launches
        t0 = time.perf_counter()
        with torch.cuda.stream(ctx.stream):
            ev[0].record()
            for k in range(n):
                if host_path and not os.environ.get("RGBID_BENCH_NO_PREFETCH"):
                    # steady-state pipeline: every step issues exactly one upload (of the NEXT frame) inside the
                    # timed region; the frame tracked in step 0 was uploaded by the last warm-up step the same way
                    tr.prefetch(frames_d[k + 1], frames_c[k + 1])
                res = tr.track(frames_d[k], frames_c[k])
                if gather and tracker is None:
                    gather(res)
                poses0.append((np.array(res[0].R[:]).reshape(3, 3), np.array(res[0].t[:]), res[0].status))
                if k + 1 < n:
                    ev[k + 1].record()
            if side is not None and tracker is None:
                ctx.stream.wait_stream(side)  # the collective belongs to the step: join it before the closing event
            ev[n].record()
        sync_all()
        wall = time.perf_counter() - t0
        per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]
        dev_s = ev[0].elapsed_time(ev[n]) / 1e3
        lost = sum(1 for r in res if r.status != 0)
        return dev_s, wall, ctx.launches - launches0, lost, per_step, poses0

    # ---- device-resident arm ----------------------------------------------------------------------------
    trk.track(depth[0], rgb[0])                      # frame 0: keyframe initialisation
    for k in range(1, 1 + W):
        trk.track(depth[k], rgb[k])
    sampler = ClockSampler(local)
    sampler.start()
    dev_s, wall_s, launches, lost, per_step, poses0 = timed(depth[1 + W:], rgb[1 + W:], False)
    clocks = sampler.stop()
    ms_build = trk.time_build(level=0, reps=20)

    # ---- end-to-end arm: host buffers, H2D + D2H inside the timed region ----------------------------------
    h_depth = depth.cpu().pin_memory()
    h_rgb = rgb.cpu().pin_memory()
    trk.reset()
    trk.track(h_depth[0], h_rgb[0])
    for k in range(1, 1 + W):
        if not os.environ.get("RGBID_BENCH_NO_PREFETCH"):
            trk.prefetch(h_depth[k + 1], h_rgb[k + 1])
        trk.track(h_depth[k], h_rgb[k])
    e2e_dev_s, e2e_wall_s, _, _, e2e_per_step, _ = timed(h_depth[1 + W:], h_rgb[1 + W:], True)
    # batch-1 latency and KeyframeAlign mode are single-GPU numbers: the N = 1 run carries them
    extras = {} if (args.no_extras or world > 1) else side_measurements(args, ctx, depth, rgb, h_depth, h_rgb, intr, its, timed)

    def allmax(x):
        if world == 1:
            return x
        import torch.distributed as dist
        tns = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    dev_s, e2e_s = allmax(dev_s), allmax(max(e2e_dev_s, e2e_wall_s))
    total_frames = S * world * K
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        algo_bytes = ALGO_BYTES_PER_PX * args.rows * args.cols * S
        achieved = algo_bytes / (ms_build * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        bytes_in = S * args.rows * args.cols * 5
        out = {
            "metric": METRIC, "value": total_frames / dev_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3,
            "ms_per_step_spread": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "tum_synth_640x480_4lvl_tracker" if (args.rows, args.cols, args.levels) == (480, 640, 4)
                       else "tum_synth_%dx%d_%dlvl_tracker" % (args.cols, args.rows, args.levels),
                       "streams_per_gpu": S, "frames_per_step": S * world, "iterations": its,
                       "sigma_estimator": "sigmaML", "m_estimator": "Student", "warp_order": args.warp_order,
                       "l2_policy": "inputs larger than L2: %.0f MB of pyramids touched per step, new frames every step"
                                    % (S * 39.0)},
            "clocks": clocks,
            "e2e": {"value": total_frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bytes_in,
                    "d2h_bytes_per_step": int(S * (state_bytes + 32)),  # per stream: solver state + 8 covisibility counters
                    "device_ms_per_step": e2e_dev_s / K * 1e3, "wall_ms_per_step": e2e_wall_s / K * 1e3,
                    "upload": "next frame prefetched on a copy stream (rgbid_tracker_prefetch)"
                              if not os.environ.get("RGBID_BENCH_NO_PREFETCH") else "inside rgbid_tracker_track"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "gn_build_fast_kernel<tracker> level 0 (fused warp+residual+JtJ, TMA-staged)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "committed ncu --set full capture (profiles/roofline_traffic.json), not measured in this run",
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "ms_per_launch": ms_build, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                         "timed": "the shipped launch: pixel loop + final sum + 6x6 solve + pose update in the last CTA of every "
                                  "stream (rgbid_aligner_time_build, CUDA events on the launching stream, 20 back-to-back launches)",
                         "note": "pixel loop alone: 64-65 us = 0.74 of the peak (profiles/r02_tail_probe.txt); the launch adds reduction + election "
                                 "and the single-thread 6x6 solve / pose update on its critical path, see profiles/README.md"},
            "lost_streams": lost,
            "collective": None if world == 1 else "NCCL all-gather of [streams, 48] f64 per step from device memory, side stream, inside `value`",
        }
        out.update(extras)
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only: the scaling runs time the GPUs, not the host
            out["cpu_baseline"], out["parity"] = cpu_baseline(args, depth[:, 0].cpu(), rgb[:, 0].cpu(), intr, its, poses0, 1 + W)
    trk.close()
    ctx.close()
    return out


def run_align(args, world, rank, local, stream_ids):
    """--mode align: the KeyframeAlign schedule as the headline of the line (pairs/s), device-resident pyramids."""
    from rgbid_slam_b200 import host
    device = torch.device("cuda", local)
    depth, rgb, intr = make_frames(args, stream_ids, 4, device)
    ctx = host.Context(local)
    a = align_throughput(args, ctx, depth, rgb, intr, reps=max(1, args.steps))
    ctx.close()
    val = a["pairs_per_s"]
    if world > 1:
        import torch.distributed as dist
        tns = torch.tensor([val], dtype=torch.float64, device=device)
        dist.all_reduce(tns, op=dist.ReduceOp.SUM)
        val = float(tns.item())
    if rank != 0:
        return None
    return {"metric": "frame pairs/s (640x480, 4-level pyr, KeyframeAlign {5,5,3,0})", "value": val, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": a["ms_per_batch"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "tum_synth_640x480_4lvl_keyframe_align", "pairs_per_gpu": len(stream_ids)}, "align_mode": a}


def cpu_baseline(args, depth, rgb, intr, its, gpu_poses=None, first_timed=0, max_seconds=20.0):
    """The CPU oracle port (oracle.c, OpenMP) tracking stream 0 of the same workload for a bounded sample.  The same
    run is the checker of the `parity` object: the poses the CUDA path produced for stream 0 in the timed region
    (recorded outside of it) against the oracle's on the same frames."""
    import oracle as orc  # the one place bench.py may execute the oracle: the reported CPU baseline (and its parity check)
    from oracle.tracker import OracleTracker
    orc.lib()
    levels = args.levels
    ot = OracleTracker(args.rows, args.cols, intr, levels=levels, iterations=tuple(its), kind="cpu")
    ot.track(depth[0].numpy().astype(np.uint16), rgb[0].numpy())
    threads = int(orc.lib().orc_max_threads())
    worst_t = worst_r = 0.0
    compared = 0

    def step(k):
        nonlocal worst_t, worst_r, compared
        o = ot.track(depth[k].numpy().astype(np.uint16), rgb[k].numpy())
        j = k - first_timed
        if gpu_poses is not None and 0 <= j < len(gpu_poses) and o["status"] == 0 and gpu_poses[j][2] == 0:
            R, t, _ = gpu_poses[j]
            dR = R @ o["R"].T
            worst_t = max(worst_t, float(np.linalg.norm(t - o["t"])))
            worst_r = max(worst_r, float(np.arccos(np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0))))
            compared += 1

    n, t0, k = 0, time.perf_counter(), 1
    many_until = depth.shape[0] if threads == 1 else max(2, (3 * depth.shape[0]) // 4)  # keep frames for the 1-core leg
    while k < many_until:
        step(k)
        n += 1; k += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    # the same port on ONE core, on the frames that follow, for a quarter of the time
    one = None
    if threads > 1 and k < depth.shape[0]:
        orc.lib().orc_set_num_threads(1)
        n1, t1 = 0, time.perf_counter()
        while k < depth.shape[0]:
            step(k)
            n1 += 1; k += 1
            if time.perf_counter() - t1 > max_seconds / 4:
                break
        one = n1 / (time.perf_counter() - t1)
        orc.lib().orc_set_num_threads(threads)
    out = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d frames of one stream of the same workload (oracle.c, OpenMP over rows)" % n}
    if one is not None:
        out["value_1core"] = one
    parity = {"checker": "CPU oracle (oracle.c), stream 0, outside the timed region", "frames_compared": compared,
              "max_translation_err_m": worst_t, "max_rotation_err_rad": worst_r, "bar": "1e-4 m / 1e-4 rad",
              "ok": bool(compared > 0 and worst_t < 1e-4 and worst_r < 1e-4)}
    return out, parity


def side_measurements(args, ctx, depth, rgb, h_depth, h_rgb, intr, its, timed):
    """Two numbers the headline does not show: single-stream latency (batch 1) and KeyframeAlign::alignKeyframes."""
    from rgbid_slam_b200 import capi, host
    out = {}
    W = args.warmup
    n = min(20, depth.shape[0] - 2 - W)
    # ---- latency: ONE stream, one frame per call (what a live camera sees) ---------------------------------------
    acfg = host.make_align_config(args.rows, args.cols, args.levels, capi.MODE_TRACKER, batch=1, iterations=its,
                                  warp_first=int(args.warp_order == "warpFirst"), **intr)
    t1 = host.Tracker(ctx, host.make_tracker_config(acfg))
    d1, c1 = depth[:, :1].contiguous(), rgb[:, :1].contiguous()
    for k in range(0, 1 + W):
        t1.track(d1[k], c1[k])
    dev_s, _, _, _, per, _ = timed(d1[1 + W:], c1[1 + W:], False, tracker=t1, steps=n)
    hd1, hc1 = h_depth[:, :1].contiguous().pin_memory(), h_rgb[:, :1].contiguous().pin_memory()
    t1.reset()
    for k in range(0, 1 + W):
        t1.track(hd1[k], hc1[k])
    _, wall_s, _, _, _, _ = timed(hd1[1 + W:], hc1[1 + W:], False, tracker=t1, steps=n)  # plain host path, no prefetch
    t1.close()
    out["latency"] = {"streams": 1, "frames": n, "device_ms_per_frame": dev_s / n * 1e3,
                      "device_ms_per_frame_median": float(np.median(per)),
                      "e2e_ms_per_frame": wall_s / n * 1e3,
                      "note": "one stream, one frame per call; e2e = wall clock around rgbid_tracker_track with host buffers"}
    # ---- KeyframeAlign mode: {5,5,3,0}, 19 200 samples, resident pyramids ---------------------------------------------
    if (args.rows, args.cols, args.levels) == (480, 640, 4):
        out["align_mode"] = align_throughput(args, ctx, depth, rgb, intr, reps=10)
    return out


def align_throughput(args, ctx, depth, rgb, intr, reps=10):
    """KeyframeAlign::alignKeyframes (src/keyframe_align.cpp:115-357) for a batch of resident frame pairs: the metric's
    "4-level pyramid" schedule {5,5,3,0}; BASELINE.md bounds it at 63.3 MB per pair (about 103 k pairs/s/GPU at the HBM
    roofline)."""
    from rgbid_slam_b200 import capi, host
    S = depth.shape[1]
    cfg = host.make_align_config(args.rows, args.cols, 4, capi.MODE_ALIGN, batch=S, **intr)
    al = host.Aligner(ctx, cfg)
    for b in range(S):
        W0, I0 = ctx.convert_depth_to_invdepth(depth[0, b]), ctx.compute_intensity(rgb[0, b])
        al.set_keyframe(b, W0, I0)
        al.set_current_rgbd(b, depth[3, b], rgb[3, b])
    R0 = np.tile(np.eye(3), (S, 1, 1))
    t0 = np.zeros((S, 3))
    out0 = al.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(ctx.stream):
        e0.record()
        for _ in range(reps):
            al.enqueue(R0, t0)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    al.close()
    bytes_pair = 32 * sum(n * (args.rows >> l) * (args.cols >> l) for l, n in enumerate([5, 5, 3, 0]))
    return {"pairs_per_s": S / (ms * 1e-3), "ms_per_batch": ms, "pairs_per_batch": S, "schedule": [5, 5, 3, 0], "samples": 19200,
            "algorithmic_bytes_per_pair": bytes_pair, "hbm_bound_pairs_per_s": 6545e9 / bytes_pair,
            "failed_pairs": int((out0["status"] != 0).sum())}


def run_reference(args, world, rank, local):
    """Reference arm: the reference's OWN implementation of the path -- its CUDA kernels and bridge functions
    compiled verbatim (oracle/_ref/libref_oracle.so) driven by the restated host loop -- on the same workload,
    frame by frame and stream by stream as the reference processes them.  Falls back to the CPU port when the
    reference library is not present."""
    if rank != 0:
        return None
    import oracle as orc
    from oracle import ref as refk
    from oracle.tracker import OracleTracker
    use_ref = refk.available()
    device = torch.device("cuda", local) if use_ref else torch.device("cpu")
    S, K, W = max(1, args.ref_streams), args.steps, args.warmup
    n_frames = 1 + W + K
    depth, rgb, intr = make_frames(args, list(range(S)), n_frames, "cuda" if torch.cuda.is_available() else "cpu")
    from rgbid_slam_b200 import capi, host
    its = host.default_iterations(args.levels, capi.MODE_TRACKER)
    trackers = [OracleTracker(args.rows, args.cols, intr, levels=args.levels, iterations=tuple(its),
                              kind="ref" if use_ref else "cpu") for _ in range(S)]
    if use_ref:
        feed = lambda k, j: (depth[k, j].contiguous(), rgb[k, j].contiguous())
    else:
        dn, cn = depth.cpu().numpy().astype(np.uint16), rgb.cpu().numpy()
        feed = lambda k, j: (dn[k, j], cn[k, j])
    for k in range(0, 1 + W):
        for j in range(S):
            trackers[j].track(*feed(k, j))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(1 + W, n_frames):
        for j in range(S):
            trackers[j].track(*feed(k, j))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    val = S * K / dt
    kind = "reference" if use_ref else "port"
    return {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "tum_synth_640x480_4lvl_tracker", "streams_per_gpu": S, "frames_per_step": S,
                   "iterations": its,
                   "note": ("the reference's own CUDA kernels + bridge functions (sm_100a build of src/cuda/*.cu) with the "
                            "restated trackNewFrame host loop; the reference has no CPU implementation of this path and "
                            "no batching, frames are processed one at a time") if use_ref else
                           "CPU oracle port (reference library not present)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                         "sample": "%d frames x %d streams, one frame at a time" % (K, S)},
        "gpus_used": 1,  # rank 0 alone runs this arm at every N (bench contract): compare per GPU, or at N = 1
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    args = parse()
    world, rank, local = dist_setup(args)
    try:
        if args.impl == "reference":
            out = run_reference(args, world, rank, local)
        else:
            out = run_b200(args, world, rank, local)
        if rank == 0 and out is not None:
            print(json.dumps(out))
            sys.stdout.flush()
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
