"""Randomised parity sweep (diagnostic, not part of the test suite): tracks a batch of DISTINCT synthetic sequences with
fresh seeds on the B200 path and on the reference's own kernels (oracle/_ref) and reports the worst pose difference, the
keyframe decisions that differ and the exactness of the covisibility ratios.  Usage: python tools/stress_parity.py [base_seed] [batches]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import rot_angle  # noqa: E402
from oracle import ref as refk  # noqa: E402
from oracle.tracker import OracleTracker  # noqa: E402
from rgbid_slam_b200 import capi, host, synth  # noqa: E402


def main():
    base = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    batches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    kind = "ref" if refk.available() else "cpu"
    ctx = host.Context(0)
    rows, cols, levels, its, n_frames, B = 240, 320, 3, [10, 5, 3], 12, 6
    worst_t = worst_r = worst_vis = 0.0
    mism = 0
    total = 0
    for wf in (0, 1):
        for bi in range(batches):
            seeds = [base + 100 * bi + s + 7 * wf for s in range(B)]
            seqs = [synth.make_sequence(seed=s, n_frames=n_frames, rows=rows, cols=cols, noise=True, device="cuda") for s in seeds]
            intr = seqs[0]["intr"]
            acfg = host.make_align_config(rows, cols, levels, capi.MODE_TRACKER, batch=B, iterations=its, warp_first=wf, **intr)
            trk = host.Tracker(ctx, host.make_tracker_config(acfg))
            ots = [OracleTracker(rows, cols, intr, levels=levels, iterations=tuple(its), kind=kind, warp_first=wf) for _ in seeds]
            for k in range(n_frames):
                dd = torch.stack([q["depth"][k] for q in seqs]).contiguous()
                cc = torch.stack([q["rgb"][k] for q in seqs]).contiguous()
                res = trk.track(dd.cpu(), cc.cpu())
                for b in range(B):
                    r = res[b]
                    if kind == "ref":
                        o = ots[b].track(dd[b].cuda(), cc[b].cuda())
                    else:
                        o = ots[b].track(dd[b].cpu().numpy().astype(np.uint16), cc[b].cpu().numpy())
                    total += 1
                    if (r.status == 0) != (o["status"] == 0) or r.new_odo_keyframe != o["new_odo_keyframe"] or \
                            r.new_integr_keyframe != o["new_integr_keyframe"]:
                        mism += 1
                        print("decision mismatch: warp_first %d batch %d frame %d stream %d" % (wf, bi, k, b))
                        continue
                    if r.status != 0:
                        continue
                    worst_t = max(worst_t, float(np.linalg.norm(np.array(r.t[:]) - o["t"])))
                    worst_r = max(worst_r, rot_angle(np.array(r.R[:]).reshape(3, 3), o["R"]))
                    if k > 0:
                        worst_vis = max(worst_vis, abs(r.visibility_odo - o["visibility_odo"]), abs(r.visibility_integr - o["visibility_integr"]))
            trk.close()
    print("stress parity vs %s: %d frames, decision mismatches %d, worst |dt| %.2e m, worst angle %.2e rad, worst visibility diff %.2e"
          % (kind, total, mism, worst_t, worst_r, worst_vis))


if __name__ == "__main__":
    main()
