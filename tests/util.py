"""Shared helpers for the parity tests (numpy / torch glue only)."""
import numpy as np
import torch

import rgbid_slam_b200  # noqa: F401
from rgbid_slam_b200 import synth
import oracle as orc


def pair_maps(seed=20261018, rows=480, cols=640, noise=False, max_trans=0.02, max_rot_deg=1.0):
    """Synthetic frame pair -> float maps via the oracle's ingest restatement (numpy)."""
    p = synth.make_pair(seed=seed, rows=rows, cols=cols, noise=noise, max_trans=max_trans, max_rot_deg=max_rot_deg)
    dA = p["depth_a"].numpy().astype(np.uint16)
    dB = p["depth_b"].numpy().astype(np.uint16)
    out = dict(p)
    out.update(dA=dA, dB=dB, cA=p["rgb_a"].numpy(), cB=p["rgb_b"].numpy())
    out["WA"], out["WB"] = orc.depth_to_invdepth(dA), orc.depth_to_invdepth(dB)
    out["IA"], out["IB"] = orc.intensity(out["cA"]), orc.intensity(out["cB"])
    return out


def cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def same_nan(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b))


def max_abs_diff(a, b):
    m = ~(np.isnan(a) | np.isnan(b))
    return float(np.max(np.abs(a[m] - b[m]))) if m.any() else 0.0


def rot_angle(Ra, Rb):
    dR = np.asarray(Ra) @ np.asarray(Rb).T
    return float(np.arccos(np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)))


def sums_rel_err(a, b, ignore_b=False):
    """Relative error of two 27-vectors [A00..A05,b0,A11..]: each entry is scaled by the geometric mean of the
    diagonal entries of its row/column (b entries by sqrt(A_ii) * ||b||-scale), i.e. by the magnitude the
    entry would have without cancellation."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    diag, idx, shift = np.zeros(6), {}, 0
    for i in range(6):
        for j in range(i, 7):
            idx[(i, j)] = shift
            shift += 1
    for i in range(6):
        diag[i] = abs(b[idx[(i, i)]])
    bscale = max(abs(b[idx[(i, 6)]]) / np.sqrt(diag[i]) for i in range(6) if diag[i] > 0) if diag.max() > 0 else 1.0
    worst = 0.0
    for (i, j), k in idx.items():
        if ignore_b and j == 6:
            continue
        scale = np.sqrt(diag[i] * diag[j]) if j < 6 else np.sqrt(diag[i]) * max(bscale, 1e-30)
        if scale > 0:
            worst = max(worst, abs(a[k] - b[k]) / scale)
    return worst
