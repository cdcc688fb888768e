"""Host-side geometry of the extractor: scale tables, per-level quotas, level sizes and the
algorithmic-byte count of SURVEY.md §8(d).  Pure float32 restatement of the few lines of the
reference constructor that size the work (src/ORBextractor.cc:415-446, :1111-1112); used by
bench.py and the multi-GPU runner, never for pixels or keypoints."""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def scale_tables(scale_factor: float, nlevels: int):
    sfd = float(f32(scale_factor))                 # `double scaleFactor` initialised from a float
    sf = np.ones(nlevels, f32)
    for i in range(1, nlevels):
        sf[i] = f32(float(sf[i - 1]) * sfd)
    sigma2 = (sf * sf).astype(f32)
    return sf, (f32(1) / sf).astype(f32), sigma2, (f32(1) / sigma2).astype(f32)


def level_quotas(nfeatures: int, scale_factor: float, nlevels: int):
    sfd = float(f32(scale_factor))
    factor = f32(1.0 / sfd)
    nd = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
    out, s = [], 0
    for _ in range(nlevels - 1):
        out.append(int(np.rint(nd)))
        s += out[-1]
        nd = f32(nd * factor)
    out.append(max(nfeatures - s, 0))
    return out


def level_sizes(width: int, height: int, scale_factor: float, nlevels: int):
    inv = scale_tables(scale_factor, nlevels)[1]
    return [(int(np.rint(f32(f32(width) * inv[l]))), int(np.rint(f32(f32(height) * inv[l])))) for l in range(nlevels)]


def algorithmic_bytes(width: int, height: int, nfeatures: int, scale_factor: float, nlevels: int) -> int:
    """B_alg = sum_l [A_max(l-1,0) + P_l] + 3*sum_l A_l + nfeatures*1321."""
    sizes = level_sizes(width, height, scale_factor, nlevels)
    total, prev = 0, 0
    for l, (w, h) in enumerate(sizes):
        a = w * h
        total += (a if l == 0 else prev) + (w + 38) * (h + 38) + 3 * a
        prev = a
    return total + nfeatures * (749 + 512 + 32 + 28)


def stage_algorithmic_bytes(width: int, height: int, nfeatures: int, scale_factor: float, nlevels: int):
    """The same figure split per stage kernel."""
    sizes = level_sizes(width, height, scale_factor, nlevels)
    areas = [w * h for w, h in sizes]
    pyr = sum((areas[0] if l == 0 else areas[l - 1]) + (w + 38) * (h + 38) for l, (w, h) in enumerate(sizes))
    # "describe" = IC_Angle + GaussianBlur + computeOrbDescriptor, fused per keypoint; its algorithmic bytes stay the
    # stage sum of what the reference does (whole-level blur included), SURVEY.md §8(d): no fusion credit.
    return {"pyramid": pyr, "fast_cells": sum(areas), "octree": 0,
            "describe": nfeatures * 749 + 2 * sum(areas) + nfeatures * (512 + 32 + 28)}
