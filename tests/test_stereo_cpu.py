"""match::stereo: the C oracle (oracle/stereo_oracle.c) against a literal numpy walk of stereo.cc:20-251 that keeps the
reference's per-row candidate vectors and float32 patches.  No reference test exists for this matcher: unpinned."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from workloads import synth

FXB, BASELINE = 435.2 * 0.11, 0.11  # EuRoC-like focal_x_baseline / true_baseline -> max_disp = fx


def _cv_round(v):
    return int(np.rint(np.float32(v)))


def literal_stereo(pl, pr, kl, dl, kr, dr, fxb, baseline):
    f32 = np.float32
    sf, inv = O.scale_factors(1.2, len(pl))[:2]
    n = len(kl)
    xr_out, depth = np.full(n, -1, f32), np.full(n, -1, f32)
    rows = [[] for _ in range(pl[0].shape[0])]
    for j in range(len(kr)):
        r = f32(2.0) * sf[kr["octave"][j]]
        for row in range(math.floor(f32(kr["y"][j] - r)), math.ceil(f32(kr["y"][j] + r)) + 1):
            rows[row].append(j)
    max_disp = f32(f32(fxb) / f32(baseline))
    corr_idx = []
    for i in range(n):
        lv, x_left, y_left = int(kl["octave"][i]), f32(kl["x"][i]), f32(kl["y"][i])
        cands = rows[int(y_left)]
        if not cands:
            continue
        min_x, max_x = f32(x_left - max_disp), f32(x_left - f32(0))
        if max_x < 0:
            continue
        best, best_j = 75, 0
        for j in cands:
            if kr["octave"][j] < lv - 1 or kr["octave"][j] > lv + 1:
                continue
            if kr["x"][j] < min_x or max_x < kr["x"][j]:
                continue
            d = int(np.unpackbits(dl[i] ^ dr[j]).sum())
            if d < best:
                best, best_j = d, j
        if 75 <= best:
            continue
        sxl, syl, sxr = _cv_round(x_left * inv[lv]), _cv_round(y_left * inv[lv]), _cv_round(f32(kr["x"][best_j]) * inv[lv])
        if sxr - 10 < 0 or pr[lv].shape[1] <= sxr + 10:
            continue
        patch_l = pl[lv][syl - 5:syl + 6, sxl - 5:sxl + 6].astype(f32)
        patch_l = patch_l - patch_l[5, 5]
        corrs, best_c, best_off = [], np.finfo(f32).max, 0
        for off in range(-5, 6):
            patch_r = pr[lv][syl - 5:syl + 6, sxr + off - 5:sxr + off + 6].astype(f32)
            patch_r = patch_r - patch_r[5, 5]
            c = f32(np.abs(patch_l - patch_r).astype(np.float64).sum())
            if c < best_c:
                best_c, best_off = c, off
            corrs.append(c)
        if best_off in (-5, 5):
            continue
        c1, c2, c3 = corrs[5 + best_off - 1], corrs[5 + best_off], corrs[5 + best_off + 1]
        x_delta = f32(float(f32(c1 - c3)) / (2.0 * float(f32(c1 + c3)) - 4.0 * float(c2)))
        if x_delta < -1.0 or 1.0 < x_delta:
            continue
        best_x = f32(sf[lv] * f32(f32(sxr + best_off) + x_delta))
        disp = f32(x_left - best_x)
        if disp < 0 or max_disp <= disp:
            continue
        if disp <= 0:
            disp = f32(0.01)
            best_x = f32(x_left - disp)
        depth[i], xr_out[i] = f32(f32(fxb) / disp), best_x
        corr_idx.append((int(best_c), i))
    corr_idx.sort()
    if corr_idx:
        med = f32(corr_idx[len(corr_idx) // 2][0])
        thr = f32(2.0 * float(med))
        for c, i in corr_idx[len(corr_idx) // 2:]:
            if thr < f32(c):
                xr_out[i], depth[i] = -1, -1
    return xr_out, depth


@pytest.fixture(scope="module")
def pair():
    left, right = synth.make_stereo_pair(480, 320, seed=5, disparities=(7, 19))
    a = O.orb_extract(left, min_area=300, want_pyramid=True)
    b = O.orb_extract(right, min_area=300, want_pyramid=True)
    return a, b


def test_stereo_oracle_matches_literal_walk(pair):
    a, b = pair
    xr, dep, n = O.stereo_compute(a["pyramid"], b["pyramid"], a["kps"], a["desc"], b["kps"], b["desc"], FXB, BASELINE)
    xr_want, dep_want = literal_stereo(a["pyramid"], b["pyramid"], a["kps"], a["desc"], b["kps"], b["desc"], FXB, BASELINE)
    assert np.array_equal(xr, xr_want) and np.array_equal(dep, dep_want)
    assert n == (xr >= 0).sum() > 0.3 * len(xr)


def test_stereo_oracle_recovers_band_disparities(pair):
    a, b = pair
    xr, dep, n = O.stereo_compute(a["pyramid"], b["pyramid"], a["kps"], a["desc"], b["kps"], b["desc"], FXB, BASELINE)
    ok = xr >= 0
    disp = a["kps"]["x"][ok] - xr[ok]
    truth = np.where(a["kps"]["y"][ok] < 160, 7.0, 19.0)
    inner = np.abs(a["kps"]["y"][ok] - 160) > 12          # away from the seam between the two bands
    sf = O.scale_factors(1.2, 8)[0][a["kps"]["octave"][ok]]                # sub-pixel at the keypoint's own pyramid level
    assert (np.abs(disp - truth)[inner] < 1.0 * sf[inner]).mean() > 0.97
    assert np.allclose(dep[ok], np.float32(FXB) / disp.astype(np.float32), rtol=1e-6)
    assert (dep[~ok] == -1).all()
