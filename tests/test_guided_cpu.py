"""Grid-guided projection matchers: the C oracle (oracle/guided_oracle.c) against an independent, literal Python walk of
match/projection.cc:13-207 + data/common.cc:83-190.  The reference ships no test or golden vector for these functions
(test/stella_vslam/match/ only has base.cc), so this restatement is the only pin: parity for rows a12/a13 is "unpinned"."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from workloads import synth


def _popcount(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _angle_diff(a1, a2):  # util/angle.cc:7-16 in float
    r = np.float32(a1) - np.float32(a2)
    if r <= -180.0:
        r = np.float32(float(r) + 360.0)
    if r > 180.0:
        r = np.float32(float(r) - 360.0)
    return r


def literal_guided(pr, mode, thr=100, lowe=0.8, check_orientation=True):
    f32 = np.float32
    min_x, max_x, min_y, max_y = [f32(v) for v in pr["bounds"]]
    gc, gr = pr.get("grid", (64, 48))
    inv_w, inv_h = gc / float(max_x - min_x), gr / float(max_y - min_y)
    n = len(pr["t_x"])
    grid = [[[] for _ in range(gr)] for _ in range(gc)]
    for i in range(n):  # data/common.cc:83-118
        cx = math.floor(float(f32(pr["t_x"][i]) - min_x) * inv_w)
        cy = math.floor(float(f32(pr["t_y"][i]) - min_y) * inv_h)
        if 0 <= cx < gc and 0 <= cy < gr:
            grid[cx][cy].append(i)
    occ = np.array(pr["t_occupied"], np.uint8).copy() if pr.get("t_occupied") is not None else np.zeros(n, np.uint8)
    xr = pr.get("t_x_right")
    out = np.full(len(pr["q_x"]), -1, np.int32)
    dist_state, owner = np.full(n, 256, np.int64), np.full(n, -1, np.int64)
    for q in range(len(pr["q_x"])):
        if pr.get("q_valid") is not None and not pr["q_valid"][q]:
            continue
        rx, ry, m = f32(pr["q_x"][q]), f32(pr["q_y"][q]), f32(pr["q_margin"][q])
        lo, hi = int(pr["q_min_level"][q]), int(pr["q_max_level"][q])
        min_cx = max(0, math.floor(float(f32(f32(rx - min_x) - m)) * inv_w))
        max_cx = min(gc - 1, math.ceil(float(f32(f32(rx - min_x) + m)) * inv_w))
        min_cy = max(0, math.floor(float(f32(f32(ry - min_y) - m)) * inv_h))
        max_cy = min(gr - 1, math.ceil(float(f32(f32(ry - min_y) + m)) * inv_h))
        if min_cx >= gc or max_cx < 0 or min_cy >= gr or max_cy < 0:
            continue
        best, second, best_lv, second_lv, best_idx = 256, 256, -1, -1, -1
        for cx in range(min_cx, max_cx + 1):
            for cy in range(min_cy, max_cy + 1):
                for idx in grid[cx][cy]:
                    o = int(pr["t_octave"][idx])
                    if 0 <= lo and o < lo:
                        continue
                    if 0 <= hi and hi < o:
                        continue
                    if not (abs(f32(pr["t_x"][idx]) - rx) < m and abs(f32(pr["t_y"][idx]) - ry) < m):
                        continue
                    if mode == 4:  # area.cc:41-62
                        if check_orientation and abs(_angle_diff(pr["q_angle"][q], pr["t_angle"][idx])) > 30.0:
                            continue
                        d = _popcount(pr["q_desc"][q], pr["t_desc"][idx])
                        if dist_state[idx] <= d:
                            continue
                        if d < best:
                            second, best, best_idx = best, d, idx
                        elif d < second:
                            second = d
                        continue
                    if occ[idx]:
                        continue
                    if mode <= 1 and xr is not None and xr[idx] > 0 and m < abs(f32(pr["q_x_right"][q]) - f32(xr[idx])):
                        continue
                    if mode == 1 and check_orientation and abs(_angle_diff(pr["q_angle"][q], pr["t_angle"][idx])) > 30.0:
                        continue
                    if mode == 3 and pr.get("do_reprojection_matching"):  # fuse.cc:93-120
                        e_x = float(pr["q_reproj"][q][0]) - float(f32(pr["t_x"][idx]))
                        e_y = float(pr["q_reproj"][q][1]) - float(f32(pr["t_y"][idx]))
                        inv_sigma = float(f32(pr["inv_level_sigma_sq"][o]))
                        if xr is not None and xr[idx] >= 0:
                            e_xr = f32(pr["q_x_right"][q]) - f32(xr[idx])
                            if float(f32(7.81473)) < (e_x * e_x + e_y * e_y + float(f32(e_xr * e_xr))) * inv_sigma:
                                continue
                        elif float(f32(5.99146)) < (e_x * e_x + e_y * e_y) * inv_sigma:
                            continue
                    d = _popcount(pr["q_desc"][q], pr["t_desc"][idx])
                    if d < best:
                        second, second_lv, best, best_lv, best_idx = best, best_lv, d, o, idx
                    elif mode == 0 and d < second:
                        second, second_lv = d, o
        if best_idx < 0 or best > thr:
            continue
        if mode == 0 and best_lv == second_lv and f32(best) > f32(lowe) * f32(second):
            continue
        if mode == 4:
            if f32(second) * f32(lowe) < f32(best):
                continue
            if owner[best_idx] >= 0:
                out[owner[best_idx]] = -1
            owner[best_idx] = q
            dist_state[best_idx] = best
        elif mode != 2:
            # `lm && lm->has_observation()` (projection.cc:50-53, 163-166): a landmark without observations does not close the keypoint
            if pr.get("q_has_observation") is None or pr["q_has_observation"][q]:
                occ[best_idx] = 1
        out[q] = best_idx
    return out, occ


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("stereo", [False, True])
def test_oracle_matches_literal_walk(mode, stereo):
    pr = synth.make_guided_problem(11 + mode, n_train=700, n_queries=500, mode=mode, stereo=stereo)
    got, occ, n = O.match_guided(pr, mode, lowe_ratio=0.8, check_orientation=True)
    want, occ_want = literal_guided(pr, mode)
    assert np.array_equal(got, want)
    assert np.array_equal(occ, occ_want)
    assert n == (want >= 0).sum() > 50


@pytest.mark.parametrize("mode,thr", [(2, 100), (3, 50), (4, 50)])
@pytest.mark.parametrize("stereo", [False, True])
def test_oracle_matches_literal_walk_occasional_variants(mode, thr, stereo):
    pr = synth.make_guided_problem(40 + mode, n_train=700, n_queries=600, mode=mode, stereo=stereo)
    if mode != 3:
        pr["t_occupied"] = np.zeros(700, np.uint8) if mode == 4 else pr["t_occupied"]
    got, occ, n = O.match_guided(pr, mode, thr=thr, lowe_ratio=0.9, check_orientation=True)
    want, occ_want = literal_guided(pr, mode, thr=thr, lowe=0.9)
    assert np.array_equal(got, want)
    assert n == (want >= 0).sum() > 30
    if mode == 3:
        assert np.array_equal(occ, occ_want)
    else:
        assert np.array_equal(occ, pr["t_occupied"])               # modes 2 and 4 do not write occupancy back
    if mode == 2:
        assert len(np.unique(got[got >= 0])) < (got >= 0).sum()    # stateless: several landmarks may pick one keypoint
    if mode == 4:
        hit = got[got >= 0]
        assert len(np.unique(hit)) == len(hit)                    # after stealing every keypoint has one owner


def test_cross_check():
    a = np.array([3, -1, 0, 2, 7], np.int32)
    b = np.array([2, 1, 5, 0], np.int32)
    out, n = O.cross_check(a, b)
    assert out.tolist() == [3, -1, 0, -1, -1] and n == 2


def test_oracle_guided_properties():
    pr = synth.make_guided_problem(5, n_train=1500, n_queries=1200, mode=0)
    got, occ, n = O.match_guided(pr, 0, lowe_ratio=0.8)
    hit = got[got >= 0]
    assert len(np.unique(hit)) == len(hit)                        # a keypoint receives at most one landmark
    assert not pr["t_occupied"][hit].any()                        # never a pre-occupied keypoint
    assert (got[pr["q_valid"] == 0] == -1).all()
    assert occ.sum() == pr["t_occupied"].sum() + n
    # every accepted pair is inside the window, the level range and the distance threshold
    for q in np.flatnonzero(got >= 0)[:200]:
        i = got[q]
        assert abs(pr["t_x"][i] - pr["q_x"][q]) < pr["q_margin"][q] and abs(pr["t_y"][i] - pr["q_y"][q]) < pr["q_margin"][q]
        assert _popcount(pr["q_desc"][q], pr["t_desc"][i]) <= 100
        if pr["q_min_level"][q] >= 0:
            assert pr["q_min_level"][q] <= pr["t_octave"][i] <= pr["q_max_level"][q]


def test_oracle_guided_empty():
    pr = synth.make_guided_problem(1, n_train=50, n_queries=40)
    empty_q = dict(pr, q_desc=np.zeros((0, 32), np.uint8), q_x=np.zeros(0), q_y=np.zeros(0), q_margin=np.zeros(0), q_min_level=np.zeros(0),
                   q_max_level=np.zeros(0), q_angle=np.zeros(0), q_valid=np.zeros(0))
    got, occ, n = O.match_guided(empty_q, 0)
    assert n == 0 and len(got) == 0 and np.array_equal(occ, pr["t_occupied"])


def test_library_cross_check_matches_oracle_without_gpu():
    """b200_match_cross_check is host-side glue (the closing loop of match_keyframes_mutually): same answer as the oracle, no device."""
    from stella_vslam_b200 import match
    rng = np.random.default_rng(4)
    a = rng.integers(-1, 50, 200).astype(np.int32)
    b = rng.integers(-1, 200, 50).astype(np.int32)
    got, n = match.cross_check(a, b)
    want, n_want = O.cross_check(a, b)
    assert np.array_equal(got, want) and n == n_want


@pytest.mark.parametrize("mode", [0, 1])
def test_landmarks_without_observations_do_not_close_keypoints(mode):
    """projection.cc:50-53 / 163-166: the occupancy test is `lm && lm->has_observation()`.  Temporal landmarks (stereo / RGBD last
    frame, no observation yet) leave the keypoint they were attached to open: a later landmark may take the same keypoint."""
    pr = synth.make_guided_problem(91 + mode, n_train=1500, n_queries=2500, mode=mode, stereo=bool(mode))
    rng = np.random.default_rng(7)
    pr["q_has_observation"] = (rng.random(len(pr["q_x"])) > 0.5).astype(np.uint8)
    got, occ, n = O.match_guided(pr, mode, lowe_ratio=0.8, check_orientation=True)
    want, want_occ = literal_guided(pr, mode, lowe=0.8, check_orientation=True)
    assert np.array_equal(got, want) and np.array_equal(occ.astype(bool), want_occ.astype(bool))
    hit = got[got >= 0]
    assert len(hit) > len(np.unique(hit)) > 100          # some keypoints were taken more than once ...
    all_obs = dict(pr, q_has_observation=None)
    base, _, _ = O.match_guided(all_obs, mode, lowe_ratio=0.8, check_orientation=True)
    assert len(base[base >= 0]) == len(np.unique(base[base >= 0])) and not np.array_equal(base, got)   # ... which never happens otherwise
