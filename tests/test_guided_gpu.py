"""GPU parity: b200_match_guided (match::projection::match_frame_and_landmarks / match_current_and_last_frames) against the
oracle, bit-exact match indices and occupancy."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import match
from workloads import synth
from stella_vslam_b200._lib import ERR_CAPACITY, B200Error

pytestmark = pytest.mark.gpu


def _check(problems, mode, lowe=0.8, check_orientation=True, max_candidates=0, thr=100):
    res = match.match_guided_batch(problems, mode, thr, lowe, check_orientation, max_candidates)
    total = 0
    for pr, (got, occ, n) in zip(problems, res):
        want, occ_want, n_want = O.match_guided(pr, mode, thr=thr, lowe_ratio=lowe, check_orientation=check_orientation)
        assert np.array_equal(got, want)
        assert n == n_want
        if pr.get("t_occupied") is not None:
            assert np.array_equal(occ, occ_want)
        total += n
    return total


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("stereo", [False, True])
def test_guided_parity(mode, stereo):
    assert _check([synth.make_guided_problem(21 + mode, mode=mode, stereo=stereo)], mode) > 300


@pytest.mark.parametrize("mode", [0, 1])
def test_guided_batch_ragged(mode):
    sizes = [(2000, 1500), (1, 1), (300, 2500), (4000, 100), (0, 50), (50, 0), (2017, 2017)]
    probs = []
    for k, (nt, nq) in enumerate(sizes):
        pr = synth.make_guided_problem(100 + k, n_train=max(nt, 1), n_queries=max(nq, 1), mode=mode, stereo=bool(k & 1))
        if nt == 0:
            for key in ("t_x", "t_y", "t_octave", "t_angle", "t_desc", "t_occupied", "t_x_right"):
                if key in pr:
                    pr[key] = pr[key][:0]
        if nq == 0:
            for key in [k2 for k2 in pr if k2.startswith("q_")]:
                pr[key] = pr[key][:0]
        probs.append(pr)
    assert _check(probs, mode) > 500


def test_guided_no_orientation_and_ratio_off():
    pr = synth.make_guided_problem(7, mode=1)
    _check([pr], 1, check_orientation=False)
    _check([pr], 0, lowe=1.0)
    _check([pr], 0, lowe=0.5)


def test_guided_contention_one_keypoint():
    """Every landmark wants the same few keypoints: the result is decided purely by the sequential order."""
    pr = synth.make_guided_problem(9, n_train=64, n_queries=3000, mode=0)
    pr["t_occupied"] = np.zeros(64, np.uint8)
    pr["q_min_level"][:] = -1
    pr["q_max_level"][:] = -1
    n = _check([pr], 0, lowe=1.0)
    assert 0 < n <= 64


def test_guided_wide_margin_capacity():
    pr = synth.make_guided_problem(13, n_train=3000, n_queries=200, mode=0)
    pr["q_margin"] = np.full(200, 80.0, np.float32)
    pr["q_min_level"][:] = -1
    pr["q_max_level"][:] = -1
    with pytest.raises(B200Error) as e:
        match.projection(0.8, True).match_guided_batch([pr], 0, max_candidates=16)
    assert e.value.code == ERR_CAPACITY
    _check([pr], 0, max_candidates=2048)


def test_guided_named_methods():
    pr = synth.make_guided_problem(31, mode=0)
    frm = {k: v for k, v in pr.items() if k.startswith("t_") or k in ("bounds", "grid", "scale_factors")}
    lv = np.random.default_rng(0).integers(0, 8, len(pr["q_x"]))
    P = match.projection(0.8, True)
    reproj = np.stack([pr["q_x"], pr["q_y"]], 1).astype(np.float64)
    got, occ, n = P.match_frame_and_landmarks(frm, pr["q_desc"], reproj, lv, margin=5.0, valid=pr["q_valid"])
    ref = dict(pr, q_margin=np.float32(5.0) * pr["scale_factors"][lv], q_min_level=np.maximum(0, lv - 1), q_max_level=np.minimum(7, lv + 1))
    want, _, n_want = O.match_guided(ref, 0, lowe_ratio=0.8)
    assert np.array_equal(got, want) and n == n_want
    got, occ, n = P.match_current_and_last_frames(frm, pr["q_desc"], reproj, lv, pr["q_angle"], 10.0, valid=pr["q_valid"], assume_forward=True)
    ref = dict(pr, q_margin=np.float32(10.0) * pr["scale_factors"][lv], q_min_level=lv, q_max_level=np.minimum(7, lv + 1))
    want, _, n_want = O.match_guided(ref, 1, check_orientation=True)
    assert np.array_equal(got, want) and n == n_want


@pytest.mark.parametrize("mode,thr", [(2, 100), (3, 50), (4, 50)])
@pytest.mark.parametrize("stereo", [False, True])
def test_guided_occasional_variants(mode, thr, stereo):
    """mode 2: one direction of match_keyframes_mutually; 3: fuse::detect_duplication with the chi-square gate; 4: area matcher."""
    probs = [synth.make_guided_problem(60 + mode + k, n_train=2000, n_queries=2500, mode=mode, stereo=stereo) for k in range(3)]
    if mode == 3:
        probs[1]["do_reprojection_matching"] = False
        for pr in probs:
            pr["t_occupied"] = None
    assert _check(probs, mode, lowe=0.9, thr=thr) > 300


def test_guided_area_steals_under_contention():
    pr = synth.make_guided_problem(77, n_train=120, n_queries=4000, mode=4)
    pr["q_margin"][:] = 40.0
    n = _check([pr], 4, lowe=1.0, thr=100, check_orientation=False)
    assert 0 < n <= 120


def _frame_of(pr):
    return {k: v for k, v in pr.items() if (k.startswith("t_") and k != "t_occupied") or k in ("bounds", "grid", "scale_factors")}


def test_keyframes_mutually_and_cross_check():
    """Keyframe 2 sees a shuffled, slightly moved copy of keyframe 1's keypoints; landmark i of 1 reprojects near keypoint perm[i] of 2."""
    a = synth.make_guided_problem(81, n_train=1500, n_queries=10, mode=2)
    rng = np.random.default_rng(3)
    perm = rng.permutation(1500)
    kf1 = _frame_of(a)
    kf2 = dict(kf1)
    flips = rng.integers(0, 256, (1500, 32), dtype=np.uint8) & rng.integers(0, 256, (1500, 32), dtype=np.uint8) & rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    for k in ("t_x", "t_y", "t_octave", "t_angle"):
        kf2[k] = np.empty_like(kf1[k])
        kf2[k][perm] = kf1[k]
    kf2["t_desc"] = np.empty_like(kf1["t_desc"])
    kf2["t_desc"][perm] = kf1["t_desc"] ^ flips
    kf2["t_x"] = (kf2["t_x"] + rng.normal(0, 1.5, 1500)).astype(np.float32)
    inv = np.argsort(perm)
    lv1, lv2 = kf1["t_octave"].astype(np.int64), kf2["t_octave"].astype(np.int64)
    noise = lambda: rng.normal(0, 2.0, (1500, 2))
    lms_1 = dict(desc=kf1["t_desc"], reproj=np.stack([kf2["t_x"][perm], kf2["t_y"][perm]], 1) + noise(), level=lv1, valid=rng.random(1500) > 0.1)
    lms_2 = dict(desc=kf2["t_desc"], reproj=np.stack([kf1["t_x"][inv], kf1["t_y"][inv]], 1) + noise(), level=lv2, valid=rng.random(1500) > 0.1)
    got, n = match.projection(0.8, True).match_keyframes_mutually(kf1, kf2, lms_1, lms_2, margin=7.5)

    def one(target, lms, lv):
        pr = dict({k: v for k, v in target.items() if k != "scale_factors"}, q_desc=lms["desc"], q_x=lms["reproj"][:, 0].astype(np.float32),
                  q_y=lms["reproj"][:, 1].astype(np.float32), q_margin=np.float32(7.5) * target["scale_factors"][lv],
                  q_min_level=np.maximum(0, lv - 1), q_max_level=np.minimum(7, lv + 1), q_valid=lms["valid"])
        return O.match_guided(pr, 2, thr=100)[0]

    want, n_want = O.cross_check(one(kf2, lms_1, lv1), one(kf1, lms_2, lv2))
    assert np.array_equal(got, want) and n == n_want and n > 500
    hit = got >= 0
    assert (got[hit] == perm[hit]).mean() > 0.95


def test_fuse_and_area_named_methods():
    pr = synth.make_guided_problem(91, mode=3, stereo=True)
    kf = _frame_of(pr)
    lv = np.random.default_rng(1).integers(0, 8, len(pr["q_x"]))
    got, n = match.fuse(0.6, True).detect_duplication(kf, pr["q_desc"], pr["q_reproj"], lv, 3.0, x_right=pr["q_x_right"], valid=pr["q_valid"],
                                                      do_reprojection_matching=True, inv_level_sigma_sq=pr["inv_level_sigma_sq"])
    ref = dict(pr, t_occupied=None, q_x=pr["q_reproj"][:, 0].astype(np.float32), q_y=pr["q_reproj"][:, 1].astype(np.float32),
               q_margin=np.float32(3.0) * pr["scale_factors"][lv], q_min_level=np.maximum(0, lv - 1), q_max_level=np.minimum(7, lv + 1))
    want, _, n_want = O.match_guided(ref, 3, thr=50)
    assert np.array_equal(got, want) and n == n_want

    f1 = synth.make_guided_problem(92, n_train=1800, n_queries=10, mode=4)
    f2 = synth.make_guided_problem(93, n_train=1800, n_queries=10, mode=4)
    f2["t_desc"][:900] = f1["t_desc"][:900]
    f2["t_octave"][:900] = f1["t_octave"][:900]
    f2["t_angle"][:900] = f1["t_angle"][:900]
    f2["t_x"][:900], f2["t_y"][:900] = f1["t_x"][:900] + 3, f1["t_y"][:900] - 2
    fr1, fr2 = _frame_of(f1), _frame_of(f2)
    prev = np.stack([f1["t_x"], f1["t_y"]], 1)
    got, n, new_prev = match.area(0.9, True).match_in_consistent_area(fr1, fr2, prev, 50)
    ref = dict({k: v for k, v in fr2.items() if k != "t_x_right"}, q_desc=f1["t_desc"], q_x=f1["t_x"], q_y=f1["t_y"],
               q_margin=np.full(1800, 50, np.float32), q_min_level=np.zeros(1800, np.int8), q_max_level=np.zeros(1800, np.int8),
               q_angle=f1["t_angle"], q_valid=(f1["t_octave"] == 0).astype(np.uint8))
    want, _, n_want = O.match_guided(ref, 4, thr=50, lowe_ratio=0.9)
    assert np.array_equal(got, want) and n == n_want and n > 20
    hit = want >= 0
    assert np.array_equal(new_prev[hit, 0], f2["t_x"][want[hit]]) and np.array_equal(new_prev[~hit], prev[~hit])


@pytest.mark.parametrize("mode", [0, 1])
def test_guided_landmarks_without_observations(mode):
    # projection.cc:50-53, 163-166 (`lm && lm->has_observation()`): keypoints given to temporal landmarks stay open
    probs = []
    for k in range(3):
        pr = synth.make_guided_problem(95 + 2 * k + mode, n_train=1800, n_queries=2600, mode=mode, stereo=bool(k & 1))
        pr["q_has_observation"] = (np.random.default_rng(k).random(len(pr["q_x"])) > 0.4).astype(np.uint8)
        probs.append(pr)
    n = _check(probs, mode)
    assert n > 300
