"""GPU parity: b200_match_guided (match::projection::match_frame_and_landmarks / match_current_and_last_frames) against the
oracle, bit-exact match indices and occupancy."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import match, synth
from stella_vslam_b200._lib import ERR_CAPACITY, B200Error

pytestmark = pytest.mark.gpu


def _check(problems, mode, lowe=0.8, check_orientation=True, max_candidates=0):
    P = match.projection(lowe, check_orientation)
    res = P.match_guided_batch(problems, mode, max_candidates=max_candidates)
    total = 0
    for pr, (got, occ, n) in zip(problems, res):
        want, occ_want, n_want = O.match_guided(pr, mode, lowe_ratio=lowe, check_orientation=check_orientation)
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
