"""GPU parity: b200_match_pairs (bow_tree::match_frame_and_keyframe / match_keyframes / match_for_triangulation,
robust::match_for_triangulation) against the oracle: identical match index per row."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import match
from workloads import synth
from stella_vslam_b200._lib import ERR_CAPACITY, B200Error

pytestmark = pytest.mark.gpu
THR = 0.2 * np.pi / 180.0


def _tri(seed, nodes, stereo=False, n1=2000, n2=2000, **kw):
    k1, k2, g = synth.make_keyframe_pair(seed, n1=n1, n2=n2, stereo=stereo, **kw)
    return match._triangulation_problem(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"], True, THR, nodes)


def _bow(seed, both_sides, n1=2000, n2=2000, **kw):
    k1, k2, _ = synth.make_keyframe_pair(seed, n1=n1, n2=n2, **kw)
    pr = dict(desc1=k1["desc"], angle1=k1["angle"], valid1=k1["has_landmark"], node1=k1["node"], desc2=k2["desc"], angle2=k2["angle"],
              node2=k2["node"])
    if both_sides:
        pr["valid2"] = k2["has_landmark"]
    return pr


def _check(problems, variant, lowe=0.6, check_orientation=True, max_candidates=0):
    res = match.match_pairs_batch(problems, variant, lowe, check_orientation, max_candidates)
    total = 0
    for pr, (got, n) in zip(problems, res):
        want, n_want = O.match_pairs(pr, variant, lowe, check_orientation)
        assert np.array_equal(got, want)
        assert n == n_want
        total += n
    return total


@pytest.mark.parametrize("nodes", [False, True])
@pytest.mark.parametrize("stereo", [False, True])
def test_triangulation_parity(nodes, stereo):
    assert _check([_tri(11, nodes, stereo)], match.PAIRS_TRIANGULATION) > 80


@pytest.mark.parametrize("both_sides", [False, True])
@pytest.mark.parametrize("lowe", [0.6, 0.75, 1.0])
def test_bow_parity(both_sides, lowe):
    assert _check([_bow(12, both_sides)], match.PAIRS_BOW, lowe=lowe) > 100


@pytest.mark.parametrize("variant", [0, 1])
def test_pairs_batch_ragged(variant):
    sizes = [(2000, 1800), (1, 1), (300, 2500), (3000, 70), (2, 500), (1999, 2001)]
    make = (lambda k, a, b: _bow(30 + k, bool(k & 1), a, b, n_nodes=40)) if variant == 0 else (lambda k, a, b: _tri(40 + k, bool(k & 1), bool(k & 2), a, b, n_nodes=40))
    probs = [make(k, a, b) for k, (a, b) in enumerate(sizes)]
    assert _check(probs, variant, lowe=0.75) > 100


def test_pairs_no_orientation_no_nodes_all_pairs():
    pr = _bow(50, False)
    pr["node1"] = pr["node2"] = None
    _check([pr], match.PAIRS_BOW, lowe=0.8, check_orientation=False)


def test_pairs_heavy_contention():
    """Many rows share a handful of look-alike candidates: the outcome is decided by the row order alone."""
    k1, k2, _ = synth.make_keyframe_pair(60, n1=1500, n2=40, n_nodes=1)
    base = k2["desc"][0].copy()
    rng = np.random.default_rng(0)
    noise = lambda n: rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
    k2["desc"][:] = base ^ noise(40)
    k1["desc"][:] = base ^ noise(1500)
    pr = dict(desc1=k1["desc"], angle1=k1["angle"], desc2=k2["desc"], angle2=k2["angle"])
    n = _check([pr], match.PAIRS_BOW, lowe=1.0, check_orientation=False, max_candidates=64)
    assert n == 40


def test_pairs_capacity_error():
    k1, k2, _ = synth.make_keyframe_pair(61, n1=200, n2=300, n_nodes=1)
    k2["desc"][:] = k2["desc"][0]
    k1["desc"][:] = k2["desc"][0]
    pr = dict(desc1=k1["desc"], angle1=k1["angle"], desc2=k2["desc"], angle2=k2["angle"])
    with pytest.raises(B200Error) as e:
        match.match_pairs_batch([pr], match.PAIRS_BOW, 0.6, False, 32)
    assert e.value.code == ERR_CAPACITY
    _check([pr], match.PAIRS_BOW, check_orientation=False, max_candidates=512)


def test_named_methods():
    k1, k2, g = synth.make_keyframe_pair(70, stereo=True)
    R = match.robust(0.75, True)
    got = R.match_for_triangulation(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"])
    want, n = O.match_pairs(match._triangulation_problem(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"], True, THR, False), 1, 0.75, True)
    assert len(got) == n and np.array_equal(got[:, 1], want[got[:, 0]]) and (np.diff(got[:, 0]) > 0).all()
    B = match.bow_tree(0.75, True)
    got = B.match_for_triangulation(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"])
    want, n = O.match_pairs(match._triangulation_problem(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"], True, THR, True), 1, 0.75, True)
    assert len(got) == n and np.array_equal(got[:, 1], want[got[:, 0]])
    got, n = B.match_frame_and_keyframe(k1, k2)
    want, n_want = O.match_pairs(_bow_from(k1, k2, False), 0, 0.75, True)
    assert np.array_equal(got, want) and n == n_want
    got, n = B.match_keyframes(k1, k2)
    want, n_want = O.match_pairs(_bow_from(k1, k2, True), 0, 0.75, True)
    assert np.array_equal(got, want) and n == n_want


def _bow_from(k1, k2, both_sides):
    pr = dict(desc1=k1["desc"], angle1=k1["angle"], valid1=k1["has_landmark"], node1=k1["node"], desc2=k2["desc"], angle2=k2["angle"],
              node2=k2["node"])
    if both_sides:
        pr["valid2"] = k2["has_landmark"]
    return pr
