"""GPU parity: b200_stereo_compute (match::stereo::compute) against the oracle.  x_right and depth are products of a short,
order-fixed float32 expression on exact integer correlations, so the comparison is bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import feature, match
from workloads import synth

pytestmark = pytest.mark.gpu
FXB, BASELINE = 435.2 * 0.11, 0.11


def _oracle(left, right, min_area):
    a = O.orb_extract(left, min_area=min_area, want_pyramid=True)
    b = O.orb_extract(right, min_area=min_area, want_pyramid=True)
    return a, b, O.stereo_compute(a["pyramid"], b["pyramid"], a["kps"], a["desc"], b["kps"], b["desc"], FXB, BASELINE)


@pytest.mark.parametrize("size,disp", [((752, 480), (9, 23, 41)), ((1241, 376), (5, 60)), ((640, 360), (0, 3))])
def test_stereo_parity_two_extractors(size, disp):
    left, right = synth.make_stereo_pair(size[0], size[1], seed=21, disparities=disp)
    a, b, (xr_want, dep_want, n_want) = _oracle(left, right, 400)
    prm = feature.orb_params()
    ex_l, ex_r = feature.orb_extractor(prm, 400), feature.orb_extractor(prm, 400)
    kl, dl = ex_l.extract(left)
    kr, dr = ex_r.extract(right)
    assert np.array_equal(kl, a["kps"]) and np.array_equal(dr, b["desc"])
    st = match.stereo(ex_l, ex_r, kl, kr, dl, dr, FXB, BASELINE)
    xr, dep = st.compute()
    assert np.array_equal(xr, xr_want) and np.array_equal(dep, dep_want)
    assert st.num_matched_ == n_want > 0.3 * len(kl)


def test_stereo_parity_one_batched_extractor():
    left, right = synth.make_stereo_pair(752, 480, seed=22)
    a, b, (xr_want, dep_want, n_want) = _oracle(left, right, 800)
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=2)
    kps, descs = ex.extract_batch(np.stack([left, right]))
    st = match.stereo(ex, ex, kps[0], kps[1], descs[0], descs[1], FXB, BASELINE, frame_left=0, frame_right=1)
    xr, dep = st.compute()
    assert np.array_equal(xr, xr_want) and np.array_equal(dep, dep_want) and st.num_matched_ == n_want


def test_stereo_filtered_and_empty_sides():
    left, right = synth.make_stereo_pair(752, 480, seed=23)
    a, b, _ = _oracle(left, right, 800)
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=2)
    kps, descs = ex.extract_batch(np.stack([left, right]))
    keep_l, keep_r = np.arange(len(kps[0]))[::2], np.arange(len(kps[1]))[1::3]
    want = O.stereo_compute(a["pyramid"], b["pyramid"], a["kps"][keep_l], a["desc"][keep_l], b["kps"][keep_r], b["desc"][keep_r], FXB, BASELINE)
    st = match.stereo(ex, ex, kps[0][keep_l], kps[1][keep_r], descs[0][keep_l], descs[1][keep_r], FXB, BASELINE, 0, 1)
    xr, dep = st.compute()
    assert np.array_equal(xr, want[0]) and np.array_equal(dep, want[1]) and st.num_matched_ == want[2]
    st = match.stereo(ex, ex, kps[0], kps[1][:0], descs[0], descs[1][:0], FXB, BASELINE, 0, 1)
    xr, dep = st.compute()
    assert (xr == -1).all() and (dep == -1).all() and st.num_matched_ == 0
    st = match.stereo(ex, ex, kps[0][:0], kps[1], descs[0][:0], descs[1], FXB, BASELINE, 0, 1)
    assert len(st.compute()[0]) == 0
