"""CPU tests: the C oracle against the committed golden fixtures (cv2 4.13.0 outputs + the reference's own KATs).

These pin the oracle BEFORE it is used to judge the CUDA path.
"""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as O


@pytest.fixture(scope="module")
def prims(golden_dir):
    return np.load(os.path.join(golden_dir, "prims_cv2.npz"))


def test_scale_factor_recurrence():
    # test/stella_vslam/feature/orb_params.cc:29-71: float recurrences
    sf, inv, sig, isig = O.scale_factors(1.2, 8)
    s = np.float32(1.0)
    for l in range(8):
        if l:
            s = np.float32(1.2) * s
        assert sf[l] == s
        assert sig[l] == (np.float32(1.0) if l == 0 else s * s)
        assert isig[l] == (np.float32(1.0) if l == 0 else np.float32(1.0) / (s * s))
    assert O.level_sizes(1920, 1080)[1:4] == [(1600, 900), (1333, 750), (1111, 625)]
    assert O.level_sizes(752, 480)[-1] == (210, 134)


def test_resize_matches_cv2(prims):
    for name in ("base", "rnd"):
        im = prims[name]
        for tag in "abc":
            ref = prims[f"resize_{name}_{tag}"]
            got = O.resize_linear(im, ref.shape[1], ref.shape[0])
            assert np.array_equal(ref, got), (name, tag)


def test_gaussian_matches_cv2(prims):
    for name in ("base", "rnd"):
        assert np.array_equal(prims[f"gauss_{name}"], O.gaussian7(prims[name]))


def test_fast_matches_cv2(prims):
    total = 0
    for ci, (which, x0, y0, cw, ch, thr) in enumerate(prims["fast_cases"]):
        im = prims["base"] if which == 0 else prims["rnd"]
        xs, ys, sc = O.fast9_16_nms(im[y0:y0 + ch, x0:x0 + cw], int(thr))
        got = np.stack([xs, ys, sc], 1).astype(np.int32).reshape(-1, 3)
        assert np.array_equal(prims[f"fast_{ci}"], got), ci
        total += len(got)
    assert total > 500


def test_fast_atan2_matches_cv2(prims):
    got = np.array([O.fast_atan2(y, x) for y, x in prims["atan2_in"]], np.float32)
    assert np.array_equal(got, prims["atan2_out"])
    assert O.fast_atan2(1, 1) == np.float32(44.990456)


def test_trig_tolerance():
    # test/stella_vslam/util/trigonometric.cc:8-20: within 1e-3 of libm over 0..360 deg
    L = O.lib()
    for deg in range(0, 361):
        r = np.float32(deg * np.pi / 180)
        assert abs(L.orc_util_cos(r) - np.cos(r)) < 1e-3
        assert abs(L.orc_util_sin(r) - np.sin(r)) < 1e-3


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "extract_*.npz"))))
def test_extract_matches_cv2_assembly(path):
    g = np.load(path)
    mask = g["mask"] if "mask" in g.files else None
    r = O.orb_extract(g["image"], mask=mask, min_area=int(g["min_area"]), ini_fast_thr=int(g["ini_thr"]),
                      min_fast_thr=int(g["min_thr"]))
    assert r["level_counts"].tolist() == g["level_counts"].tolist()
    assert r["raw_counts"].tolist() == g["raw_counts"].tolist()
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(r["kps"][f], g["kps"][f]), f
    assert np.array_equal(r["desc"], g["desc"])


def test_toy_sample_property(golden_dir):
    # test/stella_vslam/feature/orb_extractor.cc:25-50
    g = np.load(os.path.join(golden_dir, "extract_toy_600.npz"))
    r = O.orb_extract(g["image"], min_area=1000)
    sf = O.scale_factors()[0]
    assert len(r["kps"]) > 0 and r["desc"].shape == (len(r["kps"]), 32) and r["desc"].dtype == np.uint8
    for kp in r["kps"]:
        assert abs(kp["x"] - 300) <= 2.0 * sf[kp["octave"]]
        assert abs(kp["y"] - 300) <= 2.0 * sf[kp["octave"]]


def test_mask_excludes_keypoints(golden_dir):
    # test/stella_vslam/feature/orb_extractor.cc:117-229: no keypoint inside a masked area
    g = np.load(os.path.join(golden_dir, "extract_synth_400x300_mask.npz"))
    r = O.orb_extract(g["image"], mask=g["mask"])
    assert len(r["kps"]) > 0
    for kp in r["kps"]:
        assert g["mask"][int(kp["y"]), int(kp["x"])] != 0


def test_rect_mask_zero_set():
    m = O.rect_mask(400, 300, [[0.0, 0.15, 0.0, 1.0], [0.4, 0.6, 0.4, 0.7]])
    assert (m[:, :61] == 0).all() and (m[:, 61:160] == 255).all()
    assert (m[120:211, 160:241] == 0).all() and m[119, 200] == 255 and m[211, 200] == 255


def test_hamming_kat(golden_dir):
    # test/stella_vslam/match/base.cc:11-57
    g = np.load(os.path.join(golden_dir, "hamming_kat.npz"))
    for a, b, d in zip(g["a"], g["b"], g["dist"]):
        assert O.hamming_32(a, b) == d
        assert O.hamming_64(a, b) == d
