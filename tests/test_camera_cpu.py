"""Per-keypoint steps after extraction: the undistortion oracle is PINNED against cv2.undistortPointsIter (4.13) with
the reference's own TermCriteria on the reference's example calibrations (example/euroc/EuRoC_mono.yaml, example/tum_rgbd/
TUM_RGBD_mono_1.yaml, a distortion-free KITTI camera); bearings against their closed form."""
import cv2
import numpy as np
import pytest

from oracle import pyoracle as O

CAMS = {
    "euroc": dict(fx=458.654, fy=457.296, cx=367.215, cy=248.375, k1=-0.28340811, k2=0.07395907, p1=0.00019359, p2=1.76187114e-05, k3=0.0,
                  cols=752, rows=480),
    "tum_rgbd": dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358, p2=0.002628,
                     k3=1.163314, cols=640, rows=480),
    "kitti": dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0, cols=1241, rows=376),
}


def _kps(cam, n, seed):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, O.KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, cam["cols"], n), rng.uniform(0, cam["rows"], n)
    k["x"][: n // 2] = np.round(k["x"][: n // 2])         # extractor output at level 0 is integer valued
    k["y"][: n // 2] = np.round(k["y"][: n // 2])
    k["size"], k["angle"], k["response"], k["octave"] = 31.0, rng.uniform(0, 360, n), rng.uniform(1, 200, n), rng.integers(0, 8, n)
    return k


@pytest.mark.parametrize("name", sorted(CAMS))
def test_undistort_oracle_pinned_to_cv2(name):
    cam = CAMS[name]
    k = _kps(cam, 30000, 1)
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float64)
    d = np.array([cam["k1"], cam["k2"], cam["p1"], cam["p2"], cam["k3"]], np.float64)
    want = cv2.undistortPointsIter(np.stack([k["x"], k["y"]], 1).reshape(-1, 1, 2), K, d, np.eye(3), K,
                                   (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_MAX_ITER, 20, 1e-6)).reshape(-1, 2)
    und, b = O.undistort_keypoints(cam, k)
    assert np.array_equal(und["x"], want[:, 0]) and np.array_equal(und["y"], want[:, 1])      # bit-exact
    assert np.array_equal(und["angle"], k["angle"]) and np.array_equal(und["octave"], k["octave"]) and (und["response"] == 0).all()
    xn, yn = (want[:, 0].astype(np.float64) - cam["cx"]) / cam["fx"], (want[:, 1].astype(np.float64) - cam["cy"]) / cam["fy"]
    ref = np.stack([xn, yn, np.ones_like(xn)], 1) / np.sqrt(xn * xn + yn * yn + 1.0)[:, None]
    assert np.array_equal(b, ref)
    assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-15)


def test_equirectangular_bearings():
    cam = dict(model="equirectangular", cols=3840, rows=1920)
    k = _kps(cam, 5000, 2)
    und, b = O.undistort_keypoints(cam, k)
    assert np.array_equal(und, k)                                              # no undistortion for this model
    # float / unsigned int is a float division in the reference (equirectangular.cc:45-46)
    lon = ((k["x"] / np.float32(3840)).astype(np.float64) - 0.5) * (2 * np.pi)
    lat = -((k["y"] / np.float32(1920)).astype(np.float64) - 0.5) * np.pi
    ref = np.stack([np.cos(lat) * np.sin(lon), -np.sin(lat), np.cos(lat) * np.cos(lon)], 1)
    assert np.allclose(b, ref, rtol=0, atol=1e-15)
