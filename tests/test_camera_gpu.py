"""GPU parity: b200_keypoints_undistort against the (cv2-pinned) oracle: undistorted keypoints bit-exact; perspective bearings
bit-exact (sqrt / divide only), equirectangular bearings within 4e-16 (sin / cos)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import feature
from test_camera_cpu import CAMS, _kps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CAMS))
def test_undistort_and_bearings_parity(name):
    cam = CAMS[name]
    ex = feature.orb_extractor(feature.orb_params(), 800)
    for n in (1, 37, 20000):
        k = _kps(cam, n, 3 + n)
        und, b = ex.undistort_keypoints(cam, k)
        und_w, b_w = O.undistort_keypoints(cam, k)
        assert np.array_equal(und, und_w)
        assert np.array_equal(b, b_w)
    und, b = ex.undistort_keypoints(cam, k[:0])
    assert len(und) == 0 and b.shape == (0, 3)


def test_equirectangular_parity():
    cam = dict(model="equirectangular", cols=3840, rows=1920)
    ex = feature.orb_extractor(feature.orb_params(), 800)
    k = _kps(cam, 8000, 9)
    und, b = ex.undistort_keypoints(cam, k)
    und_w, b_w = O.undistort_keypoints(cam, k)
    assert np.array_equal(und, und_w)
    assert np.allclose(b, b_w, rtol=0, atol=4e-16)
