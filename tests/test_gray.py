"""util::convert_to_grayscale: the oracle is PINNED bit-exactly against cv2.cvtColor (4.13) for the four conversions the
reference uses; the CUDA kernel is compared with the oracle on the GPU."""
import cv2
import numpy as np
import pytest

from oracle import pyoracle as O

CODES = {("BGR", 3): cv2.COLOR_BGR2GRAY, ("RGB", 3): cv2.COLOR_RGB2GRAY, ("BGR", 4): cv2.COLOR_BGRA2GRAY, ("RGB", 4): cv2.COLOR_RGBA2GRAY}


@pytest.mark.parametrize("order,ch", sorted(CODES))
def test_oracle_pinned_to_cv2(order, ch):
    rng = np.random.default_rng(ch)
    for h, w in ((480, 752), (33, 641), (1, 1), (7, 3)):
        img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        assert np.array_equal(O.convert_to_grayscale(img, order), cv2.cvtColor(img, CODES[(order, ch)]))
    a, b = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    ramp = np.stack([a, b, np.full((256, 256), 77)], -1).astype(np.uint8)
    if ch == 4:
        ramp = np.concatenate([ramp, np.full((256, 256, 1), 200, np.uint8)], -1)
    assert np.array_equal(O.convert_to_grayscale(ramp, order), cv2.cvtColor(ramp, CODES[(order, ch)]))


@pytest.mark.gpu
@pytest.mark.parametrize("order,ch", sorted(CODES))
def test_gpu_matches_oracle(order, ch):
    from stella_vslam_b200 import feature
    ex = feature.orb_extractor(feature.orb_params(), 800)
    rng = np.random.default_rng(10 + ch)
    for h, w in ((1080, 1920), (33, 641), (1, 1), (5, 1023)):
        img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        assert np.array_equal(ex.convert_to_grayscale(img, order), O.convert_to_grayscale(img, order))
    gray = rng.integers(0, 256, (10, 10), dtype=np.uint8)
    assert ex.convert_to_grayscale(gray) is gray
