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


@pytest.mark.gpu
@pytest.mark.parametrize("ch", [3, 4])
def test_gpu_device_batch_both_paths(ch):
    """The batched device entry: 128-bit path (16-byte aligned rows) and the 32-bit path (rows only 4-byte aligned)."""
    import torch
    from stella_vslam_b200 import feature
    from stella_vslam_b200._lib import check, lib
    ex = feature.orb_extractor(feature.orb_params(), 800)
    for w, pad in ((1920, 0), (1000, 0), (1001, 1), (37, 0)):
        b, h = 3, 21
        sp = (w * ch + 3) // 4 * 4 + 4 * pad                     # source pitch: multiple of 4, multiple of 16 only for some widths
        gp = (w + 3) // 4 * 4
        src = torch.randint(0, 256, (b, h, sp), dtype=torch.uint8, device="cuda")
        dst = torch.zeros((b, h, gp), dtype=torch.uint8, device="cuda")
        fs = (h * sp + 15) // 16 * 16
        srcbuf = torch.zeros(b * fs + 16, dtype=torch.uint8, device="cuda")
        for f in range(b):
            srcbuf[f * fs:f * fs + h * sp] = src[f].reshape(-1)
        check(lib().b200_convert_to_grayscale_device(ex._h, srcbuf.data_ptr(), w, h, sp, fs, ch, 1, dst.data_ptr(), gp, h * gp, b))
        torch.cuda.synchronize()
        for f in range(b):
            img = src[f].cpu().numpy()[:, :w * ch].reshape(h, w, ch)
            assert np.array_equal(dst[f].cpu().numpy()[:, :w], O.convert_to_grayscale(img, "RGB")), (w, f)
