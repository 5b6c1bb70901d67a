"""GPU parity tests (run with -m gpu on the B200 box): the CUDA extractor, called through the C ABI, against
(a) the committed cv2 golden fixtures and (b) the CPU oracle on seeded inputs.  Bit-exact: integer/byte work and the
explicitly-rounded fp32 pieces must match in every bit."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ("x", "y", "size", "angle", "response", "octave")


@pytest.fixture(scope="module")
def mods():
    from oracle import pyoracle as O
    from stella_vslam_b200 import feature
    from workloads import synth
    return O, feature, synth


def assert_same(kps, desc, ref):
    assert len(kps) == len(ref["kps"]), (len(kps), len(ref["kps"]))
    for f in FIELDS:
        assert np.array_equal(kps[f], ref["kps"][f]), f
    assert np.array_equal(desc, ref["desc"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "extract_*.npz"))))
def test_golden_cv2_assembly(mods, path):
    O, feature, synth = mods
    g = np.load(path)
    prm = feature.orb_params("golden", 1.2, 8, int(g["ini_thr"]), int(g["min_thr"]))
    ex = feature.orb_extractor(prm, int(g["min_area"]))
    mask = g["mask"] if "mask" in g.files else None
    kps, desc = ex.extract(g["image"], mask)
    assert_same(kps, desc, dict(kps=g["kps"], desc=g["desc"]))
    ex.close()


def test_toy_sample_property(mods, golden_dir):
    # test/stella_vslam/feature/orb_extractor.cc:25-50
    O, feature, synth = mods
    g = np.load(os.path.join(golden_dir, "extract_toy_600.npz"))
    prm = feature.orb_params("ORB setting for test")
    ex = feature.orb_extractor(prm, 1000)
    kps, desc = ex.extract(g["image"])
    assert len(kps) > 0 and desc.shape == (len(kps), 32) and desc.dtype == np.uint8
    for kp in kps:
        assert abs(kp["x"] - 300) <= 2.0 * prm.scale_factors_[kp["octave"]]
        assert abs(kp["y"] - 300) <= 2.0 * prm.scale_factors_[kp["octave"]]


@pytest.mark.parametrize("w,h,seed,min_area", [(752, 480, 1, 800), (1241, 376, 2, 800), (640, 480, 3, 2000), (333, 217, 4, 300)])
def test_vs_oracle_sizes(mods, w, h, seed, min_area):
    O, feature, synth = mods
    img = synth.make_frame(w, h, seed=seed)
    ex = feature.orb_extractor(feature.orb_params(), min_area)
    kps, desc = ex.extract(img)
    ref = O.orb_extract(img, min_area=min_area, want_pyramid=True)
    assert_same(kps, desc, ref)
    # image_pyramid_ (orb_extractor.h:71) is part of the surface: match::stereo reads it
    pyr = ex.image_pyramid()
    for a, b in zip(pyr, ref["pyramid"]):
        assert np.array_equal(a, b)


def test_vs_oracle_1080p_batch(mods):
    O, feature, synth = mods
    frames = np.stack([synth.make_frame(1920, 1080, seed=10 + i, shift=(3 * i, -2 * i)) for i in range(3)])
    ex = feature.orb_extractor(feature.orb_params(), 7000, max_batch=3)
    kps, desc = ex.extract_batch(frames)
    for f in range(3):
        ref = O.orb_extract(frames[f], min_area=7000)
        assert 1500 < len(ref["kps"]) < 2600
        assert_same(kps[f], desc[f], ref)


def test_random_noise_and_low_contrast(mods):
    # dense corners (every cell full) and a flat image (every cell empty -> min-threshold retry -> still empty)
    O, feature, synth = mods
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    flat = np.full((300, 400), 127, np.uint8)
    faint = (127 + 6 * (rng.random((300, 400)) > 0.98)).astype(np.uint8)   # only the min threshold (7) can fire... or nothing
    grad = (np.add.outer(np.arange(300), np.arange(400)) % 256).astype(np.uint8)
    ex = feature.orb_extractor(feature.orb_params(), 400)
    for img in (noise, flat, faint, grad):
        kps, desc = ex.extract(img)
        assert_same(kps, desc, O.orb_extract(img, min_area=400))
    kps, desc = ex.extract(flat)
    assert len(kps) == 0 and desc.shape == (0, 32)


def test_thresholds_and_levels(mods):
    O, feature, synth = mods
    img = synth.make_frame(500, 400, seed=21)
    for (sf, nl, ini, mn, area) in [(1.2, 8, 12, 7, 800), (1.5, 4, 30, 5, 500), (1.1, 12, 20, 7, 1000), (1.2, 1, 20, 20, 800)]:
        ex = feature.orb_extractor(feature.orb_params("t", sf, nl, ini, mn), area)
        kps, desc = ex.extract(img)
        assert_same(kps, desc, O.orb_extract(img, scale_factor=sf, num_levels=nl, ini_fast_thr=ini, min_fast_thr=mn, min_area=area))
        ex.close()


def test_masks(mods):
    # test/stella_vslam/feature/orb_extractor.cc:117-330: image masks and rectangle masks exclude keypoints
    O, feature, synth = mods
    img = synth.make_frame(640, 480, seed=31)
    mask = np.full((480, 640), 255, np.uint8)
    mask[:, :100] = 0
    mask[200:300, 300:500] = 0
    yy, xx = np.mgrid[0:480, 0:640]
    mask[(yy - 380) ** 2 + (xx - 150) ** 2 < 60 ** 2] = 0
    ex = feature.orb_extractor(feature.orb_params(), 800)
    kps, desc = ex.extract(img, mask)
    assert_same(kps, desc, O.orb_extract(img, mask=mask))
    assert len(kps) > 50
    for kp in kps:
        assert mask[int(kp["y"]), int(kp["x"])] != 0
    # rectangle masks through the ctor (equirectangular.yaml style fractions)
    rects = [[0.0, 0.2, 0.0, 1.0], [0.5, 0.7, 0.3, 0.6]]
    ex2 = feature.orb_extractor(feature.orb_params(), 800, mask_rects=rects)
    kps2, desc2 = ex2.extract(img)
    rmask = O.rect_mask(640, 480, rects)
    assert_same(kps2, desc2, O.orb_extract(img, mask=rmask))
    # an explicit image mask takes precedence over the rectangle mask (orb_extractor.cc:50-60)
    kps3, desc3 = ex2.extract(img, mask)
    assert_same(kps3, desc3, O.orb_extract(img, mask=mask))


def test_empty_and_tiny_inputs(mods):
    O, feature, synth = mods
    ex = feature.orb_extractor(feature.orb_params(), 800)
    kps, desc = ex.extract(np.zeros((0, 0), np.uint8))          # orb_extractor.cc:30-32
    assert len(kps) == 0
    small = synth.make_frame(120, 90, seed=2)                  # upper levels fall below the 19-px border
    kps, desc = ex.extract(small)
    assert_same(kps, desc, O.orb_extract(small))


def test_reconfigure_between_sizes(mods):
    O, feature, synth = mods
    ex = feature.orb_extractor(feature.orb_params(), 800)
    for (w, h) in [(320, 240), (400, 300), (320, 240)]:
        img = synth.make_frame(w, h, seed=w)
        kps, desc = ex.extract(img)
        assert_same(kps, desc, O.orb_extract(img))


def test_capacity_error(mods):
    O, feature, synth = mods
    from stella_vslam_b200 import _lib
    img = synth.make_frame(320, 240, seed=11)
    ex = feature.orb_extractor(feature.orb_params(), 800)
    with pytest.raises(_lib.B200Error) as e:
        ex.extract_batch(img[None], cap=10)
    assert e.value.code == _lib.ERR_CAPACITY


def test_determinism_and_idempotence(mods):
    O, feature, synth = mods
    img = synth.make_frame(1920, 1080, seed=77)
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=4)
    k1, d1 = ex.extract_batch(np.stack([img] * 4))
    for f in range(1, 4):
        assert np.array_equal(k1[0], k1[f]) and np.array_equal(d1[0], d1[f])
    k2, d2 = ex.extract(img)
    assert np.array_equal(k1[0], k2) and np.array_equal(d1[0], d2)
    # size-independent properties at the full size: keypoints inside the border, octave range, at most one per grid cell
    assert (k2["octave"] >= 0).all() and (k2["octave"] < 8).all()
    assert len(k2) <= ex_max(ex, 1920, 1080)


def ex_max(ex, w, h):
    from stella_vslam_b200._lib import lib
    return lib().b200_orb_max_keypoints(ex._h, w, h)


def test_device_frames_unaligned_pitch_uses_non_tma_path(mods):
    """Caller frames whose pitch is not a multiple of 16 cannot be described by a tensor map: the extractor falls back to
    ordinary loads for level 0 and must produce the same bits."""
    import ctypes as C

    import torch

    O, feature, synth = mods
    from stella_vslam_b200._lib import KP_DTYPE, check, lib
    L = lib()
    w, h = 333, 217
    img = synth.make_frame(w, h, seed=4)
    ex = feature.orb_extractor(feature.orb_params(), 300)
    d = torch.from_numpy(img).cuda()                     # pitch 333
    check(L.b200_orb_extract_device(ex._h, C.c_void_p(d.data_ptr()), w, h, w, w * h, 1, None, 0))
    cap = L.b200_orb_max_keypoints(ex._h, w, h)
    kps, desc, cnt = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8), np.zeros(1, np.int32)
    check(L.b200_orb_fetch(ex._h, kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, cnt.ctypes.data_as(C.c_void_p)))
    ref = O.orb_extract(img, min_area=300)
    assert_same(kps[:cnt[0]], desc[:cnt[0]], ref)
