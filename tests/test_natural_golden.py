"""Round-2 pins on NATURAL imagery and at the sizes of BASELINE.json's configs (fixtures: tests/golden/make_golden_r2.py).

The goldens are the keypoints / descriptors of the extractor assembled from the REAL cv2 primitives on the reference's own test
images (test/data/equirectangular_image_00{1,2}.jpg, committed as decoded grey pixels): 1920x960 with the four mask rectangles of
example/aist/equirectangular.yaml, a 752x480 (EuRoC) and a 1241x376 (KITTI, thresholds 12/7) crop, and BASELINE config 2
(3840x1920 + the same rectangles) stored as level counts + SHA-256 + every 64th row.  CPU tests check the C oracle, `-m gpu` tests
the CUDA path through the C ABI (rectangles through the extractor's ctor argument, like the reference).  Bit-exact."""
import hashlib
import os

import numpy as np
import pytest

from golden.natural import AIST_MASK_RECTS, upsample2x

FIELDS = ("x", "y", "size", "angle", "response", "octave")
CASES = [("natural_1920x960_masks", "equirect_1920x960"), ("natural_752x480", "euroc_752x480"), ("natural_1241x376", "kitti_1241x376")]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def images(golden_dir):
    return np.load(os.path.join(golden_dir, "natural_images.npz"))


def kps_as_golden(kps):
    out = np.zeros(len(kps), dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    for f in FIELDS:
        out[f] = kps[f]
    return out


def check_full(kps, desc, g):
    assert len(kps) == len(g["kps"]), (len(kps), len(g["kps"]))
    for f in FIELDS:
        assert np.array_equal(kps[f], g["kps"][f]), f
    assert np.array_equal(desc, g["desc"])


def check_digest(kps, desc, g):
    assert len(kps) == int(g["n"])
    kk = kps_as_golden(kps)
    assert np.array_equal(kk[::64], g["kps_sample"]) and np.array_equal(desc[::64], g["desc_sample"])
    assert digest(kk) == str(g["kps_sha256"]) and digest(desc) == str(g["desc_sha256"])


# ---- CPU: the oracle ----------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,key", CASES)
def test_oracle_natural(golden_dir, images, name, key):
    from oracle import pyoracle as O
    g = np.load(os.path.join(golden_dir, f"nat_{name}.npz"))
    img = images[key]
    mask = None
    if "mask_rects" in g.files:
        mask = O.rect_mask(img.shape[1], img.shape[0], g["mask_rects"].tolist())
        # create_rectangle_mask (orb_extractor.cc:138-151): same zero set as the real cv2.rectangle(LINE_AA)
        assert np.array_equal(np.packbits(mask == 0), g["rect_mask_zero_rows"])
    r = O.orb_extract(img, mask=mask, min_area=int(g["min_area"]), ini_fast_thr=int(g["ini_thr"]), min_fast_thr=int(g["min_thr"]))
    assert r["level_counts"].tolist() == g["level_counts"].tolist() and r["raw_counts"].tolist() == g["raw_counts"].tolist()
    check_full(r["kps"], r["desc"], g)


def test_oracle_equirect_3840x1920_masks(golden_dir, images):
    # BASELINE config 2: 22.8 Mpx pyramid, 8 levels, the four rectangles of example/aist/equirectangular.yaml
    from oracle import pyoracle as O
    g = np.load(os.path.join(golden_dir, "nat_equirect_3840x1920_masks.npz"))
    img = upsample2x(images["equirect_1920x960"])
    assert img.shape == (1920, 3840)
    mask = O.rect_mask(3840, 1920, AIST_MASK_RECTS)
    assert digest(mask == 0) == str(g["mask_zero_sha256"])
    r = O.orb_extract(img, mask=mask, min_area=800)
    assert r["level_counts"].tolist() == g["level_counts"].tolist() and r["raw_counts"].tolist() == g["raw_counts"].tolist()
    check_digest(r["kps"], r["desc"], g)
    for kp in r["kps"][::97]:   # test/stella_vslam/feature/orb_extractor.cc:231-330: no keypoint inside a masked rectangle
        assert mask[int(kp["y"]), int(kp["x"])] != 0


def test_hamming_and_best_two_match_opencv(golden_dir):
    """a9 / the inputs of a10: popcount distances, best and second-best of every keyframe row equal cv2.BFMatcher(NORM_HAMMING)."""
    from oracle import pyoracle as O
    g = np.load(os.path.join(golden_dir, "match_bf_cv2.npz"))
    for ci in range(3):
        d1, d2, knn = g[f"d1_{ci}"], g[f"d2_{ci}"], g[f"knn_{ci}"]
        rows = np.random.default_rng(ci).choice(len(d2), 40, replace=False)
        for q in rows:
            dist = np.array([O.hamming_32(d2[q], d1[t]) for t in range(len(d1))])
            order = np.argsort(dist, kind="stable")
            assert dist[order[0]] == knn[q, 1] and dist[order[1]] == knn[q, 3]
            assert dist[knn[q, 0]] == knn[q, 1] and dist[knn[q, 2]] == knn[q, 3]
    # robust::brute_force_match (robust.cc:232-328) replayed literally on OpenCV's distance matrix
    d1, d2, D = g["d1_0"], g["d2_0"], g["dist_0"].astype(np.int64)
    for lowe in (0.8, 0.95, 0.6):
        taken, match = set(), {}
        for i2 in range(len(d2)):
            best, second, bi = 256, 256, -1
            for i1 in range(len(d1)):
                if i1 in taken:
                    continue
                h = D[i2, i1]
                if h < best:
                    second, best, bi = best, h, i1
                elif h < second:
                    second = h
            if 50 < best or bi < 0 or np.float32(lowe) * np.float32(second) < np.float32(best):
                continue
            match[bi] = i2
            taken.add(bi)
        want = np.array(sorted(match.items()), np.int32).reshape(-1, 2)
        z1, z2 = np.zeros(len(d1), np.float32), np.zeros(len(d2), np.float32)
        got = O.brute_force_match(d1, z1, d2, z2, None, lowe, False)
        assert len(want) > 50 and np.array_equal(got, want), lowe


# ---- GPU: the CUDA path through the C ABI ------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name,key", CASES)
def test_gpu_natural(golden_dir, images, name, key):
    from stella_vslam_b200 import feature
    g = np.load(os.path.join(golden_dir, f"nat_{name}.npz"))
    rects = g["mask_rects"].tolist() if "mask_rects" in g.files else None
    prm = feature.orb_params("golden", 1.2, 8, int(g["ini_thr"]), int(g["min_thr"]))
    ex = feature.orb_extractor(prm, int(g["min_area"]), mask_rects=rects) if rects else feature.orb_extractor(prm, int(g["min_area"]))
    kps, desc = ex.extract(images[key])
    check_full(kps, desc, g)
    ex.close()


@pytest.mark.gpu
def test_gpu_equirect_3840x1920_masks(golden_dir, images):
    from stella_vslam_b200 import feature
    g = np.load(os.path.join(golden_dir, "nat_equirect_3840x1920_masks.npz"))
    img = upsample2x(images["equirect_1920x960"])
    ex = feature.orb_extractor(feature.orb_params(), 800, mask_rects=AIST_MASK_RECTS)
    kps, desc = ex.extract(img)
    check_digest(kps, desc, g)
    # a batch of two frames of this size through the same handle (pyramid arena 2 x 22.8 MB), second frame shifted by a row
    img2 = np.roll(img, 1, axis=0)
    from oracle import pyoracle as O
    kb, db = ex.extract_batch(np.stack([img, img2]))
    check_digest(kb[0], db[0], g)
    ref = O.orb_extract(img2, mask=O.rect_mask(3840, 1920, AIST_MASK_RECTS), min_area=800)
    check_full(kb[1], db[1], dict(kps=ref["kps"], desc=ref["desc"]))
    ex.close()


@pytest.mark.gpu
def test_gpu_hamming_matches_opencv(golden_dir):
    from stella_vslam_b200 import match
    g = np.load(os.path.join(golden_dir, "match_bf_cv2.npz"))
    D = match.hamming_matrix(g["d2_0"], g["d1_0"])
    assert np.array_equal(D.astype(np.int64), g["dist_0"].astype(np.int64))
    for ci in range(3):
        d1, d2, knn = g[f"d1_{ci}"], g[f"d2_{ci}"], g[f"knn_{ci}"]
        D = match.hamming_matrix(d2, d1).astype(np.int64)
        part = np.partition(D, 1, axis=1)
        assert np.array_equal(part[:, 0], knn[:, 1]) and np.array_equal(part[:, 1], knn[:, 3])
        assert np.array_equal(D[np.arange(len(d2)), knn[:, 0]], knn[:, 1])
