#!/usr/bin/env python3
"""Round-2 golden fixtures (run in the BUILD container only; needs cv2 and /root/reference/test/data).

Adds to make_golden.py (whose cv2-assembled extractor `extract_cv2` is reused unchanged):

  1. NATURAL images: the reference's own test images (test/data/equirectangular_image_00{1,2}.jpg, 1920x960), decoded once with
     cv2 and committed as decoded grey pixels (JPEG decoding is library dependent, the fixture must not be);
  2. end-to-end extractor goldens at the sizes of BASELINE.json's configs:
        natural_1920x960_masks : image_001, the four mask rectangles of example/aist/equirectangular.yaml, min_size 800
        natural_752x480        : EuRoC-sized crop of image_002 (config 0), EuRoC thresholds 20 / 7
        natural_1241x376       : KITTI-sized crop of image_002 (config 3), KITTI thresholds 12 / 7
        equirect_3840x1920_masks : config 2 -- image_001 upsampled 2x by tests/golden/natural.py::upsample2x (pure integer numpy, so
                                 the test can rebuild the input), the same four rectangles; 71k-cell grid => the keypoints are stored
                                 as per-level counts + SHA-256 of the keypoint / descriptor bytes + every 64th row
     The rectangle mask is drawn by the REAL cv2.rectangle(..., LINE_AA) exactly as orb_extractor.cc:138-151 does.
  3. match_bf_cv2.npz: cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) on seeded descriptor sets -- the distance, best and second-best
     that robust::brute_force_match consumes (robust.cc:270-295), from OpenCV's own Hamming kernel.

Usage:  PYTHONPATH=/root/repo python tests/golden/make_golden_r2.py
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import c_round, extract_cv2  # noqa: E402
from natural import AIST_MASK_RECTS, upsample2x  # noqa: E402

REF_DATA = "/root/reference/test/data"


def rect_mask_cv2(cols, rows, rects):
    """orb_extractor::create_rectangle_mask (orb_extractor.cc:138-151) with the real cv2.rectangle."""
    m = np.full((rows, cols), 255, np.uint8)
    for r in rects:
        x0, x1 = c_round(cols * float(np.float32(r[0]))), c_round(cols * float(np.float32(r[1])))
        y0, y1 = c_round(rows * float(np.float32(r[2]))), c_round(rows * float(np.float32(r[3])))
        cv2.rectangle(m, (x0, y0), (x1, y1), 0, -1, cv2.LINE_AA)
    return m


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    im1 = cv2.imread(os.path.join(REF_DATA, "equirectangular_image_001.jpg"), cv2.IMREAD_GRAYSCALE)
    im2 = cv2.imread(os.path.join(REF_DATA, "equirectangular_image_002.jpg"), cv2.IMREAD_GRAYSCALE)
    assert im1.shape == (960, 1920) and im2.shape == (960, 1920)
    crop_euroc = np.ascontiguousarray(im2[300:780, 500:1252])
    crop_kitti = np.ascontiguousarray(im2[330:706, 300:1541])
    np.savez_compressed(os.path.join(HERE, "natural_images.npz"), equirect_1920x960=im1, euroc_752x480=crop_euroc, kitti_1241x376=crop_kitti)

    cases = {
        "natural_1920x960_masks": (im1, AIST_MASK_RECTS, dict(min_area=800)),
        "natural_752x480": (crop_euroc, None, dict(min_area=800)),
        "natural_1241x376": (crop_kitti, None, dict(min_area=800, ini_thr=12, min_thr=7)),
    }
    for name, (im, rects, kw) in cases.items():
        mask = rect_mask_cv2(im.shape[1], im.shape[0], rects) if rects else None
        kps, desc, lc, rc, _ = extract_cv2(im, mask, **kw)
        arrs = dict(kps=kps, desc=desc, level_counts=lc, raw_counts=rc, min_area=kw.get("min_area", 800), ini_thr=kw.get("ini_thr", 20),
                    min_thr=kw.get("min_thr", 7))
        if rects:
            arrs["mask_rects"] = np.array(rects, np.float32)
            arrs["rect_mask_zero_rows"] = np.packbits(mask == 0)   # the zero set of cv2.rectangle(LINE_AA): pins create_rectangle_mask
        np.savez_compressed(os.path.join(HERE, f"nat_{name}.npz"), **arrs)
        print(name, len(kps), lc.tolist(), flush=True)

    big = upsample2x(im1)
    assert big.shape == (1920, 3840)
    mask = rect_mask_cv2(3840, 1920, AIST_MASK_RECTS)
    kps, desc, lc, rc, _ = extract_cv2(big, mask, min_area=800)
    np.savez_compressed(os.path.join(HERE, "nat_equirect_3840x1920_masks.npz"), level_counts=lc, raw_counts=rc, n=len(kps),
                        kps_sha256=digest(kps), desc_sha256=digest(desc), kps_sample=kps[::64], desc_sample=desc[::64],
                        mask_rects=np.array(AIST_MASK_RECTS, np.float32), mask_zero_sha256=digest(mask == 0), min_area=800, ini_thr=20, min_thr=7)
    print("equirect_3840x1920_masks", len(kps), lc.tolist(), flush=True)

    # ---- cv2.BFMatcher pins for the brute-force matcher's inputs ---------------------------------------------------------------
    rng = np.random.default_rng(77)
    out = {}
    for ci, (n1, n2) in enumerate([(300, 280), (1000, 1100), (64, 2000)]):
        d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        src = rng.integers(0, n1, n2)
        near = rng.random(n2) < 0.6
        for j in np.nonzero(near)[0]:                         # 60 % of side 2 = a row of side 1 with 0..40 flipped bits
            row = d1[src[j]].copy()
            bits = rng.choice(256, int(rng.integers(0, 41)), replace=False)
            for b in bits:
                row[b >> 3] ^= np.uint8(1 << (b & 7))
            d2[j] = row
        knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d2, d1, k=2)     # query = keyframe side (robust.cc outer loop), train = frame side
        best = np.array([[m[0].trainIdx, m[0].distance, m[1].trainIdx, m[1].distance] for m in knn], np.int32)
        out[f"d1_{ci}"], out[f"d2_{ci}"], out[f"knn_{ci}"] = d1, d2, best
        if ci == 0:   # the full distance matrix from OpenCV's own Hamming kernel (rows = keyframe side, columns = frame side)
            full = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d2, d1, k=n1)
            D = np.zeros((n2, n1), np.int16)
            for q, ms in enumerate(full):
                for m in ms:
                    D[q, m.trainIdx] = int(m.distance)
            out["dist_0"] = D
    np.savez_compressed(os.path.join(HERE, "match_bf_cv2.npz"), **out)


if __name__ == "__main__":
    main()
