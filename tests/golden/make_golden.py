#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ (run in the BUILD container only; needs cv2).

What this pins
--------------
The reference's ORB path delegates its bit-exact arithmetic to OpenCV (cv::resize, cv::FAST, cv::GaussianBlur,
cv::fastAtan2 -- orb_extractor.cc:103,160,228-235, orb_impl.cc:90).  OpenCV's C++ headers are not installed, so the
reference cannot be compiled here, but the cv2 4.13.0 wheel exposes the same four primitives.  This script

  1. records cv2's outputs for those primitives on seeded inputs  -> prims_*.npz  (pins oracle + CUDA primitives), and
  2. assembles the whole extractor in Python from the REAL cv2 primitives plus an independent numpy restatement of
     the reference's glue (cells/retry/mask, grid arg-max, ic_angle, util::cos/sin, rBRIEF, scale correction)
     -> extract_*.npz (pins the oracle's and the CUDA path's end-to-end keypoints + descriptors),
  3. records the Hamming known-answer vectors of the reference's own unit test
     (test/stella_vslam/match/base.cc:11-57) -> hamming_kat.npz.

Usage:  PYTHONPATH=/root/repo python tests/golden/make_golden.py
"""
import math
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from workloads import synth  # noqa: E402

f32 = np.float32


def load_pattern():
    txt = open(os.path.join(ROOT, "stella_vslam_b200", "csrc", "orb_pattern.inc")).read()
    nums = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.split(",") if t.strip()]
    assert len(nums) == 1024
    return np.array(nums, np.float32).reshape(256, 4)


PATTERN = load_pattern()
U_MAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def scale_factors(scale_factor, n):
    sf = [f32(1.0)]
    for _ in range(1, n):
        sf.append(f32(scale_factor) * sf[-1])
    return sf


def c_round(v):  # std::round (half away from zero) on a double
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def util_cos(v):
    PI = f32(3.14159265358979)
    PI_2 = f32(PI / f32(2.0))
    TWO_PI = f32(f32(2.0) * PI)
    INV_TWO_PI = f32(f32(1.0) / TWO_PI)
    THREE_PI_2 = f32(f32(3.0) * PI_2)

    def poly(x):
        x2 = f32(x * x)
        return f32(f32(0.99940307) + f32(x2 * f32(f32(-0.49558072) + f32(f32(0.03679168) * x2))))

    v = f32(v)
    v = f32(v - f32(f32(math.floor(f32(v * INV_TWO_PI))) * TWO_PI))
    v = v if f32(0.0) < v else f32(-v)
    if v < PI_2:
        return poly(v)
    if v < PI:
        return f32(-poly(f32(PI - v)))
    if v < THREE_PI_2:
        return f32(-poly(f32(v - PI)))
    return poly(f32(TWO_PI - v))


def util_sin(v):
    PI_2 = f32(f32(3.14159265358979) / f32(2.0))
    return util_cos(f32(PI_2 - f32(v)))


def ic_angle(img, x, y):
    m01 = m10 = 0
    for u in range(-15, 16):
        m10 += u * int(img[y, x + u])
    for v in range(1, 16):
        d = U_MAX[v]
        vs = 0
        for u in range(-d, d + 1):
            p, m = int(img[y + v, x + u]), int(img[y - v, x + u])
            vs += p - m
            m10 += u * (p + m)
        m01 += v * vs
    return f32(cv2.fastAtan2(float(m01), float(m10)))


def rbrief(blur, x, y, angle_deg):
    angle = f32(float(angle_deg) * math.pi / 180.0)
    ca, sa = util_cos(angle), util_sin(angle)
    px, py, qx, qy = PATTERN[:, 0], PATTERN[:, 1], PATTERN[:, 2], PATTERN[:, 3]

    def rnd(a):  # cvRound: half to even
        return np.rint(a).astype(np.int64)

    with np.errstate(all="ignore"):
        r0 = rnd((px * sa).astype(f32) + (py * ca).astype(f32))
        c0 = rnd((px * ca).astype(f32) - (py * sa).astype(f32))
        r1 = rnd((qx * sa).astype(f32) + (qy * ca).astype(f32))
        c1 = rnd((qx * ca).astype(f32) - (qy * sa).astype(f32))
    bits = (blur[y + r0, x + c0] < blur[y + r1, x + c1]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)


def extract_cv2(img, mask=None, scale_factor=1.2, num_levels=8, ini_thr=20, min_thr=7, min_area=800):
    """orb_extractor::extract assembled from cv2 primitives (orb_extractor.cc:28-136), non-OpenMP order."""
    h, w = img.shape
    sf = scale_factors(scale_factor, num_levels)
    pyr = [img]
    for l in range(1, num_levels):
        scale = float(sf[l])
        size = (c_round(w * 1.0 / scale), c_round(h * 1.0 / scale))
        pyr.append(cv2.resize(pyr[l - 1], size, interpolation=cv2.INTER_LINEAR))
    min_area_sqrt = int(math.sqrt(min_area))
    det_ini = cv2.FastFeatureDetector_create(int(ini_thr), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(int(min_thr), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def in_mask(y, x, s):  # unsigned y, x; float products truncated (orb_extractor.cc:168-170)
        return mask[int(f32(f32(y) * s)), int(f32(f32(x) * s))] == 0

    kps, descs, level_counts, raw_counts = [], [], [], []
    for l in range(num_levels):
        im = pyr[l]
        H, W = im.shape
        s = sf[l]
        B, overlap, cell = 19, 6, 64
        max_bx, max_by = W - B, H - B
        width, height = max_bx - B, max_by - B
        ncols, nrows = width // cell + 1, height // cell + 1
        cand = []
        for i in range(nrows):
            min_y = B + i * cell
            if max_by - overlap <= min_y:
                continue
            max_y = min(min_y + cell + overlap, max_by)
            for j in range(ncols):
                min_x = B + j * cell
                if max_bx - overlap <= min_x:
                    continue
                max_x = min(min_x + cell + overlap, max_bx)
                if mask is not None and (in_mask(min_y, min_x, s) or in_mask(max_y, min_x, s) or in_mask(min_y, max_x, s)
                                         or in_mask(max_y, max_x, s)):
                    continue
                sub = im[min_y:max_y, min_x:max_x]
                found = det_ini.detect(sub)
                if len(found) == 0:
                    found = det_min.detect(sub)
                for k in found:
                    px, py = f32(k.pt[0]) + f32(j * cell), f32(k.pt[1]) + f32(i * cell)
                    if mask is not None and in_mask(int(f32(B) + py), int(f32(B) + px), s):
                        continue
                    cand.append((px, py, f32(k.response)))
        raw_counts.append(len(cand))
        # distribute_keypoints (orb_extractor.cc:289-329)
        smas = float(f32(f32(min_area_sqrt) / s))
        nx, ny = int(math.ceil((max_bx - B) / smas)), int(math.ceil((max_by - B) / smas))
        dx, dy = float(max_bx - B) / nx, float(max_by - B) / ny
        win = {}
        for (px, py, r) in cand:
            idx = int(float(px) / dx) + int(float(py) / dy) * nx
            if idx not in win or r > win[idx][2]:
                win[idx] = (px, py, r)
        level = []
        for idx in sorted(win):
            px, py, r = win[idx]
            x, y = f32(px + f32(B)), f32(py + f32(B))
            ang = ic_angle(im, int(np.rint(x)), int(np.rint(y)))
            level.append([x, y, f32(int(f32(31) * s)), ang, r, l])
        level_counts.append(len(level))
        if level:
            blur = cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            for kp in level:
                descs.append(rbrief(blur, int(np.rint(kp[0])), int(np.rint(kp[1])), kp[3]))
                if l > 0:
                    kp[0], kp[1] = f32(kp[0] * s), f32(kp[1] * s)
            kps += level
    n = len(kps)
    out = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                             ("octave", "<i4")])
    for i, kp in enumerate(kps):
        out[i] = tuple(kp)
    return out, (np.stack(descs) if n else np.zeros((0, 32), np.uint8)), np.array(level_counts), np.array(raw_counts), pyr


def toy_image():
    """test/stella_vslam/feature/orb_extractor.cc:25-50: white 600x600 with a black rectangle, corner at (300,300)."""
    img = np.full((600, 600), 255, np.uint8)
    cv2.rectangle(img, (300, 300), (600, 600), 0, -1, cv2.LINE_AA)
    return img


def main():
    rng = np.random.default_rng(2024)
    # ---- 1. primitives -------------------------------------------------------------------------------------------
    base = synth.make_frame(400, 300, seed=5)
    rnd = rng.integers(0, 256, (211, 317), dtype=np.uint8)
    prim = {"base": base, "rnd": rnd}
    for name, im in (("base", base), ("rnd", rnd)):
        h, w = im.shape
        for tag, (dw, dh) in (("a", (c_round(w / 1.2), c_round(h / 1.2))), ("b", (w // 2 + 3, h // 3 + 1)), ("c", (w + 5, h + 9))):
            prim[f"resize_{name}_{tag}"] = cv2.resize(im, (dw, dh), interpolation=cv2.INTER_LINEAR)
        prim[f"gauss_{name}"] = cv2.GaussianBlur(im, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    fast_cases = []
    for ci in range(24):
        im = base if ci % 2 == 0 else rnd
        h, w = im.shape
        cw, ch = int(rng.integers(7, 71)), int(rng.integers(7, 71))
        x0, y0 = int(rng.integers(0, w - cw)), int(rng.integers(0, h - ch))
        thr = [20, 7, 12][ci % 3]
        det = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        found = det.detect(im[y0:y0 + ch, x0:x0 + cw])
        fast_cases.append((ci % 2, x0, y0, cw, ch, thr))
        prim[f"fast_{ci}"] = np.array([(k.pt[0], k.pt[1], k.response) for k in found], np.int32).reshape(-1, 3)
    prim["fast_cases"] = np.array(fast_cases, np.int32)
    yx = rng.integers(-60000, 60000, (4000, 2)).astype(np.float32)
    yx[:8] = [(0, 0), (1, 1), (0, 5), (5, 0), (-5, 0), (0, -5), (3, 3), (-3, -3)]
    prim["atan2_in"] = yx
    prim["atan2_out"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "prims_cv2.npz"), **prim)

    # ---- 2. assembled extractor ------------------------------------------------------------------------------------
    cases = {
        "synth_320x240": (synth.make_frame(320, 240, seed=11), None, dict(min_area=800)),
        "synth_400x300_thr12": (synth.make_frame(400, 300, seed=12), None, dict(min_area=400, ini_thr=12, min_thr=5)),
        "toy_600": (toy_image(), None, dict(min_area=1000)),
    }
    m = np.full((300, 400), 255, np.uint8)
    m[:, :60] = 0
    m[120:200, 150:260] = 0
    cases["synth_400x300_mask"] = (synth.make_frame(400, 300, seed=13), m, dict(min_area=800))
    for name, (im, mask, kw) in cases.items():
        kps, desc, lc, rc, _ = extract_cv2(im, mask, **kw)
        arrs = dict(kps=kps, desc=desc, level_counts=lc, raw_counts=rc, min_area=kw.get("min_area", 800),
                    ini_thr=kw.get("ini_thr", 20), min_thr=kw.get("min_thr", 7))
        arrs["image"] = im
        if mask is not None:
            arrs["mask"] = mask
        np.savez_compressed(os.path.join(HERE, f"extract_{name}.npz"), **arrs)
        print(name, len(kps), lc.tolist())

    # ---- 3. Hamming KATs (test/stella_vslam/match/base.cc:11-57) -----------------------------------------------------
    a = np.zeros((3, 32), np.uint8)
    b = np.zeros((3, 32), np.uint8)
    a[0], b[0] = 0b01010101, 0b01010101   # identical -> 0
    a[1], b[1] = 0b01010101, 0b10101010   # complement -> 256
    a[2], b[2] = 0b01100110, 0b00111100   # -> 128
    np.savez_compressed(os.path.join(HERE, "hamming_kat.npz"), a=a, b=b, dist=np.array([0, 256, 128], np.int32))


if __name__ == "__main__":
    main()
