"""data::frame::can_observe: oracle vs a literal numpy walk (CPU); CUDA vs oracle (GPU): observability flags and
predicted levels identical, perspective reprojections bit-exact, equirectangular ones within 1e-12 (asin / atan2)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from workloads import synth

KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, fxb=386.1448, cols=1241, rows=376)
LOG_SF = np.float32(np.log(np.float32(1.2)))


def _scene(seed, n, equirect=False):
    rng = np.random.default_rng(seed)
    T = np.eye(4)
    T[:3, :3] = synth._rot_y(0.4 * rng.standard_normal()) @ synth._rodrigues(0.1 * rng.standard_normal(3))
    T[:3, 3] = rng.normal(0, 3, 3)
    c = -T[:3, :3].T @ T[:3, 3]
    pos = c + rng.normal(0, 25, (n, 3))                      # all around the camera: behind, outside the image, too far, ...
    front = rng.random(n) < 0.6                              # ... and most of them inside the viewing frustum
    z = rng.uniform(3, 60, n)
    pc = np.stack([(rng.uniform(-100, 1341, n) - KITTI["cx"]) / KITTI["fx"] * z, (rng.uniform(-50, 426, n) - KITTI["cy"]) / KITTI["fy"] * z, z], 1)
    pos[front] = ((pc - T[:3, 3]) @ T[:3, :3])[front]
    d = np.linalg.norm(pos - c, axis=1)
    ref = d * rng.uniform(0.4, 2.5, n)                       # distance at which the landmark was first seen
    nml = (c - pos) / d[:, None] + rng.normal(0, 0.6, (n, 3))
    nml = -nml / np.linalg.norm(nml, axis=1, keepdims=True) * -1.0
    lms = dict(pos_w=pos, mean_normal=-nml, min_valid_dist=(ref / 1.2 ** 7 * 0.9).astype(np.float32), max_valid_dist=(ref * 1.1).astype(np.float32))
    cam = dict(model="equirectangular", cols=3840, rows=1920) if equirect else dict(KITTI)
    return cam, T, lms


def _literal(cam, T, lms, thr=0.5):
    f32 = np.float32
    n = len(lms["pos_w"])
    ok, rp, xr, lv = np.zeros(n, bool), np.zeros((n, 2)), np.zeros(n, f32), np.zeros(n, np.uint32)
    R, t = T[:3, :3], T[:3, 3]
    twc = np.array([-((R[0, r] * t[0] + R[1, r] * t[1]) + R[2, r] * t[2]) for r in range(3)])
    for i in range(n):
        p = lms["pos_w"][i]
        pc = np.array([(R[r, 0] * p[0] + R[r, 1] * p[1]) + R[r, 2] * p[2] + t[r] for r in range(3)])
        if pc[2] <= 0:
            continue
        zi = 1.0 / pc[2]
        x, y = cam["fx"] * pc[0] * zi + cam["cx"], cam["fy"] * pc[1] * zi + cam["cy"]
        if not (0 < x < f32(cam["cols"]) and 0 < y < f32(cam["rows"])):
            continue
        v = p - twc
        dist = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
        df = f32(dist)
        if not (f32(1.0 / 1.3) * lms["min_valid_dist"][i] <= df <= f32(1.3) * lms["max_valid_dist"][i]):
            continue
        nml = lms["mean_normal"][i]
        if ((v[0] * nml[0] + v[1] * nml[1]) + v[2] * nml[2]) / dist < f32(thr):
            continue
        lvl = int(np.ceil(f32(np.log(f32(lms["max_valid_dist"][i] / df))) / LOG_SF))
        ok[i], rp[i], xr[i] = True, (x, y), f32(x - cam["fxb"] * zi)
        lv[i] = 0 if lvl < 0 else (7 if 8 <= lvl else lvl)
    return dict(observable=ok, reproj=rp, x_right=xr, pred_scale_level=lv)


def test_oracle_matches_literal_walk():
    cam, T, lms = _scene(0, 4000)
    got = O.can_observe(cam, T, lms, log_scale_factor=LOG_SF)
    want = _literal(cam, T, lms)
    assert np.array_equal(got["observable"], want["observable"]) and 200 < want["observable"].sum() < 3000
    assert np.array_equal(got["reproj"], want["reproj"]) and np.array_equal(got["x_right"], want["x_right"])
    assert np.array_equal(got["pred_scale_level"], want["pred_scale_level"])
    assert len(np.unique(want["pred_scale_level"][want["observable"]])) >= 6        # the whole pyramid is predicted


@pytest.mark.gpu
@pytest.mark.parametrize("equirect", [False, True])
def test_gpu_matches_oracle(equirect):
    from stella_vslam_b200 import feature
    ex = feature.orb_extractor(feature.orb_params(), 800)
    for seed, n in ((1, 5000), (2, 1), (3, 33)):
        cam, T, lms = _scene(seed, n, equirect)
        got = ex.can_observe(cam, T, lms)
        want = O.can_observe(cam, T, lms, log_scale_factor=ex.orb_params_.log_scale_factor_)
        assert np.array_equal(got["observable"], want["observable"])
        assert np.array_equal(got["pred_scale_level"], want["pred_scale_level"])
        if equirect:
            assert np.allclose(got["reproj"], want["reproj"], rtol=1e-12, atol=1e-9) and (got["x_right"] == 0).all()
        else:
            assert np.array_equal(got["reproj"], want["reproj"]) and np.array_equal(got["x_right"], want["x_right"])
    empty = dict(pos_w=np.zeros((0, 3)), mean_normal=np.zeros((0, 3)), min_valid_dist=np.zeros(0), max_valid_dist=np.zeros(0))
    assert len(ex.can_observe(cam, T, empty)["observable"]) == 0
