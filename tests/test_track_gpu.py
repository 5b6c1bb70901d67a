"""GPU parity of the device-resident tracking chain (b200_track_local_map) against the oracle's stage-by-stage composition
(oracle.pyoracle.track_local_map): landmark slots, observability and outlier flags bit-exact, pose within 1e-5 (BASELINE north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KITTI_CAM = dict(model="perspective", fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, fxb=386.1448, cols=1241.0, rows=376.0)
EUROC_CAM = dict(model="perspective", fx=458.654, fy=457.296, cx=367.215, cy=248.375, k1=-0.28340811, k2=0.07395907, p1=0.00019359, p2=1.76187114e-05,
                 k3=0.0, fxb=0.0, cols=752.0, rows=480.0)


@pytest.fixture(scope="module")
def mods():
    from oracle import pyoracle as O
    from stella_vslam_b200 import feature, tracking
    from workloads import synth
    return O, feature, tracking, synth


def _check(O, ex, tr, camera, kps, descs, frames, monocular, bounds=None, margin=5.0):
    prm = ex.orb_params_
    got = tr.track(frames)
    total = 0
    for f, (fr, g) in enumerate(zip(frames, got)):
        ref = O.track_local_map(camera, kps[fr.get("frame", f)], descs[fr.get("frame", f)], fr, prm.scale_factors_, prm.inv_level_sigma_sq_,
                                prm.log_scale_factor_, margin=margin, monocular=monocular, img_bounds=bounds)
        assert g["n_keypoints"] == ref["n_keypoints"]
        assert np.array_equal(g["observable"], ref["observable"]), f
        assert np.array_equal(g["kp_landmark"], ref["kp_landmark"]), f
        assert g["n_matches"] == ref["n_matches"], f
        assert g["n_valid"] == ref["n_valid"], (f, g["n_valid"], ref["n_valid"])
        assert np.array_equal(g["kp_outlier"], ref["kp_outlier"]), f
        assert np.abs(g["pose_cw"] - ref["pose_cw"]).max() <= 1e-5 * max(1.0, np.abs(ref["pose_cw"]).max()), f
        total += g["n_matches"]
    return total, got


@pytest.mark.parametrize("stereo", [False, True])
def test_chain_vs_oracle_kitti(mods, stereo):
    O, feature, tracking, synth = mods
    imgs = np.stack([synth.make_frame(1241, 376, seed=50 + i) for i in range(3)])
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=3)
    kps, descs = ex.extract_batch(imgs)
    cam = dict(KITTI_CAM, setup="stereo" if stereo else "monocular")
    frames = [dict(synth.make_tracking_frame(kps[i], descs[i], cam, ex.orb_params_.scale_factors_, seed=70 + i, stereo=stereo), frame=i) for i in range(3)]
    tr = tracking.local_map_tracker(ex, cam)
    total, got = _check(O, ex, tr, cam, kps, descs, frames, monocular=not stereo)
    assert total > 0.4 * sum(len(k) for k in kps) * 0.7
    for fr, g in zip(frames, got):                              # the optimised pose moves towards the true one
        e0 = np.abs(fr["pose_cw"] - fr["gt_pose_cw"]).max()
        assert np.abs(g["pose_cw"] - fr["gt_pose_cw"]).max() < e0
    ms = tr.stage_ms()
    assert ms["chain"] > 0


def test_chain_with_distortion_and_frame_subset(mods):
    # EuRoC-like distortion: the undistortion runs inside the chain; frames given out of order and only a subset of the batch
    O, feature, tracking, synth = mods
    imgs = np.stack([synth.make_frame(752, 480, seed=80 + i) for i in range(4)])
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=4)
    kps, descs = ex.extract_batch(imgs)
    und = [O.undistort_keypoints(EUROC_CAM, k)[0] for k in kps]
    bounds = (-30.0, 790.0, -25.0, 510.0)
    frames = [dict(synth.make_tracking_frame(und[i], descs[i], EUROC_CAM, ex.orb_params_.scale_factors_, seed=90 + i), frame=i) for i in (2, 0)]
    tr = tracking.local_map_tracker(ex, EUROC_CAM, margin=10.0, img_bounds=bounds)
    total, _ = _check(O, ex, tr, EUROC_CAM, kps, descs, frames, monocular=True, bounds=bounds, margin=10.0)
    assert total > 200


def test_chain_equirectangular(mods):
    # BASELINE configs 1 / 2: equirectangular camera (identity undistortion, every reprojection inside the image, equirect pose edges)
    O, feature, tracking, synth = mods
    imgs = np.stack([synth.make_frame(1920, 960, seed=40 + i) for i in range(2)])
    ex = feature.orb_extractor(feature.orb_params(), 2500, max_batch=2)
    kps, descs = ex.extract_batch(imgs)
    cam = dict(model="equirectangular", cols=1920.0, rows=960.0, fxb=0.0, setup="monocular")
    frames = [dict(synth.make_tracking_frame(kps[i], descs[i], cam, ex.orb_params_.scale_factors_, seed=45 + i, pixel_sigma=0.7), frame=i) for i in range(2)]
    tr = tracking.local_map_tracker(ex, cam)
    total, got = _check(O, ex, tr, cam, kps, descs, frames, monocular=True)
    assert total > 0.3 * sum(len(k) for k in kps)
    for fr, g in zip(frames, got):
        assert np.abs(g["pose_cw"] - fr["gt_pose_cw"]).max() < np.abs(fr["pose_cw"] - fr["gt_pose_cw"]).max()


def test_chain_degenerate_frames(mods):
    O, feature, tracking, synth = mods
    imgs = np.stack([synth.make_frame(640, 480, seed=3), np.full((480, 640), 90, np.uint8)])    # second frame: no keypoints at all
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=2)
    kps, descs = ex.extract_batch(imgs)
    assert len(kps[1]) == 0
    cam = dict(model="perspective", fx=500.0, fy=500.0, cx=320.0, cy=240.0, fxb=0.0, cols=640.0, rows=480.0)
    f0 = synth.make_tracking_frame(kps[0], descs[0], cam, ex.orb_params_.scale_factors_, seed=5)
    few = dict(f0, landmarks={k: (v[:3] if v is not None else None) for k, v in f0["landmarks"].items()}, kp_landmark=None)  # < 5 edges: pose untouched
    none = dict(f0, landmarks={k: (v[:0] if v is not None else None) for k, v in f0["landmarks"].items()}, kp_landmark=None)
    empty_frame = dict(f0, kp_landmark=None, frame=1)
    frames = [dict(f0, frame=0), dict(few, frame=0), dict(none, frame=0), empty_frame]
    tr = tracking.local_map_tracker(ex, cam)
    _, got = _check(O, ex, tr, cam, kps, descs, frames, monocular=True)
    assert np.array_equal(got[1]["pose_cw"], few["pose_cw"]) and got[1]["n_valid"] == 0
    assert got[3]["n_keypoints"] == 0 and got[3]["n_matches"] == 0


def test_chain_rejects_mismatched_keypoint_arrays(mods):
    O, feature, tracking, synth = mods
    from stella_vslam_b200._lib import B200Error
    img = synth.make_frame(640, 480, seed=3)
    ex = feature.orb_extractor(feature.orb_params(), 800)
    kps, descs = ex.extract_batch(img[None])
    cam = dict(model="perspective", fx=500.0, fy=500.0, cx=320.0, cy=240.0, fxb=0.0, cols=640.0, rows=480.0)
    fr = synth.make_tracking_frame(kps[0], descs[0], cam, ex.orb_params_.scale_factors_, seed=5)
    fr["kp_landmark"] = fr["kp_landmark"][:-3]
    with pytest.raises(B200Error):
        tracking.local_map_tracker(ex, cam).track([fr])
