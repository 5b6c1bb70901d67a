"""CPU checks of the oracle's tracking-chain composition (oracle.pyoracle.track_local_map) on an oracle-extracted frame: the stages
compose the way tracking_module::search_local_landmarks + pose_optimizer::optimize do."""
import numpy as np

from oracle import pyoracle as O
from stella_vslam_b200 import feature
from workloads import synth

CAM = dict(model="perspective", fx=500.0, fy=500.0, cx=320.0, cy=240.0, fxb=40.0, cols=640.0, rows=480.0)


def _frame(stereo):
    img = synth.make_frame(640, 480, seed=3)
    r = O.orb_extract(img, min_area=800)
    prm = feature.orb_params()
    fr = synth.make_tracking_frame(r["kps"], r["desc"], CAM, prm.scale_factors_, seed=5, stereo=stereo)
    out = O.track_local_map(CAM, r["kps"], r["desc"], fr, prm.scale_factors_, prm.inv_level_sigma_sq_, prm.log_scale_factor_, monocular=not stereo)
    return r, fr, out


def test_oracle_chain_properties():
    for stereo in (False, True):
        r, fr, out = _frame(stereo)
        lm = fr["landmarks"]
        n_kp = len(r["kps"])
        assert out["n_matches"] > 0.3 * n_kp and out["n_valid"] > 0.3 * n_kp
        # skipped landmarks are never searched; the landmarks the frame already carried stay where they were
        assert not out["observable"][lm["skip"].astype(bool)].any()
        pre = fr["kp_landmark"] >= 0
        keep = pre & lm["has_observation"][np.maximum(fr["kp_landmark"], 0)].astype(bool)
        assert np.array_equal(out["kp_landmark"][keep], fr["kp_landmark"][keep])      # occupied keypoints are not offered (projection.cc:50-53)
        # every newly attached landmark was observable and no landmark with observations sits on two keypoints
        new = (out["kp_landmark"] >= 0) & ~keep
        assert out["observable"][out["kp_landmark"][new & ~pre]].all()
        # the optimised pose is closer to the truth than the prior and the flagged observations are a small minority
        assert np.abs(out["pose_cw"] - fr["gt_pose_cw"]).max() < 0.3 * np.abs(fr["pose_cw"] - fr["gt_pose_cw"]).max()
        assert out["kp_outlier"].sum() < 0.1 * (out["kp_landmark"] >= 0).sum()
        assert not out["kp_outlier"][out["kp_landmark"] < 0].any()


def test_oracle_chain_few_observations_leaves_pose():
    r, fr, _ = _frame(False)
    prm = feature.orb_params()
    few = dict(fr, landmarks={k: v[:3] for k, v in fr["landmarks"].items()}, kp_landmark=None)
    out = O.track_local_map(CAM, r["kps"], r["desc"], few, prm.scale_factors_, prm.inv_level_sigma_sq_, prm.log_scale_factor_)
    assert out["n_valid"] == 0 and np.array_equal(out["pose_cw"], fr["pose_cw"])      # pose_optimizer_g2o.cc:116-118
