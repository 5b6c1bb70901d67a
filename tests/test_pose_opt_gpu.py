"""GPU parity: b200_pose_optimize (optimize::pose_optimizer_g2o::optimize) against the oracle.  fp64: optimised pose within 1e-5
relative (BASELINE.json north_star tolerance for the BA path), outlier flags and the returned inlier count identical."""
import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import optimize
from workloads import synth

pytestmark = pytest.mark.gpu


def _same(got, want):
    n, pose, flags = got
    n_w, pose_w, flags_w = want
    assert n == n_w
    assert np.array_equal(flags, flags_w)
    assert np.allclose(pose, pose_w, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("model", ["mono", "stereo", "equirect"])
def test_pose_optimizer_parity(model):
    po = optimize.pose_optimizer()
    for seed in range(3):
        pr = synth.make_pose_problem(10 + seed, n_obs=1800, model=model)
        _same(po.optimize(pr), O.pose_optimize(pr))


def test_pose_optimizer_batch_ragged_and_protocols():
    sizes = [2000, 4, 5, 37, 1200, 0, 300]
    probs = [synth.make_pose_problem(30 + k, n_obs=max(n, 1), model=["mono", "stereo", "equirect"][k % 3]) for k, n in enumerate(sizes)]
    for pr, n in zip(probs, sizes):
        if n == 0:
            for key in ("points", "point_fixed", "e_pose", "e_point", "e_cam", "e_obs", "e_inv_sigma_sq", "e_delta"):
                pr[key] = pr[key][:0]
    for cfg in ((2, 2, 10), (0, 4, 10), (4, 0, 10), (1, 1, 3)):
        po = optimize.pose_optimizer(*cfg)
        got = po.optimize_batch(probs)
        for g, pr in zip(got, probs):
            _same(g, O.pose_optimize(pr, *cfg))
    few = got[1]
    assert few[0] == 0 and np.array_equal(few[1], probs[1]["pose_cw"][0])   # < 5 observations: pose untouched (:116-118)


def test_pose_optimizer_heavy_outliers_breaks_early():
    pr = synth.make_pose_problem(50, n_obs=12, outlier_frac=0.7)
    _same(optimize.pose_optimizer().optimize(pr), O.pose_optimize(pr))
