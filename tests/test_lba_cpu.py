"""CPU tests of the local-BA oracle (oracle/lba_oracle.c).  g2o is not available, so the restatement is checked from
independent directions: (1) the analytic Jacobians of the reference's edges against numerical differentiation through an
independent numpy projection + exponential map, (2) one damped Gauss-Newton step of the Schur solver against a dense
numpy normal-equation solve, (3) convergence to the ground truth on noise-free problems, (4) the protocol quirks of
local_bundle_adjuster_g2o.cc (abort flag, gain-threshold stop skipping the second round)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from workloads import synth


def _exp_se3(upd):
    om, up = upd[:3], upd[3:]
    th = np.linalg.norm(om)
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-5:
        R = np.eye(3) + Om + 0.5 * Om @ Om
        V = np.eye(3) + 0.5 * Om + Om @ Om / 6
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, V @ up
    return T


def _project(cam, Tcw, p, stereo):
    pc = Tcw[:3, :3] @ p + Tcw[:3, 3]
    if cam["model"] == 1:
        th, ph = np.arctan2(pc[0], pc[2]), -np.arcsin(pc[1] / np.linalg.norm(pc))
        return np.array([cam["cols"] * (0.5 + th / (2 * np.pi)), cam["rows"] * (0.5 - ph / np.pi)])
    u = cam["fx"] * pc[0] / pc[2] + cam["cx"]
    v = cam["fy"] * pc[1] / pc[2] + cam["cy"]
    return np.array([u, v, u - cam["fxb"] / pc[2]]) if stereo else np.array([u, v])


def _residuals(pr, pose, pts):
    out = []
    for e in range(len(pr["e_pose"])):
        obs = pr["e_obs"][e].astype(np.float64)
        stereo = obs[2] >= 0
        z = _project(pr["cams"][0], pose[pr["e_pose"][e]], pts[pr["e_point"][e]], stereo)
        r = (obs[:3] if stereo else obs[:2]) - z
        out.append(np.sqrt(float(pr["e_inv_sigma_sq"][e])) * r)
    return np.concatenate(out)


def _dense_gn_step(pr, lam):
    """Dense damped Gauss-Newton step with numerical Jacobians in the reference's parametrisation
    (pose: T <- exp(dx) T, dx = [omega, upsilon]; landmark: p <- p + dx), no robust kernel."""
    K, L = len(pr["pose_cw"]), len(pr["points"])
    free_p = [k for k in range(K) if not pr["pose_fixed"][k]]
    n = 6 * len(free_p) + 3 * L

    def state(x):
        pose = pr["pose_cw"].copy()
        for i, k in enumerate(free_p):
            pose[k] = _exp_se3(x[6 * i:6 * i + 6]) @ pose[k]
        pts = pr["points"] + x[6 * len(free_p):].reshape(L, 3)
        return pose, pts

    r0 = _residuals(pr, *state(np.zeros(n)))
    J = np.zeros((len(r0), n))
    h = 1e-6
    for j in range(n):
        d = np.zeros(n)
        d[j] = h
        J[:, j] = (_residuals(pr, *state(d)) - _residuals(pr, *state(-d))) / (2 * h)
    H = J.T @ J + lam * np.eye(n)
    x = np.linalg.solve(H, -J.T @ r0)
    return state(x), float(r0 @ r0)


@pytest.mark.parametrize("model", ["mono", "stereo", "equirect"])
def test_one_step_matches_dense_normal_equations(model):
    pr = synth.make_ba_problem(5, 2, 24, seed=3, model=model, outlier_frac=0.0, min_obs=3, max_obs=5)
    pr["e_robust"] = np.zeros(len(pr["e_pose"]), np.uint8)       # plain least squares
    pr["e_can_be_outlier"] = np.zeros(len(pr["e_pose"]), np.uint8)
    r = O.lba_solve(pr, iters1=1, iters2=0)
    (pose, pts), chi0 = _dense_gn_step(pr, r["lambda_init"])
    # the oracle's first LM step (lambda = tau * max diag) must equal the dense solve (if it was accepted)
    assert r["iterations"][0] == 1
    assert np.abs(r["points"] - pts).max() < 2e-6 * max(1.0, np.abs(pts).max())
    assert np.abs(r["pose_cw"] - pose).max() < 2e-6 * max(1.0, np.abs(pose).max())


@pytest.mark.parametrize("model", ["mono", "stereo", "equirect"])
def test_converges_to_ground_truth_without_noise(model):
    pr = synth.make_ba_problem(8, 3, 150, seed=7, model=model, outlier_frac=0.0, pixel_sigma=0.0)
    r = O.lba_solve(pr, iters1=20, iters2=20)
    assert r["n_outliers"] == 0
    assert r["chi2"][1] < 1e-3 * max(r["chi2"][0], 1e-9) or r["chi2"][1] < 1e-6
    if model == "stereo":   # scale observable: landmarks return to the truth (obs are stored as float32)
        assert np.percentile(np.abs(r["points"] - pr["gt_points"]), 95) < 5e-3   # (a few far, narrow-baseline landmarks stay loose)


def test_outliers_are_flagged_and_stats_consistent():
    pr = synth.make_ba_problem(12, 3, 600, seed=1, model="stereo")
    r = O.lba_solve(pr)
    assert r["iterations"][0] >= 1 and r["n_outliers"] == int(r["outliers"].sum()) > 0
    # gross outliers were planted on ~5 % of the observations
    assert 0.02 * len(pr["e_pose"]) < r["n_outliers"] < 0.2 * len(pr["e_pose"])
    assert r["chi2"][1] < r["chi2"][0]
    # fixed keyframes are returned untouched
    fixed = pr["pose_fixed"].astype(bool)
    assert np.array_equal(r["pose_cw"][fixed], pr["pose_cw"][fixed])


def test_force_stop_semantics():
    pr = synth.make_ba_problem(8, 3, 200, seed=2, model="mono")
    flag = np.array([1], np.uint8)          # set before the first solve: no write-back (local_bundle_adjuster_g2o.cc:308-310)
    r = O.lba_solve(pr, force_stop=flag)
    assert r["rc"] == 1
    # gain-threshold stop in round 1 writes the caller's flag and therefore skips round 2 (:317-321)
    flag = np.array([0], np.uint8)
    r = O.lba_solve(pr, iters1=50, iters2=10, force_stop=flag)
    assert r["iterations"][0] < 50 and flag[0] == 1 and r["iterations"][1] == 0


def test_fixed_points_and_non_outlier_edges():
    pr = synth.make_ba_problem(6, 2, 80, seed=9, model="mono")
    pf = np.zeros(80, np.uint8)
    pf[:10] = 1                               # marker corners of a keep_fixed_ marker
    pr["point_fixed"] = pf
    co = np.ones(len(pr["e_pose"]), np.uint8)
    co[np.isin(pr["e_point"], np.arange(10))] = 0
    pr["e_can_be_outlier"] = co
    r = O.lba_solve(pr)
    assert np.array_equal(r["points"][:10], pr["points"][:10])
    assert r["outliers"][co == 0].sum() == 0
