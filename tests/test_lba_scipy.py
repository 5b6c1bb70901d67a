"""Pins the local-BA ORACLE from outside ("parity unpinned" for a20-a27): an independent numpy restatement of the three
reprojection residuals + scipy.optimize.least_squares.

  1. chi2 pin: the robust / plain chi2 the oracle reports after each round equals the cost of ITS OWN output state evaluated by the
     independent numpy residual functions (perspective mono / stereo: perspective_reproj_edge.h:118-120, 236-239; equirectangular:
     equirectangular_reproj_edge.h:130-134; information = inv_sigma_sq * I, Huber cost rho(chi2) of g2o's RobustKernelHuber);
  2. optimum pin: starting scipy's trust-region solver at the oracle's result lowers the cost by no more than the gain threshold the
     reference stops on (terminate_action.cc:36-76, 1e-3 relative): the oracle's LM + Schur + outlier protocol ends at the optimum
     of the same least-squares problem (second round: plain squares on the inlier set; first round: Huber);
  3. outlier pin: the flags equal the chi-square test at the round-1 state re-evaluated in numpy.

The LM path itself (iteration counts, lambda) has no external reference; this file pins what the optimisation converges TO."""
import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation as R

from oracle import pyoracle as O
from workloads import synth


class Problem:
    def __init__(self, pr):
        self.pr = pr
        self.K, self.L = len(pr["pose_cw"]), len(pr["points"])
        self.free_k = np.nonzero(np.asarray(pr["pose_fixed"]) == 0)[0]
        self.cam = pr["cams"][0]
        self.T0 = np.asarray(pr["pose_cw"], np.float64).copy()

    def unpack(self, x):
        T = self.T0.copy()
        for n, k in enumerate(self.free_k):   # T <- exp(xi) T0 (any smooth parametrisation has the same optimum)
            xi = x[6 * n:6 * n + 6]
            dR = R.from_rotvec(xi[:3]).as_matrix()
            T[k, :3, :3] = dR @ self.T0[k, :3, :3]
            T[k, :3, 3] = dR @ self.T0[k, :3, 3] + xi[3:]
        return T, x[6 * len(self.free_k):].reshape(self.L, 3)

    def pack(self, pose, pts):
        x = []
        for k in self.free_k:
            dR = pose[k, :3, :3] @ self.T0[k, :3, :3].T
            x.append(np.concatenate([R.from_matrix(dR).as_rotvec(), pose[k, :3, 3] - dR @ self.T0[k, :3, 3]]))
        return np.concatenate(x + [np.asarray(pts, np.float64).ravel()])

    def edge_chi2(self, T, P, sel):
        pr, cam = self.pr, self.cam
        ep, el, obs = pr["e_pose"][sel], pr["e_point"][sel], pr["e_obs"][sel].astype(np.float64)
        pc = np.einsum("eij,ej->ei", T[ep, :3, :3], P[el]) + T[ep, :3, 3]
        if cam["model"] == 1:
            th = np.arctan2(pc[:, 0], pc[:, 2])
            ph = -np.arcsin(pc[:, 1] / np.linalg.norm(pc, axis=1))
            e = np.stack([obs[:, 0] - cam["cols"] * (0.5 + th / (2 * np.pi)), obs[:, 1] - cam["rows"] * (0.5 - ph / np.pi)], 1)
        else:
            u = cam["fx"] * pc[:, 0] / pc[:, 2] + cam["cx"]
            v = cam["fy"] * pc[:, 1] / pc[:, 2] + cam["cy"]
            xr = np.where(obs[:, 2] >= 0, obs[:, 2] - (u - cam["fxb"] / pc[:, 2]), 0.0)
            e = np.stack([obs[:, 0] - u, obs[:, 1] - v, xr], 1)
        return (e * e).sum(1) * pr["e_inv_sigma_sq"][sel].astype(np.float64), pc

    def huber(self, chi2, sel):
        d = self.pr["e_delta"][sel].astype(np.float64)
        return np.where(chi2 <= d * d, chi2, 2 * np.sqrt(chi2) * d - d * d)


CASES = [("mono", 10, 3, 150, 5), ("stereo", 10, 3, 150, 6), ("equirect", 8, 2, 120, 7)]


@pytest.mark.parametrize("model,K,F,L,seed", CASES)
def test_chi2_outliers_and_optimum(model, K, F, L, seed):
    pr = synth.make_ba_problem(K, F, L, seed=seed, model=model, min_obs=4, max_obs=7)
    E = len(pr["e_pose"])
    every = np.ones(E, bool)
    pb = Problem(pr)
    # ---- first round alone (Huber on every edge): reported robust chi2 == numpy cost of the reported state
    r1 = O.lba_solve(pr, iters1=40, iters2=0)
    T1, P1 = r1["pose_cw"], r1["points"]
    chi_1, pc_1 = pb.edge_chi2(T1, P1, every)
    assert abs(pb.huber(chi_1, every).sum() - r1["chi2"][0]) <= 1e-9 * r1["chi2"][0]
    # optimum of the Huber problem: scipy with one scalar residual sqrt(chi2_e) per edge and loss='huber', f_scale = delta
    delta = float(pr["e_delta"][0])
    assert np.all(pr["e_delta"] == pr["e_delta"][0])

    def fun_rob(x):
        T, P = pb.unpack(x)
        return np.sqrt(pb.edge_chi2(T, P, every)[0] + 1e-300)

    x1 = pb.pack(T1, P1)
    sol = least_squares(fun_rob, x1, loss="huber", f_scale=delta, method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-10, max_nfev=60)
    cost_opt = 2.0 * sol.cost     # scipy: cost = 0.5 * sum f_scale^2 rho((f / f_scale)^2) = 0.5 * sum Huber(chi2)
    assert cost_opt <= r1["chi2"][0] * (1 + 1e-12)
    assert (r1["chi2"][0] - cost_opt) <= 2e-3 * cost_opt, (r1["chi2"][0], cost_opt)
    # ---- the full protocol: outlier flags == chi-square test + depth test at the state after the LAST activation of each edge
    ref = O.lba_solve(pr, iters1=5, iters2=10)
    keep = ref["outliers"] == 0
    chi_f, pc_f = pb.edge_chi2(ref["pose_cw"], ref["points"], every)
    thr = np.where(pr["e_obs"][:, 2] >= 0, np.float64(np.float32(7.81473)), np.float64(np.float32(5.99146)))
    if model == "equirect":
        thr[:] = np.float64(np.float32(5.99146))
    depth_bad = (pc_f[:, 2] <= 0) if model != "equirect" else np.zeros(E, bool)
    # inliers were active in round 2, so their chi2 at the final state decides; every inlier must pass, and any edge that fails at the
    # final state must be flagged
    assert np.all((chi_f[keep] <= thr[keep]) & ~depth_bad[keep])
    assert np.all(ref["outliers"][(chi_f > thr) | depth_bad] == 1)
    # second round: plain least squares on the edges that stayed active; chi2 pin + optimum pin
    r5 = O.lba_solve(pr, iters1=5, iters2=0)
    chi_5, pc_5 = pb.edge_chi2(r5["pose_cw"], r5["points"], every)
    active = ~((chi_5 > thr) | ((pc_5[:, 2] <= 0) if model != "equirect" else False))   # local_bundle_adjuster_g2o.cc:323-344
    chi_a, _ = pb.edge_chi2(ref["pose_cw"], ref["points"], active)
    assert abs(chi_a.sum() - ref["chi2"][1]) <= 1e-9 * ref["chi2"][1]

    def fun_plain(x):
        T, P = pb.unpack(x)
        return np.sqrt(pb.edge_chi2(T, P, active)[0] + 1e-300)

    x2 = pb.pack(ref["pose_cw"], ref["points"])
    sol2 = least_squares(fun_plain, x2, method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-10, max_nfev=60)
    opt2 = 2.0 * sol2.cost
    assert opt2 <= ref["chi2"][1] * (1 + 1e-12)
    assert (ref["chi2"][1] - opt2) <= 5e-3 * opt2, (ref["chi2"][1], opt2)
