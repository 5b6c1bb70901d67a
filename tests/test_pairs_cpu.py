"""All-pairs matchers with greedy state (bow_tree::*, robust::match_for_triangulation): the C oracle (oracle/pairs_oracle.c)
against a literal Python walk of bow_tree.cc:11-366 / robust.cc:14-146 that keeps the reference's containers (an ordered
node -> indices map per keyframe and the merge-join over them).  No reference test exists for these functions: unpinned."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from stella_vslam_b200 import match
from workloads import synth


def _popcount(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def _angle_diff(a1, a2):
    r = np.float32(a1) - np.float32(a2)
    if r <= -180.0:
        r = np.float32(float(r) + 360.0)
    if r > 180.0:
        r = np.float32(float(r) - 360.0)
    return r


def _epipolar_ok(b1, b2, E, thr, scale):  # match/base.h:67-79
    e = [E[0] * b2[0] + E[1] * b2[1] + E[2] * b2[2], E[3] * b2[0] + E[4] * b2[1] + E[5] * b2[2], E[6] * b2[0] + E[7] * b2[1] + E[8] * b2[2]]
    dot = e[0] * b1[0] + e[1] * b1[1] + e[2] * b1[2]
    norm = math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
    c = min(1.0, max(-1.0, dot / norm))
    return abs(math.pi / 2.0 - math.acos(c)) < float(np.float32(thr) * np.float32(scale))


def literal_pairs(pr, variant, lowe, check_orientation):
    n1, n2 = len(pr["desc1"]), len(pr["desc2"])
    out = np.full(n1, -1, np.int32)
    taken = np.zeros(n2, bool)

    def feat_vec(nodes, n):  # bow_feature_vector: std::map<node, std::vector<idx>>
        fv = {}
        for i in range(n):
            fv.setdefault(int(nodes[i]) if nodes is not None else 0, []).append(i)
        return sorted(fv.items())

    fv1, fv2 = feat_vec(pr.get("node1"), n1), feat_vec(pr.get("node2"), n2)
    E = None if variant == 0 else [float(v) for v in np.asarray(pr["E_12"]).reshape(9)]
    i1 = i2 = 0
    while i1 < len(fv1) and i2 < len(fv2):
        if fv1[i1][0] < fv2[i2][0]:
            i1 += 1
            continue
        if fv2[i2][0] < fv1[i1][0]:
            i2 += 1
            continue
        for a in fv1[i1][1]:
            if pr.get("valid1") is not None and not pr["valid1"][a]:
                continue
            best, second, best_idx = (50 if variant == 1 else 256), 256, -1
            for b in fv2[i2][1]:
                if pr.get("valid2") is not None and not pr["valid2"][b]:
                    continue
                if taken[b]:
                    continue
                if check_orientation and abs(_angle_diff(pr["angle1"][a], pr["angle2"][b])) > 30.0:
                    continue
                d = _popcount(pr["desc1"][a], pr["desc2"][b])
                if variant == 1:
                    if 50 < d or best < d:
                        continue
                    st1 = pr.get("stereo1") is not None and pr["stereo1"][a]
                    st2 = pr.get("stereo2") is not None and pr["stereo2"][b]
                    b2 = [float(v) for v in pr["bearing2"][b]]
                    if pr["valid_epiplane"] and not st1 and not st2:
                        ep = [float(v) for v in pr["epiplane_in_keyfrm_2"]]
                        if 0.99862953475 < ep[0] * b2[0] + ep[1] * b2[1] + ep[2] * b2[2]:
                            continue
                    if not _epipolar_ok([float(v) for v in pr["bearing1"][a]], b2, E, pr["residual_rad_thr"], pr["scale1"][a]):
                        continue
                if d < best:
                    second, best, best_idx = best, d, b
                elif d < second:
                    second = d
            if variant == 1:
                if best_idx < 0:
                    continue
            elif 50 < best:
                continue
            if np.float32(lowe) * np.float32(second) < np.float32(best):
                continue
            taken[best_idx] = True
            out[a] = best_idx
        i1 += 1
        i2 += 1
    return out


def _bow_problem(k1, k2, both_sides):
    pr = dict(desc1=k1["desc"], angle1=k1["angle"], valid1=k1["has_landmark"], node1=k1["node"], desc2=k2["desc"], angle2=k2["angle"],
              node2=k2["node"])
    if both_sides:
        pr["valid2"] = k2["has_landmark"]
    return pr


@pytest.mark.parametrize("with_nodes", [False, True])
@pytest.mark.parametrize("stereo", [False, True])
def test_triangulation_oracle_matches_literal_walk(with_nodes, stereo):
    k1, k2, g = synth.make_keyframe_pair(3, n1=500, n2=450, stereo=stereo, n_nodes=12)
    pr = match._triangulation_problem(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"], True, 0.2 * np.pi / 180.0, with_nodes)
    got, n = O.match_pairs(pr, 1, 0.75, True)
    want = literal_pairs(pr, 1, 0.75, True)
    assert np.array_equal(got, want) and n == (want >= 0).sum() > 15


@pytest.mark.parametrize("both_sides", [False, True])
@pytest.mark.parametrize("lowe", [0.6, 0.9])
def test_bow_oracle_matches_literal_walk(both_sides, lowe):
    k1, k2, _ = synth.make_keyframe_pair(4, n1=500, n2=520, n_nodes=10)
    pr = _bow_problem(k1, k2, both_sides)
    got, n = O.match_pairs(pr, 0, lowe, True)
    want = literal_pairs(pr, 0, lowe, True)
    assert np.array_equal(got, want) and n == (want >= 0).sum() > 30
    hit = got[got >= 0]
    assert len(np.unique(hit)) == len(hit)


def test_epipolar_constraint_kat():
    """bearing_2 on the epipolar plane passes; tilted out of it (residual ~2 degrees) it fails at 0.2 degrees x scale 1 and x 8, passes x 16."""
    E = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 0]], np.float64)  # t = z: x1 . (z x x2) = 0 for coplanar bearings
    thr = np.float32(0.2 * np.pi / 180.0)
    on = np.array([0.3, 0.0, 0.954])
    off = np.array([0.3, np.sin(np.deg2rad(1.0)), 0.954])
    b1 = np.array([0.6, 0.0, 0.8])
    f = O.lib().orc_check_epipolar_constraint
    import ctypes as C
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
    call = lambda b2, s: f(b1.ctypes.data, np.ascontiguousarray(b2).ctypes.data, E.ctypes.data, thr, s)
    assert call(on, 1.0) == 1 and call(off, 1.0) == 0 and call(off, 8.0) == 0 and call(off, 16.0) == 1
