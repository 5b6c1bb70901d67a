"""GPU parity tests for the Hamming matchers (match/base.h, match/robust.cc:232-328) against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from oracle import pyoracle as O
    from stella_vslam_b200 import match, synth
    return O, match, synth


def test_hamming_kat(mods, golden_dir):
    # test/stella_vslam/match/base.cc:11-57
    O, match, synth = mods
    g = np.load(os.path.join(golden_dir, "hamming_kat.npz"))
    for a, b, d in zip(g["a"], g["b"], g["dist"]):
        assert match.compute_descriptor_distance_32(a, b) == d
        assert match.compute_descriptor_distance_64(a, b) == d


def test_hamming_matrix_vs_oracle(mods):
    O, match, synth = mods
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (70, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (133, 32), dtype=np.uint8)
    m = match.hamming_matrix(a, b)
    ref = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2)
    assert np.array_equal(m, ref)
    assert m[3, 7] == O.hamming_32(a[3], b[7])


@pytest.mark.parametrize("n1,n2,lowe,ori,seed", [(2000, 2000, 0.8, True, 7), (2000, 2000, 0.95, False, 8), (500, 1300, 0.7, True, 9),
                                                 (1, 1, 0.8, True, 10), (37, 5, 0.6, False, 11), (3000, 2500, 0.75, True, 12)])
def test_brute_force_vs_oracle(mods, n1, n2, lowe, ori, seed):
    O, match, synth = mods
    d1, a1, d2, a2, v2 = synth.make_descriptor_pair(n1, n2, seed=seed)
    m = match.robust(lowe, ori)
    got = m.brute_force_match(d1, a1, d2, a2, v2)
    ref = O.brute_force_match(d1, a1, d2, a2, v2, lowe, ori)
    assert np.array_equal(got, ref)
    if n1 >= 500:
        assert len(ref) > 50


def test_brute_force_collisions_force_exact_fallback(mods):
    # many keyframe keypoints compete for few frame keypoints: candidate lists get exhausted by "taken" entries
    O, match, synth = mods
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    d1 = np.repeat(base, 3, axis=0)                       # 36 frame descriptors, triplicates
    d1[1::3, 0] ^= 1
    d1[2::3, 1] ^= 3
    d2 = np.repeat(base, 40, axis=0)                      # 480 keyframe descriptors hitting the same 36
    flips = rng.integers(0, 256, (480,))
    for i, b in enumerate(flips):
        d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    a1 = np.zeros(36, np.float32)
    a2 = np.zeros(480, np.float32)
    for lowe in (0.6, 0.8, 1.0):
        got = match.robust(lowe, False).brute_force_match(d1, a1, d2, a2)
        ref = O.brute_force_match(d1, a1, d2, a2, None, lowe, False)
        assert np.array_equal(got, ref)


def test_identical_descriptors_tie_break(mods):
    # all distances tie: the reference's strict '<' keeps the first frame keypoint in index order
    O, match, synth = mods
    d1 = np.zeros((50, 32), np.uint8)
    d2 = np.zeros((20, 32), np.uint8)
    a = np.zeros(50, np.float32)
    got = match.robust(1.0, True).brute_force_match(d1, a, d2, a[:20])
    ref = O.brute_force_match(d1, a, d2, a[:20], None, 1.0, True)
    assert np.array_equal(got, ref)


def test_batch_and_empty_problems(mods):
    O, match, synth = mods
    probs = []
    for s, (n1, n2) in enumerate([(300, 400), (0, 10), (10, 0), (1200, 900)]):
        d1, a1, d2, a2, v2 = synth.make_descriptor_pair(max(n1, 1), max(n2, 1), seed=40 + s)
        probs.append((d1[:n1], a1[:n1], d2[:n2], a2[:n2], v2[:n2]))
    m = match.robust(0.8, True)
    got = m.brute_force_match_batch(probs)
    for g, (d1, a1, d2, a2, v2) in zip(got, probs):
        assert np.array_equal(g, O.brute_force_match(d1, a1, d2, a2, v2, 0.8, True))


def test_orientation_gate_boundaries(mods):
    O, match, synth = mods
    d = np.zeros((4, 32), np.uint8)
    a1 = np.array([0.0, 30.0, 30.000002, 359.0], np.float32)
    for q in (0.0, 329.0, 330.0, 180.0, 59.999996):
        a2 = np.array([q], np.float32)
        got = match.robust(1.0, True).brute_force_match(d, a1, d[:1], a2)
        ref = O.brute_force_match(d, a1, d[:1], a2, None, 1.0, True)
        assert np.array_equal(got, ref)
