"""data::landmark::compute_descriptor: oracle vs a literal numpy walk (CPU), CUDA vs oracle (GPU, bit-exact index)."""
import numpy as np
import pytest

from oracle import pyoracle as O


def _literal(descs):
    n = len(descs)
    d = np.array([[int(np.unpackbits(descs[i] ^ descs[j]).sum()) for j in range(n)] for i in range(n)])
    best, best_idx = 256, 0
    for i in range(n):
        med = np.sort(d[i])[int(0.5 * (n - 1))]
        if med < best:
            best, best_idx = med, i
    return best_idx


def _landmarks(seed, n_landmarks, max_obs):
    rng = np.random.default_rng(seed)
    out = []
    for l in range(n_landmarks):
        n = int(rng.integers(1, max_obs + 1))
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        d = base ^ flips
        if l % 5 == 0 and n > 2:
            d[n // 2] = d[0]                      # exact duplicates: ties are decided by the first index
        out.append(d)
    return out


def test_oracle_matches_literal_walk():
    for d in _landmarks(0, 60, 12):
        assert O.landmark_descriptor(d) == _literal(d)
    one = _landmarks(1, 1, 1)[0][:1]
    assert O.landmark_descriptor(one) == 0


@pytest.mark.gpu
def test_gpu_matches_oracle():
    from stella_vslam_b200 import match
    lms = _landmarks(2, 3000, 9) + _landmarks(3, 40, 120) + [np.zeros((0, 32), np.uint8)] + _landmarks(4, 3, 512)
    best, rep = match.landmark_descriptors(lms)
    for l, d in enumerate(lms):
        if len(d) == 0:
            assert best[l] == -1 and not rep[l].any()
            continue
        want = O.landmark_descriptor(d)
        assert best[l] == want, l
        assert np.array_equal(rep[l], d[want])


@pytest.mark.gpu
def test_gpu_capacity_is_reported():
    from stella_vslam_b200 import match
    from stella_vslam_b200._lib import ERR_CAPACITY, B200Error
    with pytest.raises(B200Error) as e:
        match.landmark_descriptors([np.zeros((600, 32), np.uint8)])
    assert e.value.code == ERR_CAPACITY


# ---- data::landmark::update_mean_normal_and_obs_scale_variance (data/landmark.cc:256-311) ------------------------------------
def _geometry_case(seed, n):
    rng = np.random.default_rng(seed)
    pos = rng.normal(0, 20, (n, 3))
    cams = [pos[l] + rng.normal(0, 8, (int(rng.integers(1, 12)), 3)) for l in range(n)]
    cams[0][0] = pos[0]                                   # an observation from the landmark's own position: normalized() leaves the zero vector
    ref = np.stack([c[int(rng.integers(0, len(c)))] for c in cams])
    sf = (np.float32(1.2) ** rng.integers(0, 8, n)).astype(np.float32)
    return pos, cams, ref, sf, np.float32(1.0) / np.float32(1.2) ** 7


def test_geometry_oracle_matches_numpy():
    pos, cams, ref, sf, inv_last = _geometry_case(5, 300)
    mn, mx, mi = O.landmark_geometry(pos, cams, ref, sf, inv_last)
    for l in range(len(pos)):
        v = pos[l] - cams[l]
        nr = np.linalg.norm(v, axis=1, keepdims=True)
        m = np.where(nr > 0, v / np.where(nr > 0, nr, 1), v).sum(0)
        assert np.allclose(mn[l], m / np.linalg.norm(m), atol=1e-13)
        d = np.linalg.norm(pos[l] - ref[l])
        assert np.isclose(mx[l], np.float32(d * float(sf[l])), rtol=1e-6) and mi[l] == np.float32(mx[l] * inv_last)


@pytest.mark.gpu
def test_geometry_gpu_matches_oracle():
    from stella_vslam_b200 import match
    pos, cams, ref, sf, inv_last = _geometry_case(6, 5000)
    mn, mx, mi = match.landmark_geometry(pos, cams, ref, sf, inv_last)
    mn_w, mx_w, mi_w = O.landmark_geometry(pos, cams, ref, sf, inv_last)
    assert np.array_equal(mn, mn_w) and np.array_equal(mx, mx_w) and np.array_equal(mi, mi_w)   # same operations in the same order
