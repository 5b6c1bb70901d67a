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
