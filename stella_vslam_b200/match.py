"""Host-side mirror of the reference's match:: surface for the Hamming matchers
(src/stella_vslam/match/base.h:15-91, match/robust.h, match/robust.cc:232-328)."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

HAMMING_DIST_THR_LOW = 50    # match/base.h:15
HAMMING_DIST_THR_HIGH = 100  # match/base.h:16
MAX_HAMMING_DIST = 256       # match/base.h:17

_matchers = {}


def _matcher(device=0):
    if device not in _matchers:
        h = C.c_void_p()
        check(lib().b200_matcher_create(device, C.byref(h)))
        _matchers[device] = h
    return _matchers[device]


def hamming_matrix(desc_1, desc_2, device=0):
    """All-pairs compute_descriptor_distance_32 (match/base.h:20-41): (n1, n2) uint16."""
    d1, d2 = np.ascontiguousarray(desc_1, np.uint8).reshape(-1, 32), np.ascontiguousarray(desc_2, np.uint8).reshape(-1, 32)
    out = np.zeros((d1.shape[0], d2.shape[0]), np.uint16)
    check(lib().b200_hamming_matrix(_matcher(device), ptr(d1), d1.shape[0], ptr(d2), d2.shape[0], ptr(out)))
    return out


def compute_descriptor_distance_32(desc_1, desc_2, device=0):
    return int(hamming_matrix(np.asarray(desc_1).reshape(1, 32), np.asarray(desc_2).reshape(1, 32), device)[0, 0])


compute_descriptor_distance_64 = compute_descriptor_distance_32  # same value (match/base.h:44-65)


class base:
    """match::base (match/base.h:81-91)."""

    def __init__(self, lowe_ratio, check_orientation):
        self.lowe_ratio_ = np.float32(lowe_ratio)
        self.check_orientation_ = bool(check_orientation)


class robust(base):
    """match::robust: the brute-force matcher (match/robust.cc:232-328)."""

    def __init__(self, lowe_ratio, check_orientation, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def brute_force_match(self, frm_desc, frm_angles, keyfrm_desc, keyfrm_angles, keyfrm_has_landmark=None):
        """One (frame, keyframe) pair.  Returns an (n, 2) int32 array of (idx_1, idx_2) sorted by idx_1."""
        return self.brute_force_match_batch([(frm_desc, frm_angles, keyfrm_desc, keyfrm_angles, keyfrm_has_landmark)])[0]

    def brute_force_match_batch(self, problems):
        """problems: list of (frm_desc (n1,32) u8, frm_angles (n1,), keyfrm_desc (n2,32), keyfrm_angles (n2,), valid2|None)."""
        P = len(problems)
        if P == 0:
            return []
        cnt1 = np.array([len(pr[0]) for pr in problems], np.int32)
        cnt2 = np.array([len(pr[2]) for pr in problems], np.int32)
        off1 = np.concatenate([[0], np.cumsum(cnt1)[:-1]]).astype(np.int32)
        off2 = np.concatenate([[0], np.cumsum(cnt2)[:-1]]).astype(np.int32)

        def cat(k, dt, shape):
            return np.ascontiguousarray(np.concatenate([np.asarray(pr[k], dt).reshape(shape) for pr in problems]))

        d1, a1 = cat(0, np.uint8, (-1, 32)), cat(1, np.float32, (-1,))
        d2, a2 = cat(2, np.uint8, (-1, 32)), cat(3, np.float32, (-1,))
        v2 = None
        if any(pr[4] is not None for pr in problems):
            v2 = np.ascontiguousarray(np.concatenate([
                (np.asarray(pr[4], np.uint8).reshape(-1) if pr[4] is not None else np.ones(len(pr[2]), np.uint8)) for pr in problems]))
        stride = max(int(cnt1.max()), 1)
        pairs = np.zeros((P, stride, 2), np.int32)
        n_pairs = np.zeros(P, np.int32)
        check(lib().b200_match_bruteforce(_matcher(self.device), P, ptr(d1), ptr(a1), 4, ptr(off1), ptr(cnt1), ptr(d2), ptr(a2), 4,
                                          ptr(v2), ptr(off2), ptr(cnt2), float(self.lowe_ratio_), int(self.check_orientation_),
                                          ptr(pairs), stride, ptr(n_pairs)))
        return [pairs[i, :n_pairs[i]].copy() for i in range(P)]
