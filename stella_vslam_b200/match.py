"""Host-side mirror of the reference's match:: surface for the Hamming matchers
(src/stella_vslam/match/base.h:15-91, match/robust.h, match/robust.cc:232-328)."""
import ctypes as C

import numpy as np

from ._lib import GuidedProblem, check, lib, pack_guided_problem, ptr

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


class projection(base):
    """match::projection (match/projection.h:24-67): the two per-frame grid-guided matchers
    match_frame_and_landmarks (projection.cc:13-93) and match_current_and_last_frames (:95-207).

    A frame is described by a dict `frm` with t_x, t_y, t_octave, t_angle, t_desc, optional t_x_right / t_occupied,
    bounds=(min_x, max_x, min_y, max_y) = camera img_bounds_, grid=(64, 48), scale_factors (orb_params_->scale_factors_)."""

    GUIDED_LANDMARKS, GUIDED_LAST_FRAME = 0, 1

    def __init__(self, lowe_ratio=0.6, check_orientation=True, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def match_guided_batch(self, problems, mode, thr=HAMMING_DIST_THR_HIGH, max_candidates=0):
        """problems: flattened dicts (see _lib.pack_guided_problem).  Returns [(match_out, occupied_after | None, n_matches)]."""
        if not problems:
            return []
        arr = (GuidedProblem * len(problems))()
        keeps = []
        for i, pr in enumerate(problems):
            S, keep = pack_guided_problem(pr)
            arr[i] = S
            keeps.append(keep)
        check(lib().b200_match_guided(_matcher(self.device), len(problems), arr, mode, thr, float(self.lowe_ratio_), int(self.check_orientation_),
                                      max_candidates))
        return [(k["match_out"][:arr[i].n_queries].copy(), k.get("t_occupied"), int(arr[i].n_matches)) for i, k in enumerate(keeps)]

    @staticmethod
    def _query_fields(frm, desc, reproj, level, min_level, max_level, margin, x_right, angle, valid):
        sf = np.asarray(frm["scale_factors"], np.float32)
        level = np.asarray(level, np.int64)
        reproj = np.asarray(reproj, np.float64).reshape(-1, 2)
        pr = {k: v for k, v in frm.items() if k.startswith("t_") or k in ("bounds", "grid")}
        pr.update(q_desc=desc, q_x=reproj[:, 0].astype(np.float32), q_y=reproj[:, 1].astype(np.float32),
                  q_margin=np.float32(margin) * sf[level], q_min_level=min_level, q_max_level=max_level, q_x_right=x_right, q_angle=angle,
                  q_valid=valid)
        return pr

    def match_frame_and_landmarks(self, frm, lm_desc, lm_to_reproj, lm_to_scale, margin=5.0, lm_to_x_right=None, valid=None):
        """Landmark q (in local_landmarks order) reprojects to lm_to_reproj[q] at predicted level lm_to_scale[q];
        valid[q] = 0 for landmarks without a reprojection or about to be erased (projection.cc:23-28)."""
        n_levels = len(frm["scale_factors"])
        lv = np.asarray(lm_to_scale, np.int64)
        pr = self._query_fields(frm, lm_desc, lm_to_reproj, lv, np.maximum(0, lv - 1), np.minimum(n_levels - 1, lv + 1), margin, lm_to_x_right,
                                None, valid)
        return self.match_guided_batch([pr], self.GUIDED_LANDMARKS)[0]

    def match_current_and_last_frames(self, curr_frm, last_desc, last_reproj, last_octave, last_angle, margin, last_x_right=None, valid=None,
                                      assume_forward=False, assume_backward=False):
        """Query q = keypoint q of the last frame that carries a landmark (valid[q]), reprojected into the current frame
        (projection.cc:121-160); level window per :140-154."""
        n_levels = len(curr_frm["scale_factors"])
        lv = np.asarray(last_octave, np.int64)
        lo, hi = np.maximum(0, lv - 1), np.minimum(n_levels - 1, lv + 1)
        if assume_forward:
            lo = lv
        elif assume_backward:
            hi = lv
        pr = self._query_fields(curr_frm, last_desc, last_reproj, lv, lo, hi, margin, last_x_right, last_angle, valid)
        return self.match_guided_batch([pr], self.GUIDED_LAST_FRAME)[0]
