"""Host-side mirror of the reference's match:: surface for the Hamming matchers
(src/stella_vslam/match/base.h:15-91, match/robust.h, match/robust.cc:232-328)."""
import ctypes as C
import threading

import numpy as np

from ._lib import KP_DTYPE, GuidedProblem, PairsProblem, check, lib, pack_guided_problem, pack_pairs_problem, ptr

HAMMING_DIST_THR_LOW = 50    # match/base.h:15
HAMMING_DIST_THR_HIGH = 100  # match/base.h:16
MAX_HAMMING_DIST = 256       # match/base.h:17

_tls = threading.local()


def _matcher(device=0):
    """One matcher handle (stream + staging buffers) per (thread, device): a handle must not be shared by concurrent callers, exactly
    like the reference adapters keep theirs thread_local."""
    cache = getattr(_tls, "matchers", None)
    if cache is None:
        cache = _tls.matchers = {}
    if device not in cache:
        h = C.c_void_p()
        check(lib().b200_matcher_create(device, C.byref(h)))
        cache[device] = h
    return cache[device]


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


GUIDED_LANDMARKS, GUIDED_LAST_FRAME, GUIDED_INDEPENDENT, GUIDED_FUSE, GUIDED_AREA = range(5)  # b200vslam.h


def match_guided_batch(problems, mode, thr, lowe_ratio, check_orientation, max_candidates=0, device=0):
    """b200_match_guided on flattened problem dicts (see _lib.pack_guided_problem).
    Returns [(match_out, occupied_after | None, n_matches)] per problem."""
    if not problems:
        return []
    arr = (GuidedProblem * len(problems))()
    keeps = []
    for i, pr in enumerate(problems):
        S, keep = pack_guided_problem(pr)
        arr[i] = S
        keeps.append(keep)
    check(lib().b200_match_guided(_matcher(device), len(problems), arr, mode, int(thr), float(lowe_ratio), int(check_orientation), max_candidates))
    return [(k["match_out"][:arr[i].n_queries].copy(), k.get("t_occupied"), int(arr[i].n_matches)) for i, k in enumerate(keeps)]


def _guided_problem(frm, desc, reproj, level, min_level, max_level, margin, x_right=None, angle=None, valid=None, occupied=None, **extra):
    """A frame dict (t_x, t_y, t_octave, t_angle, t_desc, [t_x_right], bounds, grid, scale_factors) + per-landmark arrays."""
    sf = np.asarray(frm["scale_factors"], np.float32)
    level = np.asarray(level, np.int64)
    reproj = np.asarray(reproj, np.float64).reshape(-1, 2)
    pr = {k: v for k, v in frm.items() if k.startswith("t_") or k in ("bounds", "grid")}
    if occupied is not None:
        pr["t_occupied"] = occupied
    pr.update(q_desc=desc, q_x=reproj[:, 0].astype(np.float32), q_y=reproj[:, 1].astype(np.float32), q_margin=np.float32(margin) * sf[level],
              q_min_level=min_level, q_max_level=max_level, q_x_right=x_right, q_angle=angle, q_valid=valid, **extra)
    return pr


def _level_window(frm, level):
    lv = np.asarray(level, np.int64)
    return lv, np.maximum(0, lv - 1), np.minimum(len(frm["scale_factors"]) - 1, lv + 1)


class projection(base):
    """match::projection (match/projection.h:24-67).  A frame / keyframe is a dict with t_x, t_y, t_octave, t_angle, t_desc,
    optional t_x_right, bounds=(min_x, max_x, min_y, max_y) = camera img_bounds_, grid=(64, 48), scale_factors.  The map-side
    work of each method (walking landmarks, reprojecting, predicting the level, validity) is the caller's, as in the adapter."""

    GUIDED_LANDMARKS, GUIDED_LAST_FRAME = GUIDED_LANDMARKS, GUIDED_LAST_FRAME

    def __init__(self, lowe_ratio=0.6, check_orientation=True, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def match_guided_batch(self, problems, mode, thr=HAMMING_DIST_THR_HIGH, max_candidates=0):
        return match_guided_batch(problems, mode, thr, self.lowe_ratio_, self.check_orientation_, max_candidates, self.device)

    def match_frame_and_landmarks(self, frm, lm_desc, lm_to_reproj, lm_to_scale, margin=5.0, lm_to_x_right=None, valid=None, occupied=None):
        """projection.cc:13-93.  valid[q] = 0 for landmarks without a reprojection or about to be erased (:23-28)."""
        lv, lo, hi = _level_window(frm, lm_to_scale)
        pr = _guided_problem(frm, lm_desc, lm_to_reproj, lv, lo, hi, margin, lm_to_x_right, None, valid, occupied)
        return self.match_guided_batch([pr], GUIDED_LANDMARKS)[0]

    def match_current_and_last_frames(self, curr_frm, last_desc, last_reproj, last_octave, last_angle, margin, last_x_right=None, valid=None,
                                      occupied=None, assume_forward=False, assume_backward=False):
        """projection.cc:95-207; the level window follows :140-154."""
        lv, lo, hi = _level_window(curr_frm, last_octave)
        if assume_forward:
            lo = lv
        elif assume_backward:
            hi = lv
        pr = _guided_problem(curr_frm, last_desc, last_reproj, lv, lo, hi, margin, last_x_right, last_angle, valid, occupied)
        return self.match_guided_batch([pr], GUIDED_LAST_FRAME)[0]

    def match_frame_and_keyframe(self, frm, lm_desc, reproj, pred_scale_level, keyfrm_angle, margin, hamm_dist_thr, valid=None, occupied=None):
        """projection.cc:217-319: the keyframe's landmarks into a frame whose pose is a candidate (relocalisation); no stereo gate."""
        lv, lo, hi = _level_window(frm, pred_scale_level)
        mono = {k: v for k, v in frm.items() if k != "t_x_right"}
        pr = _guided_problem(mono, lm_desc, reproj, lv, lo, hi, margin, None, keyfrm_angle, valid, occupied)
        return self.match_guided_batch([pr], GUIDED_LAST_FRAME, thr=hamm_dist_thr)[0]

    def match_by_Sim3_transform(self, keyfrm, lm_desc, reproj, pred_scale_level, margin, valid=None, occupied=None):
        """projection.cc:321-416 (loop closure): thr = HAMMING_DIST_THR_LOW, no orientation, no stereo gate."""
        lv, lo, hi = _level_window(keyfrm, pred_scale_level)
        mono = {k: v for k, v in keyfrm.items() if k != "t_x_right"}
        pr = _guided_problem(mono, lm_desc, reproj, lv, lo, hi, margin, None, None, valid, occupied)
        return match_guided_batch([pr], GUIDED_LAST_FRAME, HAMMING_DIST_THR_LOW, self.lowe_ratio_, False, 0, self.device)[0]

    def match_keyframes_mutually(self, keyfrm_1, keyfrm_2, lms_1, lms_2, margin):
        """projection.cc:418-630.  lms_k = dict(desc, reproj (into the OTHER keyframe), level, valid) for the landmarks of keyframe k
        (valid excludes null / erased / already matched / out-of-range ones, :461-496).  Returns (idx_2 per keypoint of 1 | -1, n)."""
        prs = []
        for target, lms in ((keyfrm_2, lms_1), (keyfrm_1, lms_2)):
            lv, lo, hi = _level_window(target, lms["level"])
            mono = {k: v for k, v in target.items() if k not in ("t_x_right", "t_occupied")}
            prs.append(_guided_problem(mono, lms["desc"], lms["reproj"], lv, lo, hi, margin, None, None, lms.get("valid")))
        (m12, _, _), (m21, _, _) = match_guided_batch(prs, GUIDED_INDEPENDENT, HAMMING_DIST_THR_HIGH, self.lowe_ratio_, False, 0, self.device)
        return cross_check(m12, m21)


def cross_check(idx2_in_1, idx1_in_2):
    a, b = np.ascontiguousarray(idx2_in_1, np.int32), np.ascontiguousarray(idx1_in_2, np.int32)
    out = np.full(max(len(a), 1), -2, np.int32)
    n = C.c_int32(0)
    check(lib().b200_match_cross_check(ptr(a), len(a), ptr(b), len(b), ptr(out), C.byref(n)))
    return out[:len(a)].copy(), n.value


class fuse(base):
    """match::fuse (match/fuse.h, fuse.cc:12-154)."""

    def __init__(self, lowe_ratio=0.6, check_orientation=True, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def detect_duplication(self, keyfrm, lm_desc, reproj, pred_scale_level, margin, x_right=None, valid=None, do_reprojection_matching=False,
                           inv_level_sigma_sq=None):
        """Returns (best_idx per landmark | -1, n_fused); the caller splits the hits into duplicated_lms_in_keyfrm / new_connections
        by whether keypoint best_idx already carries a landmark (fuse.cc:131-146)."""
        lv, lo, hi = _level_window(keyfrm, pred_scale_level)
        pr = _guided_problem(keyfrm, lm_desc, reproj, lv, lo, hi, margin, x_right, None, valid, None,
                             q_reproj=np.asarray(reproj, np.float64).reshape(-1, 2), inv_level_sigma_sq=inv_level_sigma_sq,
                             do_reprojection_matching=do_reprojection_matching)
        got, _, n = match_guided_batch([pr], GUIDED_FUSE, HAMMING_DIST_THR_LOW, self.lowe_ratio_, False, 0, self.device)[0]
        return got, n


class area(base):
    """match::area (match/area.h, area.cc:8-98): the initialiser's matcher."""

    def __init__(self, lowe_ratio=0.9, check_orientation=True, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def match_in_consistent_area(self, frm_1, frm_2, prev_matched_pts, margin):
        """frm_k: frame dicts.  Returns (matched_indices_2_in_frm_1, n, updated prev_matched_pts) (area.cc:89-95)."""
        oct1 = np.asarray(frm_1["t_octave"], np.int64)
        n1 = len(oct1)
        prev = np.array(prev_matched_pts, np.float32).reshape(-1, 2)
        pr = {k: v for k, v in frm_2.items() if (k.startswith("t_") and k not in ("t_x_right", "t_occupied")) or k in ("bounds", "grid")}
        pr.update(q_desc=frm_1["t_desc"], q_x=prev[:, 0], q_y=prev[:, 1], q_margin=np.full(n1, np.float32(int(margin))),
                  q_min_level=np.zeros(n1, np.int8), q_max_level=np.zeros(n1, np.int8), q_angle=frm_1["t_angle"],
                  q_valid=(oct1 <= 0).astype(np.uint8))
        got, _, n = match_guided_batch([pr], GUIDED_AREA, HAMMING_DIST_THR_LOW, self.lowe_ratio_, self.check_orientation_, 0, self.device)[0]
        hit = got >= 0
        prev[hit, 0] = np.asarray(frm_2["t_x"], np.float32)[got[hit]]
        prev[hit, 1] = np.asarray(frm_2["t_y"], np.float32)[got[hit]]
        return got, n, prev


PAIRS_BOW, PAIRS_TRIANGULATION = 0, 1  # b200vslam.h


def match_pairs_batch(problems, variant, lowe_ratio, check_orientation, max_candidates=0, device=0):
    """b200_match_pairs on problem dicts (see _lib.pack_pairs_problem).  Returns [(match_out per row, n_matches)]."""
    if not problems:
        return []
    arr = (PairsProblem * len(problems))()
    keeps = []
    for i, pr in enumerate(problems):
        S, keep = pack_pairs_problem(pr)
        arr[i] = S
        keeps.append(keep)
    check(lib().b200_match_pairs(_matcher(device), len(problems), arr, variant, float(lowe_ratio), int(check_orientation), max_candidates))
    return [(k["match_out"][:arr[i].n1].copy(), int(arr[i].n_matches)) for i, k in enumerate(keeps)]


def _triangulation_problem(keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane, residual_rad_thr, with_nodes):
    sf = np.asarray(keyfrm_1["scale_factors"], np.float32)
    pr = dict(desc1=keyfrm_1["desc"], angle1=keyfrm_1["angle"], valid1=keyfrm_1.get("no_landmark"), bearing1=keyfrm_1["bearings"],
              scale1=sf[np.asarray(keyfrm_1["octave"], np.int64)], stereo1=keyfrm_1.get("stereo"),
              desc2=keyfrm_2["desc"], angle2=keyfrm_2["angle"], valid2=keyfrm_2.get("no_landmark"), bearing2=keyfrm_2["bearings"],
              stereo2=keyfrm_2.get("stereo"), E_12=E_12, epiplane_in_keyfrm_2=epiplane_in_keyfrm_2, valid_epiplane=valid_epiplane,
              residual_rad_thr=residual_rad_thr)
    if with_nodes:
        pr.update(node1=keyfrm_1["node"], node2=keyfrm_2["node"])
    return pr


def _pairs_of(match_out):
    idx_1 = np.flatnonzero(match_out >= 0)
    return np.stack([idx_1, match_out[idx_1]], 1).astype(np.int32)  # matched_idx_pairs, sorted by idx_1 (robust.cc:137-143)


def _match_for_triangulation(self, keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane=True, residual_rad_thr=0.2 * np.pi / 180.0,
                             with_nodes=False):
    pr = _triangulation_problem(keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane, residual_rad_thr, with_nodes)
    got, n = match_pairs_batch([pr], PAIRS_TRIANGULATION, self.lowe_ratio_, self.check_orientation_, 0, self.device)[0]
    return _pairs_of(got)


def _robust_match_for_triangulation(self, keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane=True, residual_rad_thr=0.2 * np.pi / 180.0):
    """robust::match_for_triangulation (robust.cc:14-146).  keyfrm_k: dict(desc, angle, octave, bearings (n,3) f64, scale_factors,
    no_landmark (u8: keypoint carries no landmark), stereo (u8) | None).  Returns matched_idx_pairs (n, 2)."""
    return _match_for_triangulation(self, keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane, residual_rad_thr, False)


robust.match_for_triangulation = _robust_match_for_triangulation


class bow_tree(base):
    """match::bow_tree (match/bow_tree.h, bow_tree.cc).  Keyframes / frames are dicts as for robust.match_for_triangulation plus
    node = the BoW node id of every keypoint (the bow_feat_vec_ entry that lists it)."""

    def __init__(self, lowe_ratio=0.6, check_orientation=True, device=0):
        super().__init__(lowe_ratio, check_orientation)
        self.device = device

    def match_for_triangulation(self, keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane=True, residual_rad_thr=0.2 * np.pi / 180.0):
        """bow_tree.cc:11-167."""
        return _match_for_triangulation(self, keyfrm_1, keyfrm_2, E_12, epiplane_in_keyfrm_2, valid_epiplane, residual_rad_thr, True)

    def match_frame_and_keyframe(self, keyfrm, frm):
        """bow_tree.cc:169-256.  keyfrm["has_landmark"]: live landmark per keypoint.  Returns (frame index per keyframe keypoint | -1, n):
        matched_lms_in_frm[out[i]] = keyfrm landmark i."""
        pr = dict(desc1=keyfrm["desc"], angle1=keyfrm["angle"], valid1=keyfrm["has_landmark"], node1=keyfrm["node"],
                  desc2=frm["desc"], angle2=frm["angle"], node2=frm["node"])
        return match_pairs_batch([pr], PAIRS_BOW, self.lowe_ratio_, self.check_orientation_, 0, self.device)[0]

    def match_keyframes(self, keyfrm_1, keyfrm_2):
        """bow_tree.cc:258-366: both sides restricted to keypoints with live landmarks.  Returns (idx_2 per keypoint of 1 | -1, n)."""
        pr = dict(desc1=keyfrm_1["desc"], angle1=keyfrm_1["angle"], valid1=keyfrm_1["has_landmark"], node1=keyfrm_1["node"],
                  desc2=keyfrm_2["desc"], angle2=keyfrm_2["angle"], valid2=keyfrm_2["has_landmark"], node2=keyfrm_2["node"])
        return match_pairs_batch([pr], PAIRS_BOW, self.lowe_ratio_, self.check_orientation_, 0, self.device)[0]


class stereo:
    """match::stereo (match/stereo.h:17-101, stereo.cc).  Constructed like the reference's (system.cc:443) from the two extractors
    whose last extract() produced the rectified pair -- their image pyramids stay on the device -- plus the keypoints and
    descriptors of both eyes.  frame_left / frame_right select a frame when both eyes went through one batched extract."""

    hamm_dist_thr_ = (HAMMING_DIST_THR_HIGH + HAMMING_DIST_THR_LOW) // 2  # stereo.h:99

    def __init__(self, extractor_left, extractor_right, keypts_left, keypts_right, descs_left, descs_right, focal_x_baseline, true_baseline,
                 frame_left=0, frame_right=0, device=0):
        self.extractor_left, self.extractor_right = extractor_left, extractor_right
        self.frame_left, self.frame_right = frame_left, frame_right
        self.keypts_left_ = np.ascontiguousarray(keypts_left, KP_DTYPE)
        self.keypts_right_ = np.ascontiguousarray(keypts_right, KP_DTYPE)
        self.descs_left_ = np.ascontiguousarray(descs_left, np.uint8).reshape(-1, 32)
        self.descs_right_ = np.ascontiguousarray(descs_right, np.uint8).reshape(-1, 32)
        self.focal_x_baseline_, self.true_baseline_ = float(focal_x_baseline), float(true_baseline)
        self.device = device

    def compute(self):
        """stereo::compute (stereo.cc:20-114).  Returns (stereo_x_right, depths), -1 where no stereo match survives."""
        n = len(self.keypts_left_)
        x_right, depths = np.full(max(n, 1), -1, np.float32), np.full(max(n, 1), -1, np.float32)
        kept = C.c_int32(0)
        check(lib().b200_stereo_compute(_matcher(self.device), self.extractor_left._h, self.frame_left, self.extractor_right._h, self.frame_right,
                                        ptr(self.keypts_left_), ptr(self.descs_left_), n, ptr(self.keypts_right_), ptr(self.descs_right_),
                                        len(self.keypts_right_), self.focal_x_baseline_, self.true_baseline_, ptr(x_right), ptr(depths),
                                        C.byref(kept)))
        self.num_matched_ = kept.value
        return x_right[:n], depths[:n]


def landmark_descriptors(desc_lists, device=0):
    """data::landmark::compute_descriptor (data/landmark.cc:199-256) for many landmarks: desc_lists[l] = (n_l, 32) uint8 descriptors of
    the landmark's observations.  Returns (best_idx (L,), representative descriptors (L, 32))."""
    n = len(desc_lists)
    cnt = np.array([len(d) for d in desc_lists], np.int32)
    offsets = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_lists]) if n and cnt.sum() else np.zeros((1, 32), np.uint8))
    best = np.full(max(n, 1), -9, np.int32)
    out = np.zeros((max(n, 1), 32), np.uint8)
    check(lib().b200_landmark_descriptors(_matcher(device), n, ptr(flat), ptr(offsets), ptr(best), ptr(out)))
    return best[:n], out[:n]


def landmark_geometry(pos_w, cam_center_lists, ref_center, ref_scale_factor, inv_scale_factor_last, device=0):
    """data::landmark::update_mean_normal_and_obs_scale_variance (data/landmark.cc:256-311) for many landmarks.
    cam_center_lists[l] = (n_l, 3) camera centres of the landmark's observations.  Returns (mean_normal (L,3), max_valid, min_valid)."""
    n = len(cam_center_lists)
    cnt = np.array([len(c) for c in cam_center_lists], np.int32)
    offsets = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(c, np.float64).reshape(-1, 3) for c in cam_center_lists]) if n and cnt.sum() else np.zeros((1, 3)))
    pos, ref = np.ascontiguousarray(pos_w, np.float64).reshape(-1, 3), np.ascontiguousarray(ref_center, np.float64).reshape(-1, 3)
    sf = np.ascontiguousarray(ref_scale_factor, np.float32)
    mn, mx, mi = np.zeros((max(n, 1), 3)), np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
    check(lib().b200_landmark_geometry(_matcher(device), n, ptr(pos), ptr(offsets), ptr(flat), ptr(ref), ptr(sf), float(inv_scale_factor_last),
                                       ptr(mn), ptr(mx), ptr(mi)))
    return mn[:n], mx[:n], mi[:n]
