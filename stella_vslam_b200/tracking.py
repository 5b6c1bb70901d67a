"""Host-side mirror of the device-resident tracking chain (include/b200vslam.h: b200_track_local_map):
tracking_module::search_local_landmarks (tracking_module.cc:533-606) + pose_optimizer::optimize (pose_optimizer_g2o.cc:38-175) for a
batch of frames whose keypoints / descriptors are the results of the last extract of an orb_extractor and never leave the GPU."""
import ctypes as C

import numpy as np

from . import match, optimize
from ._lib import CameraIntrinsics, check, lib, ptr


class TrackParams(C.Structure):
    """b200_track_params_t"""
    _fields_ = [("cam", CameraIntrinsics), ("focal_x_baseline", C.c_double), ("monocular", C.c_int32), ("img_bounds", C.c_float * 4),
                ("grid_cols", C.c_int32), ("grid_rows", C.c_int32), ("num_levels", C.c_uint32), ("log_scale_factor", C.c_float),
                ("scale_factors", C.c_void_p), ("inv_level_sigma_sq", C.c_void_p), ("margin", C.c_float), ("lowe_ratio", C.c_float),
                ("hamming_thr", C.c_uint32), ("ray_cos_thr", C.c_float), ("num_trials_robust", C.c_int32), ("num_trials", C.c_int32),
                ("num_each_iter", C.c_int32), ("max_candidates", C.c_int32)]


class TrackFrame(C.Structure):
    """b200_track_frame_t"""
    _fields_ = [("frame", C.c_int32), ("pose_cw", C.c_void_p), ("n_keypoints_in", C.c_int32), ("kp_x_right", C.c_void_p),
                ("kp_landmark", C.c_void_p), ("n_landmarks", C.c_int32), ("lm_pos_w", C.c_void_p), ("lm_mean_normal", C.c_void_p),
                ("lm_min_valid_dist", C.c_void_p), ("lm_max_valid_dist", C.c_void_p), ("lm_desc", C.c_void_p), ("lm_skip", C.c_void_p),
                ("lm_has_observation", C.c_void_p), ("kp_cap", C.c_int32), ("lm_observable", C.c_void_p), ("kp_landmark_out", C.c_void_p),
                ("kp_outlier", C.c_void_p), ("pose_cw_out", C.c_double * 16), ("n_keypoints", C.c_int32), ("n_matches", C.c_int32),
                ("n_valid", C.c_uint32)]


def _bind():
    L = lib()
    if not getattr(L, "_track_bound", False):
        L.b200_track_local_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TrackParams), C.c_int, C.POINTER(TrackFrame)]
        L.b200_track_stage_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L._track_bound = True
    return L


def camera_intrinsics(camera):
    return CameraIntrinsics(1 if camera.get("model", "perspective") == "equirectangular" else 0,
                            *[float(camera.get(k, 0.0)) for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "cols", "rows")])


class local_map_tracker:
    """One instance owns the pose-optimiser handle of the chain; the matcher handle is the calling thread's (match._matcher)."""

    STAGES = ("undistort_observe", "grid", "candidates", "resolve", "edges", "pose_optimize", "chain")

    def __init__(self, extractor, camera, margin=5.0, lowe_ratio=0.8, hamming_thr=match.HAMMING_DIST_THR_HIGH, ray_cos_thr=0.5,
                 num_trials_robust=2, num_trials=2, num_each_iter=10, grid=(64, 48), img_bounds=None, max_candidates=0, device=0):
        self._L = _bind()
        self.extractor = extractor
        self.device = device
        self._opt = optimize.pose_optimizer(num_trials_robust, num_trials, num_each_iter, device)
        op = extractor.orb_params_
        self._sf = np.ascontiguousarray(op.scale_factors_, np.float32)
        self._isig = np.ascontiguousarray(op.inv_level_sigma_sq_, np.float32)
        p = TrackParams()
        p.cam = camera_intrinsics(camera)
        p.focal_x_baseline = float(camera.get("fxb", 0.0))
        p.monocular = 1 if camera.get("setup", "monocular") == "monocular" else 0
        b = img_bounds if img_bounds is not None else (0.0, camera.get("cols", 0.0), 0.0, camera.get("rows", 0.0))
        p.img_bounds = (C.c_float * 4)(*[float(v) for v in b])
        p.grid_cols, p.grid_rows = int(grid[0]), int(grid[1])
        p.num_levels = int(op.num_levels_)
        p.log_scale_factor = float(op.log_scale_factor_)
        p.scale_factors = self._sf.ctypes.data
        p.inv_level_sigma_sq = self._isig.ctypes.data
        p.margin, p.lowe_ratio, p.hamming_thr, p.ray_cos_thr = float(margin), float(lowe_ratio), int(hamming_thr), float(ray_cos_thr)
        p.num_trials_robust, p.num_trials, p.num_each_iter = int(num_trials_robust), int(num_trials), int(num_each_iter)
        p.max_candidates = int(max_candidates)
        self._prm = p

    def pack(self, frames, kp_cap):
        """frames: dicts(frame, pose_cw (4,4), landmarks=dict(pos_w, mean_normal, min_valid_dist, max_valid_dist, desc[, skip, has_observation])
        [, kp_x_right, kp_landmark]).  Returns (ctypes array, keep-alive list, output arrays)."""
        arr = (TrackFrame * len(frames))()
        keep, outs = [], []
        for i, fr in enumerate(frames):
            lm = fr["landmarks"]
            pos = np.ascontiguousarray(lm["pos_w"], np.float64).reshape(-1, 3)
            n = len(pos)
            a = dict(pose=np.ascontiguousarray(fr["pose_cw"], np.float64).reshape(4, 4), pos=pos,
                     nml=np.ascontiguousarray(lm["mean_normal"], np.float64).reshape(-1, 3),
                     lo=np.ascontiguousarray(lm["min_valid_dist"], np.float32), hi=np.ascontiguousarray(lm["max_valid_dist"], np.float32),
                     desc=np.ascontiguousarray(lm["desc"], np.uint8).reshape(-1, 32),
                     skip=None if lm.get("skip") is None else np.ascontiguousarray(lm["skip"], np.uint8),
                     hobs=None if lm.get("has_observation") is None else np.ascontiguousarray(lm["has_observation"], np.uint8),
                     xr=None if fr.get("kp_x_right") is None else np.ascontiguousarray(fr["kp_x_right"], np.float32),
                     kl=None if fr.get("kp_landmark") is None else np.ascontiguousarray(fr["kp_landmark"], np.int32))
            o = dict(observable=np.zeros(max(n, 1), np.uint8), kp_landmark=np.full(max(kp_cap, 1), -1, np.int32),
                     kp_outlier=np.zeros(max(kp_cap, 1), np.uint8))
            T = arr[i]
            T.frame = int(fr.get("frame", i))
            T.pose_cw = a["pose"].ctypes.data
            T.n_keypoints_in = len(a["xr"]) if a["xr"] is not None else (len(a["kl"]) if a["kl"] is not None else 0)
            T.kp_x_right, T.kp_landmark = ptr(a["xr"]), ptr(a["kl"])
            T.n_landmarks = n
            T.lm_pos_w, T.lm_mean_normal, T.lm_min_valid_dist, T.lm_max_valid_dist = ptr(a["pos"]), ptr(a["nml"]), ptr(a["lo"]), ptr(a["hi"])
            T.lm_desc, T.lm_skip, T.lm_has_observation = ptr(a["desc"]), ptr(a["skip"]), ptr(a["hobs"])
            T.kp_cap = int(kp_cap)
            T.lm_observable, T.kp_landmark_out, T.kp_outlier = ptr(o["observable"]), ptr(o["kp_landmark"]), ptr(o["kp_outlier"])
            keep.append(a)
            outs.append(o)
        return arr, keep, outs

    def run_packed(self, packed):
        arr = packed[0]
        check(self._L.b200_track_local_map(self.extractor._h, match._matcher(self.device), self._opt._h, C.byref(self._prm), len(arr), arr))

    def track(self, frames, kp_cap=None):
        """Returns, per frame: dict(observable bool (n_lm,), kp_landmark int32 (n_kp,), kp_outlier bool (n_kp,), pose_cw (4,4), n_matches, n_valid)."""
        if kp_cap is None:
            b, h, w = self.extractor._shape
            kp_cap = lib().b200_orb_max_keypoints(self.extractor._h, w, h)
        packed = self.pack(frames, kp_cap)
        self.run_packed(packed)
        res = []
        for T, o, a in zip(packed[0], packed[2], packed[1]):
            nk = T.n_keypoints
            res.append(dict(observable=o["observable"][:len(a["pos"])].astype(bool), kp_landmark=o["kp_landmark"][:nk].copy(),
                            kp_outlier=o["kp_outlier"][:nk].astype(bool), pose_cw=np.array(T.pose_cw_out[:]).reshape(4, 4),
                            n_matches=int(T.n_matches), n_valid=int(T.n_valid), n_keypoints=int(nk)))
        return res

    def stage_ms(self):
        out = {}
        for i, nm in enumerate(self.STAGES):
            v = C.c_float()
            check(self._L.b200_track_stage_ms(match._matcher(self.device), i, C.byref(v)))
            out[nm] = v.value
        return out
