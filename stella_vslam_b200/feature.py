"""Host-side mirror of the reference's feature:: surface (src/stella_vslam/feature/orb_params.h, orb_extractor.h).

Same names, argument meaning and error behaviour as the C++ classes, on top of the C ABI (include/b200vslam.h):
  feature::orb_params      orb_params.cc:12-71
  feature::orb_extractor   orb_extractor.h:51-71, orb_extractor.cc:16-136
cv::Mat inputs become 2-D uint8 numpy arrays; std::vector<cv::KeyPoint> becomes a structured array with the
cv::KeyPoint fields the reference fills (class_id is always -1 and is not stored).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, CameraIntrinsics, OrbParams, check, lib, ptr


class descriptor_type:  # feature/orb_extractor.h:17-44
    ORB = 0
    HASH_SIFT = 1


class orb_params:
    """feature::orb_params: pyramid scale tables built with the reference's float recurrences (orb_params.cc:37-71)."""

    def __init__(self, name="default ORB feature extraction setting", scale_factor=1.2, num_levels=8, ini_fast_thr=20,
                 min_fast_thr=7):
        self.name_ = name
        self.scale_factor_ = np.float32(scale_factor)
        self.log_scale_factor_ = np.float32(np.log(np.float32(scale_factor)))
        self.num_levels_ = int(num_levels)
        self.ini_fast_thr_ = int(ini_fast_thr)
        self.min_fast_thr_ = int(min_fast_thr)
        sf, inv, sig, isig = [np.float32(1.0)], [np.float32(1.0)], [np.float32(1.0)], [np.float32(1.0)]
        s = np.float32(1.0)
        for _ in range(1, self.num_levels_):
            s = self.scale_factor_ * s
            sf.append(s)
            inv.append((np.float32(1.0) / self.scale_factor_) * inv[-1])
            sig.append(s * s)
            isig.append(np.float32(1.0) / (s * s))
        self.scale_factors_ = np.array(sf, np.float32)
        self.inv_scale_factors_ = np.array(inv, np.float32)
        self.level_sigma_sq_ = np.array(sig, np.float32)
        self.inv_level_sigma_sq_ = np.array(isig, np.float32)

    @classmethod
    def from_yaml(cls, node):
        """orb_params(const YAML::Node&) (orb_params.cc:22-27): `node` is the dict of the `Feature:` block."""
        return cls(node.get("name", "default ORB feature extraction setting"), node.get("scale_factor", 1.2),
                   node.get("num_levels", 8), node.get("ini_fast_threshold", 20), node.get("min_fast_threshold", 7))

    def to_json(self):
        return {"name": self.name_, "scale_factor": float(self.scale_factor_), "num_levels": self.num_levels_,
                "ini_fast_threshold": self.ini_fast_thr_, "min_fast_threshold": self.min_fast_thr_}


class orb_extractor:
    """feature::orb_extractor on the GPU.  `extract` handles one frame like the reference; `extract_batch` takes a
    stack of same-sized frames (the B200-native entry: one launch sequence for the whole batch)."""

    def __init__(self, orb_params_, min_area, desc_type=descriptor_type.ORB, mask_rects=(), device=0, max_batch=1):
        if desc_type == descriptor_type.HASH_SIFT:
            # orb_extractor.cc:117-122 without USE_CUDA_EFFICIENT_DESCRIPTORS
            raise RuntimeError("cuda_efficient_features is not available")
        if desc_type != descriptor_type.ORB:
            raise RuntimeError("Invalid descriptor_type")  # orb_extractor.cc:125
        self.orb_params_ = orb_params_
        self.mask_rects_ = [list(map(float, r)) for r in mask_rects]
        self.image_pyramid_ = []
        self._rects = np.ascontiguousarray(np.array(self.mask_rects_, np.float32).reshape(-1, 4))
        p = OrbParams()
        lib().b200_orb_default_params(C.byref(p))
        p.scale_factor = float(orb_params_.scale_factor_)
        p.num_levels = orb_params_.num_levels_
        p.ini_fast_thr = orb_params_.ini_fast_thr_
        p.min_fast_thr = orb_params_.min_fast_thr_
        p.min_area = int(min_area)
        p.n_mask_rects = self._rects.shape[0]
        p.mask_rects = self._rects.ctypes.data_as(C.POINTER(C.c_float)) if self._rects.size else None
        p.device = device
        p.max_batch = max_batch
        self._h = C.c_void_p()
        check(lib().b200_orb_create(C.byref(p), C.byref(self._h)))
        self._shape = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().b200_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference signature: extract(in_image, in_image_mask, keypts, out_descriptors) ------------------------------
    def extract(self, in_image, in_image_mask=None):
        """Returns (keypts, descriptors).  Empty image -> ([], None-like empty) as orb_extractor.cc:30-32,72-74."""
        img = np.asarray(in_image)
        if img.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        k, d = self.extract_batch(img[None], in_image_mask)
        return k[0], d[0]

    def extract_batch(self, images, in_image_mask=None, cap=None):
        images = np.asarray(images)
        assert images.dtype == np.uint8 and images.ndim == 3, "image.type() == CV_8UC1"  # orb_extractor.cc:36
        images = np.ascontiguousarray(images)
        b, h, w = images.shape
        if b == 0 or h == 0 or w == 0:
            return [], []
        mask = None
        if in_image_mask is not None and np.asarray(in_image_mask).size:
            mask = np.ascontiguousarray(in_image_mask, np.uint8)
            assert mask.shape == (h, w), "image mask must have the image's size"
        if cap is None:
            cap = check_pos(lib().b200_orb_max_keypoints(self._h, w, h))
        kps = np.zeros((b, max(cap, 1)), KP_DTYPE)
        desc = np.zeros((b, max(cap, 1), 32), np.uint8)
        counts = np.zeros(b, np.int32)
        check(lib().b200_orb_extract(self._h, ptr(images), w, h, images.strides[1], images.strides[0], b, ptr(mask),
                                     mask.strides[0] if mask is not None else 0, ptr(kps), ptr(desc), cap, ptr(counts)))
        self._shape = (b, h, w)
        self._last_images = images
        return [kps[f, :counts[f]].copy() for f in range(b)], [desc[f, :counts[f]].copy() for f in range(b)]

    def image_pyramid(self, frame=0):
        """orb_extractor::image_pyramid_ (orb_extractor.h:71) of the last extract, copied back to the host."""
        if self._shape is None:
            return []
        out = [self._last_images[frame]]
        for l in range(1, self.orb_params_.num_levels_):
            w, h = C.c_int(), C.c_int()
            check(lib().b200_orb_level_info(self._h, l, C.byref(w), C.byref(h), None, None))
            lv = np.empty((h.value, w.value), np.uint8)
            check(lib().b200_orb_pyramid_level_host(self._h, frame, l, ptr(lv), lv.strides[0]))
            out.append(lv)
        self.image_pyramid_ = out
        return out

    def undistort_keypoints(self, camera, dist_keypts, want_bearings=True):
        """camera::perspective / equirectangular::undistort_keypoints + camera::base::convert_keypoints_to_bearings
        (camera/perspective.cc:245-275, 117-122; equirectangular.cc:42-49; called right after extract, system.cc:386-395).
        camera: dict(model="perspective"|"equirectangular", fx, fy, cx, cy, k1, k2, p1, p2, k3, cols, rows).
        Returns (undist_keypts, bearings (n, 3) float64 | None)."""
        kps = np.ascontiguousarray(dist_keypts, KP_DTYPE)
        n = len(kps)
        cam = CameraIntrinsics(1 if camera.get("model", "perspective") == "equirectangular" else 0, *[float(camera.get(k, 0.0)) for k in
                               ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "cols", "rows")])
        out = np.zeros(n, KP_DTYPE)
        bearings = np.zeros((n, 3)) if want_bearings else None
        if n:
            check(lib().b200_keypoints_undistort(self._h, C.byref(cam), ptr(kps), n, ptr(out), ptr(bearings)))
        return out, bearings

    def can_observe(self, camera, pose_cw, landmarks, ray_cos_thr=0.5, img_bounds=None):
        """data::frame::can_observe (data/frame.cc:59-84) for the local landmarks (tracking_module.cc:559-594).
        landmarks: dict(pos_w (n,3), mean_normal (n,3), min_valid_dist (n,), max_valid_dist (n,)).
        Returns dict(observable bool (n,), reproj (n,2), x_right (n,), pred_scale_level (n,)) -- the inputs of
        match.projection.match_frame_and_landmarks."""
        pos = np.ascontiguousarray(landmarks["pos_w"], np.float64).reshape(-1, 3)
        nml = np.ascontiguousarray(landmarks["mean_normal"], np.float64).reshape(-1, 3)
        lo = np.ascontiguousarray(landmarks["min_valid_dist"], np.float32)
        hi = np.ascontiguousarray(landmarks["max_valid_dist"], np.float32)
        n = len(pos)
        cam = CameraIntrinsics(1 if camera.get("model", "perspective") == "equirectangular" else 0, *[float(camera.get(k, 0.0)) for k in
                               ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "cols", "rows")])
        bounds = np.ascontiguousarray(img_bounds if img_bounds is not None else (0.0, camera.get("cols", 0.0), 0.0, camera.get("rows", 0.0)), np.float32)
        pose = np.ascontiguousarray(pose_cw, np.float64).reshape(4, 4)
        ok, rp = np.zeros(max(n, 1), np.uint8), np.zeros((max(n, 1), 2))
        xr, lv = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.uint32)
        check(lib().b200_frame_can_observe(self._h, C.byref(cam), float(camera.get("fxb", 0.0)), ptr(bounds), ptr(pose), n, ptr(pos), ptr(nml), ptr(lo),
                                           ptr(hi), float(ray_cos_thr), int(self.orb_params_.num_levels_), float(self.orb_params_.log_scale_factor_),
                                           ptr(ok), ptr(rp), ptr(xr), ptr(lv)))
        return dict(observable=ok[:n].astype(bool), reproj=rp[:n], x_right=xr[:n], pred_scale_level=lv[:n])

    def convert_to_grayscale(self, img, in_color_order="BGR"):
        """util::convert_to_grayscale (util/image_converter.cc:8-39): (h, w, 3|4) uint8 -> (h, w) uint8; a 1-channel image or
        in_color_order == "Gray" is returned unchanged."""
        img = np.asarray(img)
        if img.ndim == 2 or in_color_order == "Gray":
            return img
        img = np.ascontiguousarray(img, np.uint8)
        h, w, c = img.shape
        out = np.empty((h, w), np.uint8)
        check(lib().b200_convert_to_grayscale(self._h, ptr(img), w, h, img.strides[0], c, 1 if in_color_order == "RGB" else 0, ptr(out), out.strides[0]))
        return out

    def enable_timing(self, on=True):
        check(lib().b200_orb_enable_timing(self._h, int(on)))

    def stage_ms(self):
        out = []
        for s in range(6):
            v = C.c_float()
            check(lib().b200_orb_stage_ms(self._h, s, C.byref(v)))
            out.append(v.value)
        return out


def check_pos(v):
    if v < 0:
        check(v)
    return v
