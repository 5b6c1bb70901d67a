"""Keyframe serialisation formats (data/common.cc:32-81, data/keyframe.cc:191-235, 298-347).
CPU: the JSON forms against a byte-level restatement and the reference's struct layout; GPU: the exported blobs against the oracle's
undistorted keypoints and the extractor's descriptors."""
import struct

import numpy as np
import pytest

from stella_vslam_b200 import data


def test_descriptor_json_round_trip_and_word_order():
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    js = data.convert_descriptors_to_json(d)
    assert len(js) == 50 and all(len(r) == 8 for r in js)
    # `*p` of a uint32_t* over the row bytes on a little-endian host (common.cc:62-65)
    for r in (0, 17, 49):
        assert js[r] == list(struct.unpack("<8I", d[r].tobytes()))
    assert np.array_equal(data.convert_json_to_descriptors(js), d)
    assert data.convert_json_to_descriptors([]).shape == (0, 32)
    with pytest.raises(AssertionError):
        data.convert_json_to_descriptors([[1 << 32] + [0] * 7])


def test_keypoint_json_keeps_pt_ang_oct_only():
    kp = np.zeros(3, data.CV_KEYPOINT_DTYPE)
    kp["x"], kp["y"], kp["size"], kp["angle"], kp["response"], kp["octave"], kp["class_id"] = [1.5, 2.25, 640.0], [3.0, 4.5, 479.0], 31.0, [0.0, 359.5, 12.0], 77.0, [0, 3, 7], 5
    js = data.convert_keypoints_to_json(kp)
    assert js[1] == {"pt": [2.25, 4.5], "ang": 359.5, "oct": 3}
    back = data.convert_json_to_keypoints(js)
    for f in ("x", "y", "angle", "octave"):
        assert np.array_equal(back[f], kp[f])
    assert (back["size"] == 0).all() and (back["response"] == 0).all() and (back["class_id"] == -1).all()   # cv::KeyPoint(x, y, 0, ang, 0, oct, -1)


def test_cv_keypoint_blob_layout():
    # sizeof(cv::KeyPoint) == 28: Point2f pt; float size, angle, response; int octave, class_id  (opencv2/core/types.hpp)
    assert data.CV_KEYPOINT_DTYPE.itemsize == 28
    assert [data.CV_KEYPOINT_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


@pytest.mark.gpu
def test_export_blobs_from_device():
    from oracle import pyoracle as O
    from stella_vslam_b200 import feature
    from workloads import synth
    cam = dict(model="perspective", fx=458.654, fy=457.296, cx=367.215, cy=248.375, k1=-0.28340811, k2=0.07395907, p1=0.00019359, p2=1.76187114e-05,
               k3=0.0, cols=752.0, rows=480.0)
    imgs = np.stack([synth.make_frame(752, 480, seed=60 + i) for i in range(2)])
    ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=2)
    kps, descs = ex.extract_batch(imgs)
    for f in range(2):
        kb, db = data.export_keyframe_blobs(ex, f, cam)
        und, _ = O.undistort_keypoints(cam, kps[f])
        assert len(kb) == len(kps[f]) and np.array_equal(db, descs[f])
        for fld in ("x", "y", "size", "angle", "response", "octave"):
            assert np.array_equal(kb[fld], und[fld]), fld
        assert (kb["class_id"] == -1).all()
        # blob bytes -> library keypoints (from_stmt direction), and the raw (no camera) export keeps the extractor's values
        back = data.keypoints_from_blob(kb.tobytes())
        assert np.array_equal(back, und)
        raw, _ = data.export_keyframe_blobs(ex, f, None)
        for fld in ("x", "y", "response"):
            assert np.array_equal(raw[fld], kps[f][fld])
    from stella_vslam_b200._lib import B200Error
    with pytest.raises(B200Error):
        data.export_keyframe_blobs(ex, 5, cam)
