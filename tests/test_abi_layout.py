"""The Python ctypes mirrors of the C-ABI structs have exactly the layout `include/b200vslam.h` gives them (sizeof and the offset of every
field, from a probe compiled with gcc): a drifted mirror would still load and call, and silently pass garbage."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mirrors():
    from oracle import pyoracle as O
    from stella_vslam_b200 import _lib, optimize, tracking
    return {
        "b200_orb_params_t": _lib.OrbParams, "b200_camera_intrinsics_t": _lib.CameraIntrinsics, "b200_guided_problem_t": _lib.GuidedProblem,
        "b200_pairs_problem_t": _lib.PairsProblem, "b200_lba_problem_t": optimize.LbaProblem, "b200_lba_stats_t": optimize.LbaStats,
        "b200_camera_t": optimize.Camera, "b200_track_params_t": tracking.TrackParams, "b200_track_frame_t": tracking.TrackFrame,
    }, {"orc_keypoint_t": O.Keypoint, "orc_orb_config_t": O.OrbConfig, "orc_guided_t": O.GuidedProblem, "orc_pairs_t": O.PairsProblem,
        "orc_camera_t": O.Camera, "orc_lba_problem_t": O.LbaProblem, "orc_lba_stats_t": O.LbaStats}


def _check(tmp_path, header_dir, header, mirrors):
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{header}"', "int main(void) {"]
    for cname, T in mirrors.items():
        lines.append(f'  printf("{cname} __sizeof__ %zu\\n", sizeof({cname}));')
        for fname, _ in T._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / (header + ".c"), tmp_path / (header + ".exe")
    src.write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-std=c11", "-I", header_dir, str(src), "-o", str(exe)])
    want = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        c, f, v = ln.split()
        want[(c, f)] = int(v)
    for cname, T in mirrors.items():
        assert C.sizeof(T) == want[(cname, "__sizeof__")], (cname, C.sizeof(T), want[(cname, "__sizeof__")])
        for fname, _ in T._fields_:
            assert getattr(T, fname).offset == want[(cname, fname)], (cname, fname)


def test_ctypes_mirrors_match_the_header(tmp_path):
    mirrors, _ = _mirrors()
    _check(tmp_path, os.path.join(ROOT, "include"), "b200vslam.h", mirrors)


def test_oracle_ctypes_mirrors_match_oracle_h(tmp_path):
    _, oracle_mirrors = _mirrors()
    _check(tmp_path, os.path.join(ROOT, "oracle"), "oracle.h", oracle_mirrors)
