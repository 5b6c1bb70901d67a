"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/b200vslam.h declares, and fails
loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

from stella_vslam_b200 import _lib, build as builder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    builder.build()
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "b200vslam.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200_last_error"} - set(_lib.SYMBOLS)
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SYMBOLS) <= declared, sorted(set(_lib.SYMBOLS) - declared)


def test_version_string(L):
    assert L.b200_version().decode().endswith("sm_100a")


def test_no_cpu_fallback(L):
    if L.b200_device_count() > 0:
        pytest.skip("GPU present")
    p = _lib.OrbParams()
    L.b200_orb_default_params(C.byref(p))
    h = C.c_void_p()
    rc = L.b200_orb_create(C.byref(p), C.byref(h))
    assert rc == _lib.ERR_CUDA
    assert b"no CPU fallback" in L.b200_last_error()
    with pytest.raises(_lib.B200Error):
        from stella_vslam_b200 import feature
        feature.orb_extractor(feature.orb_params(), 800)


def test_default_params(L):
    p = _lib.OrbParams()
    L.b200_orb_default_params(C.byref(p))
    assert (round(p.scale_factor, 4), p.num_levels, p.ini_fast_thr, p.min_fast_thr, p.min_area) == (1.2, 8, 20, 7, 800)


def test_orb_params_recurrence():
    # test/stella_vslam/feature/orb_params.cc:29-71 (EXPECT_FLOAT_EQ on the recurrences)
    import numpy as np
    from oracle import pyoracle as O
    from stella_vslam_b200 import feature
    prm = feature.orb_params("ORB setting for test")
    sf, inv, sig, isig = O.scale_factors(1.2, 8)
    assert np.array_equal(prm.scale_factors_, sf) and np.array_equal(prm.inv_scale_factors_, inv)
    assert np.array_equal(prm.level_sigma_sq_, sig) and np.array_equal(prm.inv_level_sigma_sq_, isig)
    y = feature.orb_params.from_yaml({"scale_factor": 1.3, "num_levels": 4, "ini_fast_threshold": 12})
    assert y.num_levels_ == 4 and y.ini_fast_thr_ == 12 and y.min_fast_thr_ == 7
