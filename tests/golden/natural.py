"""Helpers shared by the golden generator (make_golden_r2.py) and the tests that consume its fixtures (numpy only)."""
import numpy as np

# Preprocessing.mask_rectangles of the reference's example/aist/equirectangular.yaml: [x_min, x_max, y_min, y_max] as image fractions
AIST_MASK_RECTS = [[0.0, 1.0, 0.0, 0.1], [0.0, 1.0, 0.84, 1.0], [0.0, 0.2, 0.7, 1.0], [0.8, 1.0, 0.7, 1.0]]


def upsample2x(img):
    """Deterministic 2x upsampling in integer arithmetic (rounded averages of the 2 / 4 neighbours, edge replicated): builds the
    3840x1920 input of BASELINE config 2 from the committed 1920x960 natural image without storing 7 MB of pixels."""
    a = img.astype(np.uint16)
    h, w = a.shape
    right = np.concatenate([a[:, 1:], a[:, -1:]], 1)
    down = np.concatenate([a[1:], a[-1:]], 0)
    diag = np.concatenate([right[1:], right[-1:]], 0)
    out = np.empty((2 * h, 2 * w), np.uint16)
    out[0::2, 0::2] = a
    out[0::2, 1::2] = (a + right + 1) >> 1
    out[1::2, 0::2] = (a + down + 1) >> 1
    out[1::2, 1::2] = (a + right + down + diag + 2) >> 2
    return out.astype(np.uint8)
