"""Seeded synthetic inputs (numpy only) shaped like BASELINE.json's configs: textured frames with corners,
descriptor sets with planted near-duplicates, and local-BA problems (K poses, L landmarks, E observations).

No dataset exists in the container or on the GPU box, so every bench/test input comes from here.
"""
import numpy as np


def _value_noise(rng, h, w, cell):
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw)).astype(np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0, x0 = ys.astype(np.int32), xs.astype(np.int32)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    a = g[y0][:, x0]
    b = g[y0][:, x0 + 1]
    c = g[y0 + 1][:, x0]
    d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


_PAD = 64


def _make_pattern(w, h, seed, n_shapes=None):
    """Padded float32 scene (h+2*PAD, w+2*PAD): 4 octaves of value noise + random filled rectangles/discs."""
    rng = np.random.default_rng(seed)
    H, W = h + 2 * _PAD, w + 2 * _PAD
    img = np.zeros((H, W), np.float32)
    for cell, amp in ((96, 80.0), (32, 45.0), (11, 30.0), (4, 26.0)):
        img += amp * _value_noise(rng, H, W, cell)
    if n_shapes is None:
        n_shapes = max(40, int(400 * (w * h) / (1920 * 1080)))
    for _ in range(n_shapes):
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        s = int(rng.integers(6, 48))
        g = float(rng.integers(0, 256))
        y0, y1, x0, x1 = max(cy - s, 0), min(cy + s, H), max(cx - s, 0), min(cx + s, W)
        if rng.random() < 0.6:
            img[y0:y1, x0:x1] = g
        else:
            yy, xx = np.mgrid[y0:y1, x0:x1]
            m = (yy - cy) ** 2 + (xx - cx) ** 2 <= s * s
            img[y0:y1, x0:x1][m] = g
    return img


def _render(pattern, w, h, seed, shift, noise_sigma):
    dx, dy = int(shift[0]), int(shift[1])
    dx, dy = max(-_PAD, min(_PAD, dx)), max(-_PAD, min(_PAD, dy))
    view = pattern[_PAD + dy:_PAD + dy + h, _PAD + dx:_PAD + dx + w]
    nrng = np.random.default_rng(seed * 7919 + 17 * (dx + 101) + (dy + 103))
    view = view + noise_sigma * nrng.standard_normal(view.shape, dtype=np.float32)
    p = np.pad(view, 1, mode="edge")
    box = sum(p[i:i + h, j:j + w] for i in range(3) for j in range(3)) / np.float32(9.0)
    return np.clip(np.rint(box), 0, 255).astype(np.uint8)


def make_frame(w=1920, h=1080, seed=1234, shift=(0, 0), n_shapes=None, noise_sigma=2.0):
    """u8 HxW frame: 4 octaves of value noise + random filled rectangles/discs + iid noise, then a 3x3 box blur.

    `shift` translates the underlying pattern (pixels), so consecutive frames of a stream overlap and match.
    """
    return _render(_make_pattern(w, h, seed, n_shapes), w, h, seed, shift, noise_sigma)


def make_stream(n_frames, w=1920, h=1080, stream=0, max_step=8):
    """Frames of one synthetic stream: the pattern follows a seeded 2-D random walk (<= max_step px per frame)."""
    rng = np.random.default_rng(99 + stream)
    pattern = _make_pattern(w, h, 1234 + stream)
    pos = np.zeros(2, np.int64)
    frames = []
    for _ in range(n_frames):
        frames.append(_render(pattern, w, h, 1234 + stream, (int(pos[0]), int(pos[1])), 2.0))
        pos = np.clip(pos + rng.integers(-max_step, max_step + 1, 2), -60, 60)
    return frames


def make_descriptor_pair(n1=2000, n2=2000, seed=7, dup_frac=0.6, max_flips=40):
    """Two descriptor sets + angles: dup_frac of set 2 are rows of set 1 with k in [0,max_flips] bit flips."""
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    a1 = (rng.random(n1) * 360).astype(np.float32)
    a2 = (rng.random(n2) * 360).astype(np.float32)
    n_dup = int(dup_frac * n2)
    src = rng.integers(0, n1, n_dup)
    dst = rng.permutation(n2)[:n_dup]
    for s, t in zip(src, dst):
        row = d1[s].copy()
        k = int(rng.integers(0, max_flips + 1))
        bits = rng.choice(256, size=k, replace=False)
        for b in bits:
            row[b >> 3] ^= np.uint8(1 << (b & 7))
        d2[t] = row
        a2[t] = np.float32((a1[s] + rng.normal(0, 8)) % 360)
    valid2 = (rng.random(n2) < 0.9).astype(np.uint8)
    return d1, a1, d2, a2, valid2
