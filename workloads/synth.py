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


# ---------------------------------------------------------------------------------------------------------------------
# local-BA problems: K keyframes on an arc, L landmarks in the frustum union, E observations
# ---------------------------------------------------------------------------------------------------------------------
KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, fxb=386.1448, cols=1241, rows=376)  # example/kitti/KITTI_stereo_00-02.yaml


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def make_ba_problem(n_poses=50, n_fixed=10, n_points=10000, seed=0, model="stereo", outlier_frac=0.05, pixel_sigma=1.0,
                    min_obs=4, max_obs=8, arc_m=200.0, perturb=True):
    """Synthetic local-BA problem in the flattened layout of the C ABI.

    model: "mono" | "stereo" (perspective, KITTI intrinsics) | "equirect" (3840x1920).
    Returns dict(pose_cw (K,4,4), pose_fixed (K,), points (L,3), e_pose, e_point, e_cam, e_obs (E,3) f32,
                 e_inv_sigma_sq f32, e_delta f32, cams (list of dict), gt_pose_cw, gt_points).
    """
    rng = np.random.default_rng(seed)
    K, L = n_poses, n_points
    equirect = model == "equirect"
    cam = dict(model=1 if equirect else 0, fx=KITTI["fx"], fy=KITTI["fy"], cx=KITTI["cx"], cy=KITTI["cy"], fxb=KITTI["fxb"],
               cols=3840.0 if equirect else float(KITTI["cols"]), rows=1920.0 if equirect else float(KITTI["rows"]))
    # keyframes on a gentle arc, looking along the direction of travel
    s = np.linspace(0, arc_m, K)
    radius = 4 * arc_m
    ang = s / radius
    centers = np.stack([radius * np.sin(ang), 0.3 * np.sin(s / 15.0), radius * (1 - np.cos(ang))], 1)
    gt_pose = np.zeros((K, 4, 4))
    for k in range(K):
        Rwc = _rot_y(ang[k] + 0.02 * rng.standard_normal()) @ _rodrigues(0.01 * rng.standard_normal(3))
        Rcw = Rwc.T
        gt_pose[k, :3, :3] = Rcw
        gt_pose[k, :3, 3] = -Rcw @ centers[k]
        gt_pose[k, 3, 3] = 1
    sf = np.float32(1.0)
    inv_sigma = [np.float32(1.0)]
    for _ in range(1, 8):
        sf = np.float32(1.2) * sf
        inv_sigma.append(np.float32(1.0) / (sf * sf))
    inv_sigma = np.array(inv_sigma, np.float32)
    sigma_lvl = 1.0 / np.sqrt(inv_sigma.astype(np.float64))

    def project(Tcw, pw):
        pc = pw @ Tcw[:3, :3].T + Tcw[:3, 3]
        if equirect:
            th = np.arctan2(pc[:, 0], pc[:, 2])
            ph = -np.arcsin(pc[:, 1] / np.linalg.norm(pc, axis=1))
            uv = np.stack([cam["cols"] * (0.5 + th / (2 * np.pi)), cam["rows"] * (0.5 - ph / np.pi)], 1)
            vis = np.linalg.norm(pc, axis=1) > 2.0
            return uv, pc, vis
        z = pc[:, 2]
        uv = np.stack([cam["fx"] * pc[:, 0] / z + cam["cx"], cam["fy"] * pc[:, 1] / z + cam["cy"]], 1)
        vis = (z > 3.0) & (z < 80.0) & (uv[:, 0] > 0) & (uv[:, 0] < cam["cols"]) & (uv[:, 1] > 0) & (uv[:, 1] < cam["rows"])
        return uv, pc, vis

    # landmarks: sampled in front of random keyframes at 5..60 m
    pts = np.zeros((L, 3))
    owner = rng.integers(0, K, L)
    depth = rng.uniform(5, 60, L)
    u = rng.uniform(0.05, 0.95, L) * (KITTI["cols"] if not equirect else 1241)
    v = rng.uniform(0.05, 0.95, L) * (KITTI["rows"] if not equirect else 376)
    for l in range(L):
        Tcw = gt_pose[owner[l]]
        pc = np.array([(u[l] - KITTI["cx"]) / KITTI["fx"] * depth[l], (v[l] - KITTI["cy"]) / KITTI["fy"] * depth[l], depth[l]])
        pts[l] = Tcw[:3, :3].T @ (pc - Tcw[:3, 3])
    e_pose, e_point, e_obs, e_isq = [], [], [], []
    order = np.argsort(np.abs(np.arange(K)[None, :] - owner[:, None]), axis=1)  # nearest keyframes first
    proj = [project(gt_pose[k], pts) for k in range(K)]
    for l in range(L):
        want = int(rng.integers(min_obs, max_obs + 1))
        got = 0
        for k in order[l]:
            uv, pc, vis = proj[k]
            if not vis[l]:
                continue
            lvl = int(rng.integers(0, 8))
            noise = pixel_sigma * sigma_lvl[lvl] * rng.standard_normal(3)
            x, y = uv[l, 0] + noise[0], uv[l, 1] + noise[1]
            xr = -1.0
            if model == "stereo" and rng.random() < 0.85:
                xr = uv[l, 0] - cam["fxb"] / pc[l, 2] + noise[2]
                if xr < 0:
                    xr = -1.0
            if rng.random() < outlier_frac:
                x += rng.choice([-1, 1]) * rng.uniform(15, 30)
                y += rng.choice([-1, 1]) * rng.uniform(15, 30)
            e_pose.append(k)
            e_point.append(l)
            e_obs.append((x, y, xr))
            e_isq.append(inv_sigma[lvl])
            got += 1
            if got >= want:
                break
    E = len(e_pose)
    chi = np.float32(np.sqrt(np.float32(5.99146))) if model != "stereo" else np.float32(np.sqrt(np.float32(7.81473)))
    pose_fixed = np.zeros(K, np.uint8)
    pose_fixed[:n_fixed] = 1   # the oldest keyframes play the "fixed" role (observers outside the local window)
    pose0, pts0 = gt_pose.copy(), pts.copy()
    if perturb:
        for k in range(K):
            if pose_fixed[k]:
                continue
            dR = _rodrigues(np.deg2rad(0.5) * rng.standard_normal(3) / np.sqrt(3))
            Rn = dR @ gt_pose[k, :3, :3]                     # 0.5 deg about the camera centre, 5 cm of centre noise
            cn = centers[k] + 0.05 * rng.standard_normal(3) / np.sqrt(3)
            pose0[k, :3, :3] = Rn
            pose0[k, :3, 3] = -Rn @ cn
        pts0 = pts + 0.01 * depth[:, None] * rng.standard_normal((L, 3))
    return dict(pose_cw=pose0, pose_fixed=pose_fixed, points=pts0, point_fixed=None, e_pose=np.array(e_pose, np.int32),
                e_point=np.array(e_point, np.int32), e_cam=np.zeros(E, np.uint8), e_obs=np.array(e_obs, np.float32).reshape(E, 3),
                e_inv_sigma_sq=np.array(e_isq, np.float32), e_delta=np.full(E, chi, np.float32), e_robust=None,
                e_can_be_outlier=None, cams=[cam], gt_pose_cw=gt_pose, gt_points=pts)


def make_pose_problem(seed=0, n_obs=1500, model="stereo", outlier_frac=0.1, pixel_sigma=1.0, rot_deg=1.0, trans_m=0.15):
    """One frame for optimize::pose_optimizer in the flattened layout of b200_lba_problem_t: ONE free pose (perturbed), the
    `n_obs` landmarks it observes (fixed, exact) and one edge per observation with level-dependent noise and gross outliers."""
    rng = np.random.default_rng(seed)
    equirect = model == "equirect"
    cam = dict(model=1 if equirect else 0, fx=KITTI["fx"], fy=KITTI["fy"], cx=KITTI["cx"], cy=KITTI["cy"], fxb=KITTI["fxb"],
               cols=3840.0 if equirect else float(KITTI["cols"]), rows=1920.0 if equirect else float(KITTI["rows"]))
    Rcw = _rot_y(0.3 * rng.standard_normal()) @ _rodrigues(0.05 * rng.standard_normal(3))
    tcw = rng.normal(0, 2.0, 3)
    gt = np.eye(4)
    gt[:3, :3], gt[:3, 3] = Rcw, tcw
    depth = rng.uniform(4, 60, n_obs)
    if equirect:
        d = rng.standard_normal((n_obs, 3))
        pc = d / np.linalg.norm(d, axis=1, keepdims=True) * depth[:, None]
    else:
        u, v = rng.uniform(20, cam["cols"] - 20, n_obs), rng.uniform(20, cam["rows"] - 20, n_obs)
        pc = np.stack([(u - cam["cx"]) / cam["fx"] * depth, (v - cam["cy"]) / cam["fy"] * depth, depth], 1)
    pw = (pc - tcw) @ Rcw                                # Rcw^T (pc - tcw)
    inv_sigma = (np.float32(1.0) / np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))])).astype(np.float32) ** 2).astype(np.float32)
    lvl = rng.integers(0, 8, n_obs)
    sig = pixel_sigma / np.sqrt(inv_sigma[lvl].astype(np.float64))
    if equirect:
        th, ph = np.arctan2(pc[:, 0], pc[:, 2]), -np.arcsin(pc[:, 1] / np.linalg.norm(pc, axis=1))
        x, y = cam["cols"] * (0.5 + th / (2 * np.pi)), cam["rows"] * (0.5 - ph / np.pi)
    else:
        x, y = cam["fx"] * pc[:, 0] / pc[:, 2] + cam["cx"], cam["fy"] * pc[:, 1] / pc[:, 2] + cam["cy"]
    xr = np.full(n_obs, -1.0)
    if model == "stereo":
        has = rng.random(n_obs) < 0.8
        xr[has] = (x - cam["fxb"] / pc[:, 2] + sig * rng.standard_normal(n_obs))[has]
        xr[xr < 0] = -1.0
    x, y = x + sig * rng.standard_normal(n_obs), y + sig * rng.standard_normal(n_obs)
    bad = rng.random(n_obs) < outlier_frac
    x[bad] += rng.choice([-1, 1], bad.sum()) * rng.uniform(10, 60, bad.sum())
    y[bad] += rng.choice([-1, 1], bad.sum()) * rng.uniform(10, 60, bad.sum())
    dR = _rodrigues(np.deg2rad(rot_deg) * rng.standard_normal(3) / np.sqrt(3))
    pose0 = np.eye(4)
    pose0[:3, :3] = dR @ Rcw
    pose0[:3, 3] = dR @ tcw + trans_m * rng.standard_normal(3) / np.sqrt(3)
    chi = np.float32(np.sqrt(np.float32(7.81473))) if model == "stereo" else np.float32(np.sqrt(np.float32(5.99146)))  # setup-type dependent, :99-101
    return dict(pose_cw=pose0[None], pose_fixed=np.zeros(1, np.uint8), points=pw, point_fixed=np.ones(n_obs, np.uint8),
                e_pose=np.zeros(n_obs, np.int32), e_point=np.arange(n_obs, dtype=np.int32), e_cam=np.zeros(n_obs, np.uint8),
                e_obs=np.stack([x, y, xr], 1).astype(np.float32), e_inv_sigma_sq=inv_sigma[lvl], e_delta=np.full(n_obs, chi, np.float32),
                e_robust=None, e_can_be_outlier=None, cams=[cam], gt_pose_cw=gt, gt_outlier=bad)


def make_guided_problem(seed, n_train=2000, n_queries=1500, mode=0, stereo=False, width=640, height=480, margin=5.0, num_levels=8,
                        scale_factor=1.2):
    """A synthetic problem for the grid-guided projection matchers (match.projection): a frame with `n_train` keypoints and
    `n_queries` landmarks that reproject near some of them.  Built to exercise every gate: several landmarks compete for one
    keypoint (the sequential occupancy matters), near-duplicate descriptors sit next to each other (ratio test), octaves fall
    outside the level window, some keypoints are pre-occupied, some landmarks are invalid, undistorted bounds are fractional and
    a few keypoints / reprojections fall outside them."""
    rng = np.random.default_rng(seed)
    sf = np.float32(1.0) * np.cumprod(np.concatenate([[np.float32(1.0)], np.full(num_levels - 1, np.float32(scale_factor))])).astype(np.float32)
    bounds = (np.float32(-11.37), np.float32(width + 9.21), np.float32(-7.9), np.float32(height + 6.53))
    # keypoints: uniform background + tight clusters
    n_cl = n_train // 3
    centers = rng.uniform([0, 0], [width, height], (max(n_cl // 12, 1), 2))
    pts = np.concatenate([rng.uniform([bounds[0] - 3, bounds[2] - 3], [bounds[1] + 3, bounds[3] + 3], (n_train - n_cl, 2)),
                          centers[rng.integers(0, len(centers), n_cl)] + rng.normal(0, 4.0, (n_cl, 2))])
    pts = pts[rng.permutation(n_train)].astype(np.float32)
    octave = rng.choice(num_levels, n_train, p=np.array([.3, .22, .16, .12, .08, .06, .04, .02])).astype(np.uint8)
    angle = rng.uniform(0, 360, n_train).astype(np.float32)
    desc = rng.integers(0, 256, (n_train, 32), dtype=np.uint8)
    # near-duplicate descriptors between spatial neighbours (sorted by x so duplicates are usually inside one window)
    order = np.argsort(pts[:, 0], kind="stable")
    for a, b in zip(order[0:n_train - 1:7], order[1:n_train:7]):
        noise = rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) \
            & rng.integers(0, 256, 32, dtype=np.uint8)
        desc[b] = desc[a] ^ noise
        if rng.random() < 0.5:
            octave[b] = octave[a]
    src = rng.integers(0, n_train, n_queries)
    src[1::5] = src[0:n_queries - 1:5][:len(src[1::5])]          # two landmarks on one keypoint
    strength = rng.integers(1, 6, n_queries)                     # AND of k random bytes: ~ 256 / 2^k flipped bits
    q_desc = desc[src].copy()
    for k in range(1, 6):  # per-landmark noise level
        sel = strength == k
        fl = rng.integers(0, 256, (sel.sum(), 32), dtype=np.uint8)
        for _ in range(k):
            fl &= rng.integers(0, 256, (sel.sum(), 32), dtype=np.uint8)
        q_desc[sel] ^= fl
    level = np.clip(octave[src].astype(np.int64) + rng.integers(-2, 3, n_queries), 0, num_levels - 1)
    q_xy = pts[src].astype(np.float64) + rng.normal(0, 2.5, (n_queries, 2))
    far = rng.random(n_queries) < 0.03
    q_xy[far] += rng.normal(0, 400, (far.sum(), 2))
    q_margin = (np.float32(margin) * sf[level]).astype(np.float32)
    lo, hi = np.maximum(0, level - 1), np.minimum(num_levels - 1, level + 1)
    unchecked = rng.random(n_queries) < 0.05
    lo[unchecked], hi[unchecked] = -1, -1
    prob = dict(t_x=pts[:, 0].copy(), t_y=pts[:, 1].copy(), t_octave=octave, t_angle=angle, t_desc=desc,
                t_occupied=(rng.random(n_train) < 0.08).astype(np.uint8), bounds=bounds, grid=(64, 48), scale_factors=sf,
                q_desc=q_desc, q_x=q_xy[:, 0].astype(np.float32), q_y=q_xy[:, 1].astype(np.float32), q_margin=q_margin,
                q_min_level=lo.astype(np.int8), q_max_level=hi.astype(np.int8),
                q_angle=((angle[src] + rng.normal(0, 18, n_queries)) % 360).astype(np.float32),
                q_valid=(rng.random(n_queries) > 0.1).astype(np.uint8),
                q_reproj=q_xy.copy(), inv_level_sigma_sq=(np.float32(1.0) / (sf * sf)).astype(np.float32),
                do_reprojection_matching=(mode == 3))
    if mode == 4:  # area::match_in_consistent_area: level-0 keypoints only, one integer margin
        prob.update(q_min_level=np.zeros(n_queries, np.int8), q_max_level=np.zeros(n_queries, np.int8),
                    q_margin=np.full(n_queries, np.float32(int(margin * 4))), q_valid=(level == 0).astype(np.uint8))
    if stereo:
        xr = (pts[:, 0] - rng.uniform(2, 60, n_train)).astype(np.float32)
        xr[rng.random(n_train) < 0.3] = -1.0                    # no stereo match for this keypoint
        prob["t_x_right"] = xr
        prob["q_x_right"] = (xr[src] + rng.normal(0, q_margin * 0.6)).astype(np.float32)
    return prob


def make_keyframe_pair(seed, n1=2000, n2=2000, stereo=False, n_nodes=150, num_levels=8, scale_factor=1.2, bearing_noise=1.5e-3):
    """Two keyframes looking at the same synthetic points, for the all-pairs matchers with greedy state (bow_tree::*,
    robust::match_for_triangulation).  Returns (keyfrm_1, keyfrm_2, geometry): dicts with desc, angle, octave, bearings (unit,
    f64), node (BoW node id), no_landmark / has_landmark (u8), stereo (u8 | None), scale_factors; geometry holds E_12 with
    bearing_1 . (E_12 bearing_2) = 0 for true correspondences, the epipole of keyframe 1 in keyframe 2, and the pair list.
    Corresponding keypoints share descriptors up to noise of varied strength; some rows have two look-alike candidates (ratio
    test), some correspondences violate the epipolar constraint, some sit next to the epipole."""
    rng = np.random.default_rng(seed)
    sf = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(num_levels - 1, np.float32(scale_factor))])).astype(np.float32)
    n_common = int(0.7 * min(n1, n2))
    # relative pose: x1 = R_12 x2 + t_12  (mostly forward motion, so the epipole lies inside the image)
    R_12 = _rodrigues(rng.normal(0, 0.03, 3))
    t_12 = np.array([0.15, -0.05, 0.6]) + rng.normal(0, 0.02, 3)
    tx = np.array([[0, -t_12[2], t_12[1]], [t_12[2], 0, -t_12[0]], [-t_12[1], t_12[0], 0]])
    E_12 = tx @ R_12
    pts2 = np.stack([rng.uniform(-4, 4, n_common), rng.uniform(-3, 3, n_common), rng.uniform(4, 30, n_common)], 1)
    near_epipole = rng.random(n_common) < 0.04           # along the baseline as seen from camera 2
    c1_in_2 = -R_12.T @ t_12
    pts2[near_epipole] = c1_in_2 * rng.uniform(8, 30, (near_epipole.sum(), 1)) + rng.normal(0, 0.15, (near_epipole.sum(), 3))
    pts1 = pts2 @ R_12.T + t_12

    def unit(v):
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    def side(n, pts, perm):
        b = unit(np.stack([rng.uniform(-0.8, 0.8, n), rng.uniform(-0.6, 0.6, n), np.ones(n)], 1))
        b[perm] = unit(unit(pts) + rng.normal(0, bearing_noise, pts.shape))
        return b

    perm1, perm2 = rng.permutation(n1)[:n_common], rng.permutation(n2)[:n_common]
    bearings1, bearings2 = side(n1, pts1, perm1), side(n2, pts2, perm2)
    wrong = rng.random(n_common) < 0.1                   # correspondences that break the epipolar constraint
    bearings2[perm2[wrong]] = unit(bearings2[perm2[wrong]] + rng.normal(0, 0.05, (wrong.sum(), 3)))
    desc1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    desc2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    strength = rng.integers(2, 6, n_common)
    fl = np.full((n_common, 32), 255, np.uint8)
    for k in range(6):
        r = rng.integers(0, 256, (n_common, 32), dtype=np.uint8)
        fl = np.where((strength > k)[:, None], fl & r, fl)
    desc2[perm2] = desc1[perm1] ^ fl
    octave1 = rng.choice(num_levels, n1, p=np.array([.3, .22, .16, .12, .08, .06, .04, .02])).astype(np.uint8)
    octave2 = octave1[rng.integers(0, n1, n2)]
    octave2[perm2] = octave1[perm1]
    angle1 = rng.uniform(0, 360, n1).astype(np.float32)
    angle2 = rng.uniform(0, 360, n2).astype(np.float32)
    angle2[perm2] = (angle1[perm1] + rng.normal(0, 14, n_common)) % 360
    node1, node2 = rng.integers(0, n_nodes, n1).astype(np.int32), rng.integers(0, n_nodes, n2).astype(np.int32)
    same_node = rng.random(n_common) < 0.9
    node2[perm2[same_node]] = node1[perm1[same_node]]
    # look-alikes: a second candidate in the same node with a similar descriptor and the same bearing (passes every gate)
    free2 = np.setdiff1d(np.arange(n2), perm2)
    for k, j in zip(rng.permutation(n_common)[:len(free2) // 2], free2):
        noise = rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) \
            & rng.integers(0, 256, 32, dtype=np.uint8)
        desc2[j] = desc2[perm2[k]] ^ noise
        bearings2[j], angle2[j], node2[j] = bearings2[perm2[k]], angle2[perm2[k]], node2[perm2[k]]
    has_lm1, has_lm2 = (rng.random(n1) < 0.5).astype(np.uint8), (rng.random(n2) < 0.5).astype(np.uint8)

    def kf(desc, angle, octave, bearings, node, has_lm, n):
        return dict(desc=desc, angle=angle, octave=octave, bearings=np.ascontiguousarray(bearings), node=node, has_landmark=has_lm,
                    no_landmark=(1 - has_lm).astype(np.uint8), stereo=(rng.random(n) < 0.4).astype(np.uint8) if stereo else None, scale_factors=sf)

    geometry = dict(E_12=E_12, epiplane_in_keyfrm_2=c1_in_2 / np.linalg.norm(c1_in_2), valid_epiplane=True, perm1=perm1, perm2=perm2)
    return kf(desc1, angle1, octave1, bearings1, node1, has_lm1, n1), kf(desc2, angle2, octave2, bearings2, node2, has_lm2, n2), geometry


def make_stereo_pair(w=752, h=480, seed=77, disparities=(9, 23, 41), noise_sigma=2.0):
    """(left, right) rectified frames: the right view shows the same pattern moved left by a disparity that differs per
    horizontal band, with independent sensor noise, so match::stereo finds sub-pixel disparities around those values."""
    pattern = _make_pattern(w, h, seed)
    left = _render(pattern, w, h, seed, (0, 0), noise_sigma)
    right = np.empty_like(left)
    bands = np.linspace(0, h, len(disparities) + 1).astype(int)
    for k, d in enumerate(disparities):
        right[bands[k]:bands[k + 1]] = _render(pattern, w, h, seed + 1, (int(d), 0), noise_sigma)[bands[k]:bands[k + 1]]
    return left, right


def make_tracking_frame(kps, desc, camera, scale_factors, seed=0, stereo=False, landmark_frac=0.7, clutter_frac=0.3, pre_matched_frac=0.15,
                        pixel_sigma=1.0, rot_deg=0.5, trans_m=0.05, max_flips=30):
    """A local map for one extracted frame (track_local_map workload): most keypoints get a landmark at a random depth (descriptor = the
    keypoint's with a few flipped bits, position off by ~pixel_sigma pixels), plus clutter landmarks (random descriptors; some behind the
    camera or outside the image), a few landmarks the frame already carries (kp_landmark, skipped by the search) and a few without
    observations.  The pose handed to the tracker is the true pose perturbed by rot_deg / trans_m.  Perspective and equirectangular
    cameras (the keypoints are taken as undistorted)."""
    rng = np.random.default_rng(seed)
    n_kp = len(kps)
    sf = np.asarray(scale_factors, np.float64)
    fx, fy, cx, cy = camera.get("fx", 1.0), camera.get("fy", 1.0), camera.get("cx", 0.0), camera.get("cy", 0.0)
    Rcw = _rot_y(0.2 * rng.standard_normal()) @ _rodrigues(0.05 * rng.standard_normal(3))
    tcw = rng.normal(0, 1.0, 3)
    center = -Rcw.T @ tcw
    pick = np.nonzero(rng.random(n_kp) < landmark_frac)[0]
    z = rng.uniform(4, 40, len(pick))
    u = kps["x"][pick].astype(np.float64) + pixel_sigma * rng.standard_normal(len(pick))
    v = kps["y"][pick].astype(np.float64) + pixel_sigma * rng.standard_normal(len(pick))
    equirect = camera.get("model", "perspective") == "equirectangular"

    def back_project(uu, vv, depth):
        if equirect:                                                 # camera/equirectangular.cc:42-49: pixel -> bearing
            lon, lat = (uu / camera["cols"] - 0.5) * 2 * np.pi, -(vv / camera["rows"] - 0.5) * np.pi
            b = np.stack([np.cos(lat) * np.sin(lon), -np.sin(lat), np.cos(lat) * np.cos(lon)], 1)
            return b * depth[:, None]
        return np.stack([(uu - cx) / fx * depth, (vv - cy) / fy * depth, depth], 1)

    pc = back_project(u, v, z)
    n_cl = int(clutter_frac * len(pick))
    zc = rng.uniform(4, 60, n_cl) if equirect else rng.uniform(-10, 60, n_cl)   # (perspective: some behind the camera)
    uc, vc = rng.uniform(-200, camera["cols"] + 200, n_cl), rng.uniform(-100, camera["rows"] + 100, n_cl)
    if equirect:
        uc, vc = np.clip(uc, 0, camera["cols"] - 1), np.clip(vc, 0, camera["rows"] - 1)
    pcc = back_project(uc, vc, zc)
    pc_all = np.concatenate([pc, pcc])
    pw = (pc_all - tcw) @ Rcw
    n_lm = len(pw)
    ldesc = rng.integers(0, 256, (n_lm, 32), dtype=np.uint8)
    for j, k in enumerate(pick):
        row = desc[k].copy()
        for b in rng.choice(256, int(rng.integers(0, max_flips + 1)), replace=False):
            row[b >> 3] ^= np.uint8(1 << (b & 7))
        ldesc[j] = row
    octave = np.concatenate([kps["octave"][pick].astype(np.int64), rng.integers(0, len(sf), n_cl)])
    dist = np.linalg.norm(pw - center, axis=1)
    max_valid = (dist * sf[octave]).astype(np.float32)              # landmark::update_mean_normal_and_obs_scale_variance (landmark.cc:256-311)
    min_valid = (max_valid / np.float32(sf[-1])).astype(np.float32)
    normal = (pw - center) / np.maximum(dist, 1e-9)[:, None] + 0.2 * rng.standard_normal((n_lm, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    flip = rng.random(n_lm) < 0.03                                   # seen from behind: fails the viewing-angle test
    normal[flip] *= -1
    kp_of = np.concatenate([pick, np.full(n_cl, -1)])
    perm = rng.permutation(n_lm)                                     # the reference's local_landmarks_ order is arbitrary
    pw, ldesc, max_valid, min_valid, normal, kp_of = pw[perm], ldesc[perm], max_valid[perm], min_valid[perm], normal[perm], kp_of[perm]
    has_obs = (rng.random(n_lm) > 0.04).astype(np.uint8)
    skip = np.zeros(n_lm, np.uint8)
    kp_landmark = np.full(n_kp, -1, np.int32)
    for l in np.nonzero((kp_of >= 0) & (rng.random(n_lm) < pre_matched_frac))[0]:   # carried over from the motion-model step
        kp_landmark[kp_of[l]] = l
        skip[l] = 1
    skip[rng.random(n_lm) < 0.02] = 1                                # will_be_erased / temporal-ratio skips
    kp_x_right = None
    if stereo:
        kp_x_right = np.full(n_kp, -1.0, np.float32)
        zk = np.full(n_kp, np.nan)
        zk[pick] = z
        ok = ~np.isnan(zk) & (rng.random(n_kp) < 0.8)
        kp_x_right[ok] = (kps["x"][ok] - camera["fxb"] / zk[ok] + 0.5 * rng.standard_normal(ok.sum())).astype(np.float32)
    dR = _rodrigues(np.deg2rad(rot_deg) * rng.standard_normal(3) / np.sqrt(3))
    pose = np.eye(4)
    pose[:3, :3] = dR @ Rcw
    pose[:3, 3] = dR @ tcw + trans_m * rng.standard_normal(3) / np.sqrt(3)
    gt = np.eye(4)
    gt[:3, :3], gt[:3, 3] = Rcw, tcw
    return dict(pose_cw=pose, gt_pose_cw=gt, kp_landmark=kp_landmark, kp_x_right=kp_x_right,
                landmarks=dict(pos_w=pw, mean_normal=normal, min_valid_dist=min_valid, max_valid_dist=max_valid, desc=ldesc, skip=skip,
                               has_observation=has_obs))
