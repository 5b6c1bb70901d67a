"""Wall-clock of the host-buffer matcher entry points (H2D + kernels + D2H), batched.  Usage: python tools/match_time.py [batch]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from stella_vslam_b200 import match  # noqa: E402
from workloads import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def bench(name, fn, reps=5):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    print(f"{name:42s} {1e3 * min(t):8.3f} ms / {B} problems  = {1e6 * min(t) / B:7.1f} us each")


for mode, thr in ((0, 100), (1, 100), (2, 100), (3, 50), (4, 50)):
    probs = [synth.make_guided_problem(k, n_train=2000, n_queries=2000, mode=mode) for k in range(8)] * (B // 8)
    for p in probs:
        p.pop("t_occupied", None)
    bench(f"guided mode {mode} 2000 x 2000", lambda: match.match_guided_batch(probs, mode, thr, 0.8, True))
for variant in (0, 1):
    ps = []
    for k in range(8):
        k1, k2, g = synth.make_keyframe_pair(k)
        if variant == 0:
            ps.append(dict(desc1=k1["desc"], angle1=k1["angle"], valid1=k1["has_landmark"], node1=k1["node"], desc2=k2["desc"], angle2=k2["angle"],
                           node2=k2["node"]))
        else:
            ps.append(match._triangulation_problem(k1, k2, g["E_12"], g["epiplane_in_keyfrm_2"], True, 0.2 * np.pi / 180, False))
    ps = ps * (B // 8)
    bench(f"pairs variant {variant} 2000 x 2000", lambda: match.match_pairs_batch(ps, variant, 0.75, True))
