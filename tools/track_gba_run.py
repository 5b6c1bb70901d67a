"""One tracking-chain call (32 frames of 1241x376, one local map each) and one global bundle adjustment beyond the on-chip Cholesky
(260 keyframes) -- the workloads of profiles/r2_launches_track_gba.csv:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_track_gba.csv python tools/track_gba_run.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from stella_vslam_b200 import feature, optimize, tracking
from workloads import synth

B = 32
imgs = np.stack([synth.make_frame(1241, 376, seed=50 + i) for i in range(B)])
ex = feature.orb_extractor(feature.orb_params(), 800, max_batch=B)
kps, descs = ex.extract_batch(imgs)
cam = dict(model="perspective", fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, fxb=386.1448, cols=1241.0, rows=376.0, setup="stereo")
frames = [dict(synth.make_tracking_frame(kps[i], descs[i], cam, ex.orb_params_.scale_factors_, seed=70 + i, stereo=True), frame=i) for i in range(B)]
tr = tracking.local_map_tracker(ex, cam)
for _ in range(2):
    res = tr.track(frames)
print("tracking chain:", B, "frames,", sum(r["n_matches"] for r in res) // B, "matches / frame,", tr.stage_ms())
pr = synth.make_ba_problem(260, 1, 6000, seed=12, model="stereo")
gba = optimize.global_bundle_adjuster(3)
out = gba.optimize(pr)
print("global BA: 259 free keyframes, %d iterations, %d launches, %.2f ms on the stream" % (out["iterations"], out["launches"], out["gpu_ms"]))
