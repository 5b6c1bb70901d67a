import sys, time
sys.path.insert(0, ".")
import numpy as np
from stella_vslam_b200 import optimize, synth
model = sys.argv[1] if len(sys.argv) > 1 else "stereo"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pr = synth.make_ba_problem(50, 10, 10000, seed=0, model=model)
ba = optimize.local_bundle_adjuster()
for i in range(reps):
    t = time.time(); r = ba.optimize(pr); dt = time.time() - t
    print("lba %s E=%d wall %.2f ms gpu %.2f ms launches %d iters %s outliers %d" % (model, len(pr["e_pose"]), dt * 1e3, r["gpu_ms"], r["launches"], r["iterations"], r["n_outliers"]))
