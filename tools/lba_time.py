"""Local-BA timing: one window and batches of windows per launch sequence (b200_lba_solve_batch), with the per-kernel device times
of the profiling mode (events after every launch).
usage: python tools/lba_time.py [model] [reps] [batch sizes ...]"""
import ctypes as C
import sys, time
sys.path.insert(0, ".")
import numpy as np
from stella_vslam_b200 import optimize
from workloads import synth
model = sys.argv[1] if len(sys.argv) > 1 else "stereo"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
batches = [int(x) for x in sys.argv[3:]] or [1, 4, 16, 32]
pr = synth.make_ba_problem(50, 10, 10000, seed=0, model=model)
ba = optimize.local_bundle_adjuster()
L = ba._L
L.b200_lba_enable_profile.argtypes = [C.c_void_p, C.c_int]
L.b200_lba_kernel_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
NAMES = ["plan", "landmark_build", "pose_rows", "schur", "cholesky", "backsub+trial", "(unused)", "tail"]
for nb in batches:
    prep = ba.prepare_batch([pr] * nb)
    for i in range(reps):
        L.b200_lba_enable_profile(ba._h, 1 if i == reps - 1 else 0)
        t = time.time(); launches = ba.optimize_prepared_batch(prep); dt = time.time() - t
        ms = C.c_float(); L.b200_lba_last_profile(ba._h, C.byref(ms), None)
        st = prep["st"][0]
        print("lba %s batch %d E=%d wall %.2f ms (%.3f ms/window) gpu %.2f ms (%.3f ms/window) launches %d iters %s outliers %d" % (
            model, nb, len(pr["e_pose"]), dt * 1e3, dt * 1e3 / nb, ms.value, ms.value / nb, launches, list(st.iterations), st.n_outliers), flush=True)
    parts = []
    for k, nm in enumerate(NAMES):
        v, n = C.c_float(), C.c_int()
        L.b200_lba_kernel_ms(ba._h, k, C.byref(v), C.byref(n))
        parts.append("%s %.0f us x%d" % (nm, 1e3 * v.value / max(n.value, 1), n.value))
    print("   per launch (profiling mode): " + ", ".join(parts), flush=True)
