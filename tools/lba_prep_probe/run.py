import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np
from workloads import synth
from stella_vslam_b200.optimize import pack_problem
os.environ["B200_LBA_DEBUG"] = "1"
L = C.CDLL(os.path.join(os.path.dirname(__file__), "libprobe.so"))
pr = synth.make_ba_problem(50, 10, 10000, seed=0, model="stereo")
P, keep = pack_problem(pr)
pose, pts, outl = np.zeros((P.n_poses, 4, 4)), np.zeros((P.n_points, 3)), np.zeros(P.n_edges, np.uint8)
for _ in range(4):
    L.probe_plan(C.byref(P), pose.ctypes.data, pts.ctypes.data, outl.ctypes.data)
