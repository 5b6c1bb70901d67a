"""Wall-clock and device time of batched pose optimisation (one CUDA launch for the batch).  Usage: python tools/pose_opt_time.py [batch]"""
import sys, time
sys.path.insert(0, ".")
import ctypes as C
import numpy as np
from stella_vslam_b200 import optimize
from workloads import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
probs = [synth.make_pose_problem(k, n_obs=1500, model="stereo") for k in range(8)] * (B // 8)
po = optimize.pose_optimizer()
po.optimize_batch(probs)
t = time.perf_counter(); po.optimize_batch(probs); dt = time.perf_counter() - t
ms = C.c_float(); po._L.b200_lba_last_profile(po._h, C.byref(ms), None)
print(f"GPU: {B} frames x 1500 obs: wall {dt * 1e3:.2f} ms (incl. Python packing), kernel {ms.value:.3f} ms = {ms.value * 1e3 / B:.1f} us / frame")
