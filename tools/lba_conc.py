import sys, time
sys.path.insert(0, ".")
from concurrent.futures import ThreadPoolExecutor
from stella_vslam_b200 import optimize
from workloads import synth
pr = synth.make_ba_problem(50, 10, 10000, seed=0, model="stereo")
import os
print('host loop' if os.environ.get('B200_LBA_HOST_LOOP') == '1' else 'graph')
for nthr in (1, 2, 4, 8, 16):
    hs = [optimize.local_bundle_adjuster() for _ in range(nthr)]
    with ThreadPoolExecutor(nthr) as ex:
        list(ex.map(lambda h: h.optimize(pr), hs))  # warm
        t = time.time()
        reps = 4
        futs = [ex.submit(h.optimize, pr) for h in hs for _ in range(reps)]
        rs = [f.result() for f in futs]
        dt = time.time() - t
    print("threads %2d: %.2f ms per LBA (throughput), mean gpu_ms %.2f" % (nthr, dt * 1e3 / (nthr * reps), sum(r["gpu_ms"] for r in rs) / len(rs)))
    for h in hs:
        h.close()
