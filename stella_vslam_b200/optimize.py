"""Host-side mirror of optimize::local_bundle_adjuster (src/stella_vslam/optimize/local_bundle_adjuster.h:15-24,
local_bundle_adjuster_g2o.{h,cc}, local_bundle_adjuster_factory.h:15-33).

The reference's optimize(map_db, curr_keyfrm, force_stop_flag) first flattens the covisibility window into vertices and
edges (steps 1-4) and writes the result back under the map mutex (step 8); both stay on the host.  This module takes the
flattened problem (the layout of b200_lba_problem_t) and runs steps 5-7 on the GPU.
"""
import ctypes as C

import numpy as np

from ._lib import ERR_ABORTED, check, lib, ptr


class Camera(C.Structure):
    _fields_ = [("model", C.c_int32), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("fxb", C.c_double), ("cols", C.c_double), ("rows", C.c_double)]


class LbaProblem(C.Structure):
    _fields_ = [("n_poses", C.c_int32), ("n_points", C.c_int32), ("n_edges", C.c_int32), ("n_cams", C.c_int32),
                ("pose_cw", C.c_void_p), ("pose_fixed", C.c_void_p), ("points", C.c_void_p), ("point_fixed", C.c_void_p),
                ("e_pose", C.c_void_p), ("e_point", C.c_void_p), ("e_cam", C.c_void_p), ("e_obs", C.c_void_p),
                ("e_inv_sigma_sq", C.c_void_p), ("e_delta", C.c_void_p), ("e_robust", C.c_void_p), ("e_can_be_outlier", C.c_void_p),
                ("cams", C.c_void_p)]


class LbaStats(C.Structure):
    _fields_ = [("iterations", C.c_int32 * 2), ("n_outliers", C.c_int32), ("chi2", C.c_double * 2), ("lambda_init", C.c_double),
                ("lambda_final", C.c_double * 2)]


def _bind():
    L = lib()
    vp = C.c_void_p
    L.b200_lba_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.b200_lba_destroy.argtypes = [vp]
    L.b200_lba_solve.argtypes = [vp, C.POINTER(LbaProblem), C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(LbaStats)]
    L.b200_lba_solve_batch.argtypes = [vp, C.c_int, C.POINTER(LbaProblem), C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(LbaStats), vp]
    L.b200_lba_last_profile.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.b200_pose_optimize.argtypes = [vp, C.c_int, C.POINTER(LbaProblem), C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.b200_global_ba_solve.argtypes = [vp, C.POINTER(LbaProblem), C.c_int, C.c_double, vp, vp, vp, C.POINTER(LbaStats)]
    return L


def pack_problem(prob):
    """dict (see synth.make_ba_problem) -> (LbaProblem, keep-alive list)."""
    keep = []

    def arr(x, dt):
        if x is None:
            return None
        a = np.ascontiguousarray(x, dt)
        keep.append(a)
        return a.ctypes.data

    cams = (Camera * len(prob["cams"]))(*[Camera(c["model"], c["fx"], c["fy"], c["cx"], c["cy"], c["fxb"], c["cols"], c["rows"])
                                          for c in prob["cams"]])
    keep.append(cams)
    P = LbaProblem(len(prob["pose_cw"]), len(prob["points"]), len(prob["e_pose"]), len(prob["cams"]), arr(prob["pose_cw"], np.float64),
                   arr(prob["pose_fixed"], np.uint8), arr(prob["points"], np.float64), arr(prob.get("point_fixed"), np.uint8),
                   arr(prob["e_pose"], np.int32), arr(prob["e_point"], np.int32), arr(prob["e_cam"], np.uint8),
                   arr(prob["e_obs"], np.float32), arr(prob["e_inv_sigma_sq"], np.float32), arr(prob["e_delta"], np.float32),
                   arr(prob.get("e_robust"), np.uint8), arr(prob.get("e_can_be_outlier"), np.uint8), C.cast(cams, C.c_void_p))
    return P, keep


class local_bundle_adjuster:
    """optimize::local_bundle_adjuster with the "b200" backend (factory key Mapping.backend, local_bundle_adjuster_factory.h:17-32)."""

    def __init__(self, num_first_iter=5, num_second_iter=10, device=0):
        self.num_first_iter_ = int(num_first_iter)    # local_bundle_adjuster_g2o.h:25-27
        self.num_second_iter_ = int(num_second_iter)
        self._L = _bind()
        self._h = C.c_void_p()
        check(self._L.b200_lba_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_lba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare(self, problem):
        """Pack a problem once (ctypes struct + output buffers) for repeated, low-overhead optimize_prepared() calls."""
        P, keep = pack_problem(problem)
        K, L, E = P.n_poses, P.n_points, P.n_edges
        return dict(P=P, keep=keep, pose=np.zeros((K, 4, 4)), pts=np.zeros((L, 3)), outl=np.zeros(E, np.uint8), st=LbaStats())

    def optimize_prepared(self, prep, force_stop_flag=None):
        rc = self._L.b200_lba_solve(self._h, C.byref(prep["P"]), self.num_first_iter_, self.num_second_iter_, ptr(force_stop_flag),
                                    ptr(prep["pose"]), ptr(prep["pts"]), ptr(prep["outl"]), C.byref(prep["st"]))
        if rc == ERR_ABORTED:
            return None
        check(rc)
        launches = C.c_int()
        self._L.b200_lba_last_profile(self._h, None, C.byref(launches))
        return launches.value

    def prepare_batch(self, problems):
        """Pack several windows once (ctypes structs, pointer tables, output buffers) for repeated optimize_prepared_batch() calls."""
        packed = [pack_problem(pr) for pr in problems]
        n = len(packed)
        arr = (LbaProblem * n)(*[pk[0] for pk in packed])
        pose = [np.zeros((pk[0].n_poses, 4, 4)) for pk in packed]
        pts = [np.zeros((pk[0].n_points, 3)) for pk in packed]
        outl = [np.zeros(max(pk[0].n_edges, 1), np.uint8) for pk in packed]
        tab = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        return dict(n=n, arr=arr, keep=packed, pose=pose, pts=pts, outl=outl, pose_tab=tab(pose), pts_tab=tab(pts), outl_tab=tab(outl),
                    st=(LbaStats * n)(), status=np.zeros(n, np.int32))

    def optimize_prepared_batch(self, prep, force_stop_flags=None):
        """b200_lba_solve_batch on a prepared batch.  force_stop_flags: optional list of 1-element uint8 arrays / None per window.
        Returns the number of kernel launches the whole batch took; per-window status in prep["status"]."""
        flags = None
        if force_stop_flags is not None:
            flags = (C.c_void_p * prep["n"])(*[(f.ctypes.data if f is not None else None) for f in force_stop_flags])
        rc = self._L.b200_lba_solve_batch(self._h, prep["n"], prep["arr"], self.num_first_iter_, self.num_second_iter_, flags, prep["pose_tab"],
                                          prep["pts_tab"], prep["outl_tab"], prep["st"], ptr(prep["status"]))
        check(rc)
        launches = C.c_int()
        self._L.b200_lba_last_profile(self._h, None, C.byref(launches))
        return launches.value

    def optimize_batch(self, problems, force_stop_flags=None):
        """Several independent windows in one launch sequence (b200_lba_solve_batch).  Returns one result per window: a dict like
        optimize(), or None for a window whose flag was already set (local_bundle_adjuster_g2o.cc:308-310)."""
        if not problems:
            return []
        prep = self.prepare_batch(problems)
        launches = self.optimize_prepared_batch(prep, force_stop_flags)
        ms = C.c_float()
        self._L.b200_lba_last_profile(self._h, C.byref(ms), None)
        out = []
        for w in range(prep["n"]):
            if prep["status"][w] == ERR_ABORTED:
                out.append(None)
                continue
            st = prep["st"][w]
            E = prep["arr"][w].n_edges
            out.append(dict(pose_cw=prep["pose"][w], points=prep["pts"][w], outliers=prep["outl"][w][:E], iterations=list(st.iterations),
                            n_outliers=st.n_outliers, chi2=list(st.chi2), lambda_init=st.lambda_init, lambda_final=list(st.lambda_final),
                            gpu_ms=ms.value, launches=launches))
        return out

    def optimize(self, problem, force_stop_flag=None):
        """problem: flattened window (dict).  force_stop_flag: optional 1-element uint8 array (read AND written, like the
        reference's bool*).  Returns None if the flag was already set (local_bundle_adjuster_g2o.cc:308-310), else a dict."""
        P, keep = pack_problem(problem)
        K, L, E = P.n_poses, P.n_points, P.n_edges
        pose_out, pts_out, outl = np.zeros((K, 4, 4)), np.zeros((L, 3)), np.zeros(E, np.uint8)
        st = LbaStats()
        rc = self._L.b200_lba_solve(self._h, C.byref(P), self.num_first_iter_, self.num_second_iter_, ptr(force_stop_flag), ptr(pose_out),
                                    ptr(pts_out), ptr(outl), C.byref(st))
        if rc == ERR_ABORTED:
            return None
        check(rc)
        ms, launches = C.c_float(), C.c_int()
        self._L.b200_lba_last_profile(self._h, C.byref(ms), C.byref(launches))
        return dict(pose_cw=pose_out, points=pts_out, outliers=outl, iterations=list(st.iterations), n_outliers=st.n_outliers,
                    chi2=list(st.chi2), lambda_init=st.lambda_init, lambda_final=list(st.lambda_final), gpu_ms=ms.value,
                    launches=launches.value)


class global_bundle_adjuster:
    """optimize::global_bundle_adjuster (optimize/global_bundle_adjuster.h:18-62): one LM round over the whole map; Huber is the
    problem's e_robust array (use_huber_kernel_)."""

    def __init__(self, num_iter=10, use_huber_kernel=True, verbose=False, device=0):
        self.num_iter_, self.use_huber_kernel_, self.verbose_ = int(num_iter), bool(use_huber_kernel), bool(verbose)
        self._L = _bind()
        self._h = C.c_void_p()
        check(self._L.b200_lba_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_lba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def optimize(self, problem, force_stop_flag=None, gain_threshold=1e-3):
        """global_bundle_adjuster::optimize (:258-420; gain threshold 1e-3) / optimize_for_initialization (:201-256; the caller's
        gain_threshold).  Returns None when the caller's flag aborted the solve (the reference returns false), else a dict."""
        pr = dict(problem)
        if not self.use_huber_kernel_:
            pr["e_robust"] = np.zeros(len(pr["e_pose"]), np.uint8)
        P, keep = pack_problem(pr)
        pose_out, pts_out = np.zeros((P.n_poses, 4, 4)), np.zeros((P.n_points, 3))
        st = LbaStats()
        rc = self._L.b200_global_ba_solve(self._h, C.byref(P), self.num_iter_, float(gain_threshold), ptr(force_stop_flag), ptr(pose_out),
                                          ptr(pts_out), C.byref(st))
        if rc == ERR_ABORTED:
            return None
        check(rc)
        ms, launches = C.c_float(), C.c_int()
        self._L.b200_lba_last_profile(self._h, C.byref(ms), C.byref(launches))
        return dict(pose_cw=pose_out, points=pts_out, iterations=st.iterations[0], chi2=st.chi2[0], lambda_init=st.lambda_init,
                    lambda_final=st.lambda_final[0], gpu_ms=ms.value, launches=launches.value)


class pose_optimizer:
    """optimize::pose_optimizer (optimize/pose_optimizer.h:24-40, pose_optimizer_g2o.{h,cc}; factory defaults
    pose_optimizer_factory.h:18-47): motion-only BA of frames, one CUDA launch for a whole batch."""

    def __init__(self, num_trials_robust=2, num_trials=2, num_each_iter=10, device=0):
        self.num_trials_robust_, self.num_trials_, self.num_each_iter_ = int(num_trials_robust), int(num_trials), int(num_each_iter)
        self._L = _bind()
        self._h = C.c_void_p()
        check(self._L.b200_lba_create(device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200_lba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def optimize_batch(self, problems):
        """problems: flattened frames (synth.make_pose_problem layout: one pose, fixed landmarks, one edge per observation).
        Returns [(num_valid_obs, optimized_pose (4,4), outlier_flags (n,) bool)] like pose_optimizer::optimize (:38-44)."""
        if not problems:
            return []
        packed = [pack_problem(pr) for pr in problems]
        arr = (LbaProblem * len(problems))(*[pk[0] for pk in packed])
        n_edges = [pk[0].n_edges for pk in packed]
        pose = np.zeros((len(problems), 4, 4))
        flags = np.zeros(max(sum(n_edges), 1), np.uint8)
        valid = np.zeros(len(problems), np.uint32)
        check(self._L.b200_pose_optimize(self._h, len(problems), arr, self.num_trials_robust_, self.num_trials_, self.num_each_iter_,
                                         ptr(pose), ptr(flags), ptr(valid)))
        out, off = [], 0
        for i, ne in enumerate(n_edges):
            out.append((int(valid[i]), pose[i].copy(), flags[off:off + ne].astype(bool)))
            off += ne
        return out

    def optimize(self, problem):
        return self.optimize_batch([problem])[0]


def create(yaml_node=None, device=0):
    """local_bundle_adjuster_factory::create (local_bundle_adjuster_factory.h:17-32): Mapping.backend must be "b200" here;
    "g2o"/"gtsam" are the reference's CPU backends and are not part of this library."""
    node = yaml_node or {}
    backend = node.get("backend", "b200")
    if backend != "b200":
        raise RuntimeError(f"Invalid backend: {backend}")
    return local_bundle_adjuster(node.get("num_first_iter", 5), node.get("num_second_iter", 10), device)
