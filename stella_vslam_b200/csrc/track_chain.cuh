// track_chain.cuh -- internal interfaces of b200_track_local_map (include/b200vslam.h): the chain is driven from match_kernels.cu,
// stage A (undistort + can_observe + query build) lives in orb_kernels.cu (same device functions and -fmad=false as the stage-by-stage
// ABI), stage C (edge build + pose optimisation) in lba_kernels.cu.  Plain device pointers, no handles' internals cross a TU.
#pragma once

#include "common.cuh"

namespace b200 {
namespace chain {

struct TrackShared {  // by-value kernel parameter
    int model;
    double fx, fy, cx, cy, k1, k2, p1, p2, k3, cols, rows, fxb;
    float min_x, max_x, min_y, max_y;
    float ray_cos_thr, log_scale_factor, margin, delta;
    unsigned num_levels;
    float scale_factors[32], inv_level_sigma_sq[32];
};

struct TrackFrameDev {  // one frame; every pointer is a device pointer
    // the extractor's results for this frame
    const b200_keypoint_t* kps;
    const int* n_kp;
    int kp_cap;
    // caller inputs
    int n_kp_in, n_lm;
    const float* kp_x_right;   // may be null
    const int* kp_landmark;    // may be null
    const double *pos_w, *mean_normal;
    const float *min_d, *max_d;
    const unsigned char *lm_skip, *lm_has_obs;  // may be null
    double Rt[12], twc[3];
    // stage A
    b200_keypoint_t* undist;
    float *t_x, *t_y;
    unsigned char *t_octave, *occupied;
    unsigned char* observable;
    float *q_x, *q_y, *q_margin, *q_xr;
    signed char *q_lo, *q_hi;
    unsigned char* q_valid;
    // stage B (guided matcher)
    const int* match_out;
    // stage C
    int* kp_landmark_out;
    unsigned char* kp_outlier;
    int* status;  // [0] keypoint count, [1] != 0: n_kp_in disagrees with the extractor's count
};

// orb_kernels.cu: device views of the extractor's last batch + the stream its work is ordered on
int orb_results(b200_orb_t orb, const b200_keypoint_t** d_kps, const unsigned char** d_descs, const int** d_counts, int* stride, int* batch,
                cudaStream_t* stream, int* device);
int track_stage_a(cudaStream_t st, const TrackShared& sh, const TrackFrameDev* d_frames, int n_frames, int max_kp, int max_lm);
// lba_kernels.cu: builds one edge per keypoint that carries a landmark (keypoint order), runs pose_optimizer::optimize for every frame and
// scatters the flags back to keypoint indexing.  h_frames = the host copy of d_frames; pose_out / n_valid are device pointers.
int track_stage_c(b200_lba_t opt, cudaStream_t st, const TrackShared& sh, const TrackFrameDev* d_frames, const TrackFrameDev* h_frames,
                  const double* const* pose_cw, int n_frames, int max_kp, int trials_robust, int trials, int each_iter, double* d_pose_out,
                  unsigned* d_n_valid, cudaEvent_t ev_edges_done);

}  // namespace chain
}  // namespace b200
