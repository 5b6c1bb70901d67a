/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the stella_vslam hot path (ORB extract -> Hamming match -> local BA).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Third-party arithmetic that is NOT in the reference tree (OpenCV 4.x cv::resize / cv::FAST /
 * cv::GaussianBlur / cv::fastAtan2, g2o 20230223_git LM + Schur) is restated from the published
 * algorithms; the OpenCV pieces are pinned bit-for-bit against cv2 4.13.0 by tests/golden (see
 * tests/golden/make_golden.py), the g2o piece is "parity unpinned" (see oracle/lba_oracle.c header).
 */
#ifndef B200VSLAM_ORACLE_H
#define B200VSLAM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint fields the reference fills (class_id stays -1). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave;
} orc_keypoint_t;

/* feature::orb_params (orb_params.cc:12-27) + orb_extractor ctor's min_area (orb_extractor.cc:16-20). */
typedef struct {
    float scale_factor;    /* 1.2 */
    int32_t num_levels;    /* 8 */
    int32_t ini_fast_thr;  /* 20 */
    int32_t min_fast_thr;  /* 7 */
    uint32_t min_area;     /* Preprocessing.min_size, system.cc:95 (default 800) */
} orc_orb_config_t;

/* ---- ORB primitives -------------------------------------------------------------------------- */
void orc_scale_factors(float scale_factor, int n, float* sf, float* inv_sf, float* sigma_sq, float* inv_sigma_sq);
void orc_level_size(int w, int h, float sf, int* lw, int* lh);
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int orc_fast9_16_nms(const uint8_t* img, int stride, int w, int h, int thr, int16_t* xs, int16_t* ys, uint8_t* scores, int cap);
void orc_gaussian7_s2_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
float orc_fast_atan2(float y, float x);
float orc_ic_angle(const uint8_t* img, int stride, int x, int y);
float orc_util_cos(float v);
float orc_util_sin(float v);
void orc_rbrief(const uint8_t* blurred, int stride, float x, float y, float angle_deg, uint8_t* desc32);
void orc_rect_mask(int cols, int rows, const float* rects, int n_rects, uint8_t* mask, int mstride);

/* feature::orb_extractor::extract (orb_extractor.cc:28-136).  Returns N (>=0), or -1 if cap is too small.
 * level_counts / raw_counts: optional [num_levels] outputs (kept keypoints, raw FAST candidates).
 * pyramid_out: optional array of num_levels caller buffers (tight stride = level width) receiving image_pyramid_. */
int orc_orb_extract(const uint8_t* img, int w, int h, int stride, const uint8_t* mask, int mask_stride,
                    const orc_orb_config_t* cfg, orc_keypoint_t* kps, uint8_t* descs, int cap, int* level_counts,
                    int* raw_counts, uint8_t** pyramid_out);

/* ---- matchers --------------------------------------------------------------------------------- */
unsigned orc_hamming_32(const uint8_t* a, const uint8_t* b);
unsigned orc_hamming_64(const uint8_t* a, const uint8_t* b);
float orc_angle_diff(float a1, float a2);
/* match::robust::brute_force_match (match/robust.cc:232-328).  valid2[i]!=0 <=> keyframe keypoint i has a live
 * landmark.  pairs_out: (idx_1, idx_2) sorted by idx_1.  Returns the number of matches. */
int orc_brute_force_match(const uint8_t* desc1, const float* angle1, int n1, const uint8_t* desc2, const float* angle2,
                          const uint8_t* valid2, int n2, float lowe_ratio, int check_orientation, int32_t* pairs_out);

/* timed CPU baseline driver (batch_oracle.c): n frames on n_threads pthreads, extract then match to predecessor */
int orc_frontend_batch(const uint8_t* frames, int n_unique, int n, int w, int h, const orc_orb_config_t* cfg, int cap, float lowe,
                       int check_ori, int n_threads, int* counts, int* n_matches);

#ifdef __cplusplus
}
#endif
#endif
