/*
 * b200vslam.h -- C ABI of the B200-native stella_vslam hot path (ORB extract -> Hamming match -> local BA).
 *
 * The reference has no FFI of its own: its "operator API" for this path is three C++ class surfaces
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces (paths relative to the
 * reference checkout); the C++ adapters in stella_vslam_b200/host/ and INTEGRATION.md show the binding.
 *
 * Conventions: plain pointers and sizes only, caller-allocated outputs, int status (0 = OK, <0 = error, see
 * b200_last_error()), no exceptions cross the boundary.  One opaque handle per instance owns a CUDA stream and its
 * device arenas; handles are not thread-safe, distinct handles are independent (the reference runs the left/right
 * extractors in two threads, system.cc:427-434).  There is NO CPU fallback: every entry point fails with
 * B200_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef B200VSLAM_H
#define B200VSLAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID (-1)  /* bad argument (also: image type/size the reference would assert on) */
#define B200_ERR_CUDA (-2)     /* CUDA runtime/driver failure or no device */
#define B200_ERR_CAPACITY (-3) /* caller buffer too small; counts are still written */
#define B200_ERR_ABORTED (-4)  /* local BA: force-stop flag was set before the first solve (no write-back) */

const char* b200_last_error(void);
int b200_device_count(void);
/* bytes of the library's version string: "b200vslam <semver> sm_100a" */
const char* b200_version(void);

/* Pinned host memory helpers (the end-to-end path copies from/to pinned buffers). */
int b200_host_alloc(void** ptr, size_t bytes);
int b200_host_free(void* ptr);

/* ------------------------------------------------------------------------------------------------------------------
 * feature::orb_extractor  (src/stella_vslam/feature/orb_extractor.h:51-61, orb_extractor.cc:16-136)
 * ---------------------------------------------------------------------------------------------------------------- */

/* cv::KeyPoint as the reference fills it (orb_extractor.cc:273-283, 337-345); class_id is always -1. */
typedef struct {
    float x, y;      /* level-0 pixel coordinates (pt * scale_factor[octave]) */
    float size;      /* (unsigned)(31 * scale_factor[octave]) */
    float angle;     /* IC angle, degrees [0,360) */
    float response;  /* FAST score */
    int32_t octave;  /* pyramid level */
} b200_keypoint_t;

/* feature::orb_params (orb_params.cc:12-27) + orb_extractor ctor arguments (orb_extractor.cc:16-20). */
typedef struct {
    float scale_factor;       /* Feature.scale_factor, default 1.2 */
    int32_t num_levels;       /* Feature.num_levels, default 8 (max 16) */
    int32_t ini_fast_thr;     /* Feature.ini_fast_threshold, default 20 */
    int32_t min_fast_thr;     /* Feature.min_fast_threshold, default 7 */
    uint32_t min_area;        /* Preprocessing.min_size (system.cc:95), default 800 */
    int32_t n_mask_rects;     /* mask_rects ctor argument: n x {x_min, x_max, y_min, y_max} as image fractions */
    const float* mask_rects;  /* may be NULL when n_mask_rects == 0 */
    int32_t device;           /* CUDA device ordinal */
    int32_t max_batch;        /* frames per extract call this instance is sized for (>=1); grows on demand */
} b200_orb_params_t;

typedef struct b200_orb_s* b200_orb_t;

void b200_orb_default_params(b200_orb_params_t* p);
int b200_orb_create(const b200_orb_params_t* p, b200_orb_t* out);
int b200_orb_destroy(b200_orb_t h);

/* Upper bound of keypoints one w x h frame can yield (sum over levels of selection-grid cells). */
int b200_orb_max_keypoints(b200_orb_t h, int width, int height);

/* orb_extractor::extract (orb_extractor.cc:28-136) for `batch` same-sized CV_8UC1 frames held in HOST memory.
 *   images      : frame f starts at images + f*frame_stride, rows `pitch` bytes apart.
 *   mask        : optional CV_8UC1 image mask at level-0 resolution shared by the batch (0 = masked), or NULL; when
 *                 NULL the rectangle mask built from mask_rects is used if any (orb_extractor.cc:50-64).
 *   kps/descs   : frame f writes kps[f*cap ..], descs[(f*cap + i)*32 ..]; counts[f] = N_f.
 * width==0 || height==0 || batch==0 -> B200_OK with nothing written (orb_extractor.cc:30-32).
 * Includes the host->device copy of the frames and the device->host copy of the results. */
int b200_orb_extract(b200_orb_t h, const uint8_t* images, int width, int height, size_t pitch, size_t frame_stride,
                     int batch, const uint8_t* mask, size_t mask_pitch, b200_keypoint_t* kps, uint8_t* descs, int cap,
                     int32_t* counts);

/* Same, with frames already resident in device memory (and the mask, if any); results stay on the device until
 * b200_orb_fetch.  Enqueued on the instance's stream (see b200_orb_set_stream) without synchronising. */
int b200_orb_extract_device(b200_orb_t h, const void* d_images, int width, int height, size_t pitch, size_t frame_stride,
                            int batch, const void* d_mask, size_t mask_pitch);
/* Run on the caller's stream (a cudaStream_t, e.g. torch's current stream; NULL is the legacy default stream).
 * use_own != 0 ignores `stream` and restores the instance's own non-blocking stream. */
int b200_orb_set_stream(b200_orb_t h, void* stream, int use_own);
/* Write results into caller-owned DEVICE buffers (e.g. torch tensors) instead of the instance's arenas:
 * keypoints [batch][stride_kps], descriptors [batch][stride_kps][32], counts [batch]; stride_kps must be >=
 * b200_orb_max_keypoints().  d_kps == NULL unbinds. */
int b200_orb_bind_outputs(b200_orb_t h, void* d_kps, void* d_descs, void* d_counts, int stride_kps);
/* Size the arenas for `batch` w x h frames now (otherwise done lazily by the first extract). */
int b200_orb_reserve(b200_orb_t h, int width, int height, int batch);
/* Copy the last extract's results to host buffers (synchronises the instance stream). */
int b200_orb_fetch(b200_orb_t h, b200_keypoint_t* kps, uint8_t* descs, int cap, int32_t* counts);
/* Device views of the last extract's results: keypoints [batch][stride_kps], descriptors [batch][stride_kps][32],
 * counts [batch].  Valid until the next extract on this handle. */
int b200_orb_device_results(b200_orb_t h, const b200_keypoint_t** d_kps, const uint8_t** d_descs, const int32_t** d_counts,
                            int* stride_kps);
int b200_orb_sync(b200_orb_t h);

/* orb_extractor::image_pyramid_ (orb_extractor.h:71; read by match::stereo via system.cc:443): level geometry and a
 * device view / host copy of one level of one frame of the last extract. */
int b200_orb_level_info(b200_orb_t h, int level, int* width, int* height, size_t* pitch, float* scale_factor);
int b200_orb_pyramid_level_device(b200_orb_t h, int frame, int level, const uint8_t** d_ptr);
int b200_orb_pyramid_level_host(b200_orb_t h, int frame, int level, uint8_t* dst, size_t dst_pitch);
/* Every level including 0 (the caller's device image, or the upload staging of b200_orb_extract), with its pitch and size. */
int b200_orb_pyramid_level_view(b200_orb_t h, int frame, int level, const uint8_t** d_ptr, size_t* pitch, int* width, int* height);

/* The per-keypoint steps between extractor and matchers, for `n` keypoints in HOST buffers:
 *   camera::perspective::undistort_keypoints     (src/stella_vslam/camera/perspective.cc:245-275: cv::undistortPoints with
 *       TermCriteria(EPS | MAX_ITER, 20, 1e-6), R = I, P = K; also Perspective without distortion, e.g. KITTI)
 *   camera::equirectangular::undistort_keypoints (the identity)
 *   camera::base::convert_keypoints_to_bearings  (camera/base.cc:158-162; perspective.cc:117-122, equirectangular.cc:42-49)
 * model: 0 perspective, 1 equirectangular.  undist_keypts (n) and bearings (n x 3 doubles) may each be NULL.
 * Runs on the extractor's stream (after the extract whose keypoints it is given). */
typedef struct {
    int32_t model;
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3; /* cv_dist_params_ */
    double cols, rows;         /* equirectangular */
} b200_camera_intrinsics_t;
int b200_keypoints_undistort(b200_orb_t h, const b200_camera_intrinsics_t* cam, const b200_keypoint_t* keypts, int n,
                             b200_keypoint_t* undist_keypts, double* bearings);

/* data::frame::can_observe (src/stella_vslam/data/frame.cc:59-84) for the `n` local landmarks a frame may see
 * (tracking_module.cc:559-594): reprojection (camera/perspective.cc:130-148, equirectangular.cc:59-73), ORB scale range
 * (data/landmark.h:88-92), viewing angle, predicted pyramid level (data/landmark.cc:336-353).  pose_cw: 4x4 row-major;
 * img_bounds = {min_x, max_x, min_y, max_y} (camera::base::img_bounds_, perspective only); per landmark: pos_w (3 doubles),
 * mean_normal (3 doubles), min / max valid distance (floats).  Out per landmark: observable (0/1) and, when observable, the
 * reprojection (2 doubles), x_right and pred_scale_level -- the inputs of b200_match_guided mode 0. */
int b200_frame_can_observe(b200_orb_t h, const b200_camera_intrinsics_t* cam, double focal_x_baseline, const float* img_bounds,
                           const double* pose_cw, int n, const double* pos_w, const double* mean_normal, const float* min_valid_dist,
                           const float* max_valid_dist, float ray_cos_thr, unsigned num_levels, float log_scale_factor, uint8_t* observable,
                           double* reproj, float* x_right, uint32_t* pred_scale_level);

/* util::convert_to_grayscale (src/stella_vslam/util/image_converter.cc:8-39): cv::cvtColor(COLOR_{RGB,BGR}[A]2GRAY) of
 * 8-bit frames, the step in front of the extractor (system.cc:370-378).  channels: 3 or 4; rgb_order != 0 <=> color_order_t::RGB.
 * The host variant converts one frame; the device variant converts `batch` frames in place on the extractor's stream so that
 * b200_orb_extract_device can follow without a copy (source frames 16-byte aligned, pitches multiples of 4). */
int b200_convert_to_grayscale(b200_orb_t h, const uint8_t* src, int width, int height, size_t src_pitch, int channels, int rgb_order,
                              uint8_t* gray, size_t gray_pitch);
int b200_convert_to_grayscale_device(b200_orb_t h, const void* d_src, int width, int height, size_t src_pitch, size_t src_frame_stride,
                                     int channels, int rgb_order, void* d_gray, size_t gray_pitch, size_t gray_frame_stride, int batch);

/* Keyframe serialisation: the byte layouts in which data::keyframe stores what the extractor produced.
 *   SQLite (data/keyframe.cc:298-347 to_db, :191-235 from_stmt): `undist_keypts` = the std::vector<cv::KeyPoint> as raw bytes (28 bytes per
 *       keypoint: pt.x, pt.y, size, angle, response, octave, class_id), `descs` = cv::Mat rows, 32 bytes each;
 *   JSON / msgpack (data/common.cc:57-81): a descriptor = eight uint32 read through `desc.ptr<uint32_t>()` -- on a little-endian host the
 *       same 32 bytes, so `desc_blob` viewed as uint32[n][8] is convert_descriptors_to_json's payload.
 * Exports frame `frame` of the last extract straight from HBM: the keypoints are undistorted on the device when `cam` is given
 * (camera::*::undistort_keypoints -- the keyframe stores undist_keypts_), re-packed to cv::KeyPoint records (class_id = -1, response = 0
 * for perspective cameras as perspective.cc:266-272 leaves it) and copied out with the descriptors.  *n = keypoints of the frame;
 * B200_ERR_CAPACITY if cap is too small. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} b200_cv_keypoint_t;
int b200_orb_export_keyframe_blobs(b200_orb_t h, int frame, const b200_camera_intrinsics_t* cam, b200_cv_keypoint_t* keypts_blob,
                                   uint8_t* desc_blob, int cap, int32_t* n);
/* The inverse for a keyframe loaded from a map file (host-only byte shuffling, no GPU work): cv::KeyPoint records -> b200_keypoint_t. */
int b200_keyframe_blob_to_keypoints(const b200_cv_keypoint_t* keypts_blob, int n, b200_keypoint_t* keypts);

/* Raw FAST corners (after NMS, threshold choice and mask tests, before distribute_keypoints) of the first n frames of the last extract:
 * the candidate count of orb_extractor.cc:237-259, which prices the FAST and selection kernels.  Synchronises. */
int b200_orb_raw_corner_counts(b200_orb_t h, int32_t* counts, int n);
/* Per-stage kernel time of the last extract, in ms, measured with CUDA events on the instance stream.
 * stage: 0 pyramid, 1 FAST+NMS+grid arg-max, 2 ordered selection, 3 (unused: the descriptor blur is fused into stage 4),
 * 4 window blur + orientation + rBRIEF, 5 whole extract. */
int b200_orb_stage_ms(b200_orb_t h, int stage, float* ms);
int b200_orb_enable_timing(b200_orb_t h, int enable);

/* ------------------------------------------------------------------------------------------------------------------
 * match::compute_descriptor_distance_32 / match::robust::brute_force_match
 * (src/stella_vslam/match/base.h:15-41, match/robust.cc:232-328)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct b200_matcher_s* b200_matcher_t;

int b200_matcher_create(int device, b200_matcher_t* out);
int b200_matcher_destroy(b200_matcher_t h);

/* All-pairs 256-bit Hamming distances: dist[i*n2 + j] = popcount(desc1[i] ^ desc2[j])  (host buffers). */
int b200_hamming_matrix(b200_matcher_t h, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, uint16_t* dist);

/* robust::brute_force_match for `n_problems` independent (frame, keyframe) pairs.  Problem p reads
 *   frame side    (frm_obs):  descriptors desc1 + 32*(off1[p]+i), angle *(float*)((char*)angle1 + (off1[p]+i)*angle1_stride),
 *                             i < cnt1[p]   (angle1_stride = 4 for a float array, sizeof(b200_keypoint_t) for &kps[0].angle)
 *   keyframe side (keyfrm) :  likewise with off2/cnt2; valid2[off2[p]+i] != 0 <=> keypoint i has a live landmark
 *                             (robust.cc:255-262); valid2 == NULL means all valid.
 * and writes (idx_1, idx_2) int32 pairs sorted by idx_1 at pairs + 2*p*pairs_stride, and n_pairs[p].
 * pairs_stride >= max_p cnt1[p].  lowe_ratio / check_orientation: the matcher's ctor arguments (match/base.h:81-91).
 * Host buffers; includes the uploads and the download of the results. */
int b200_match_bruteforce(b200_matcher_t h, int n_problems, const uint8_t* desc1, const void* angle1, size_t angle1_stride,
                          const int32_t* off1, const int32_t* cnt1, const uint8_t* desc2, const void* angle2, size_t angle2_stride,
                          const uint8_t* valid2, const int32_t* off2, const int32_t* cnt2, float lowe_ratio, int check_orientation,
                          int32_t* pairs, int pairs_stride, int32_t* n_pairs);
/* Device-resident variant: every pointer is a device pointer and the work is enqueued on the matcher's stream
 * (see b200_matcher_set_stream) without synchronising.  Problem p reads
 *   frame side    : descriptors d_desc1 + 32*(off1[p]+i), angle *(float*)((char*)d_angle1 + (off1[p]+i)*angle1_stride), i < cnt1[p]
 *   keyframe side : likewise with off2/cnt2 (cnt arrays live on the device, e.g. the extractor's d_counts)
 * and writes its pairs at d_pairs + 2*p*pairs_stride and its count at d_n_pairs[p].  max_n1/max_n2 bound cnt1/cnt2 (host-side
 * upper bounds used to size the launch); pairs_stride >= max_n1. */
int b200_match_bruteforce_device(b200_matcher_t h, int n_problems, const void* d_desc1, const void* d_angle1, size_t angle1_stride,
                                 const void* d_off1, const void* d_cnt1, const void* d_desc2, const void* d_angle2,
                                 size_t angle2_stride, const void* d_valid2, const void* d_off2, const void* d_cnt2, int max_n1,
                                 int max_n2, float lowe_ratio, int check_orientation, void* d_pairs, int pairs_stride,
                                 void* d_n_pairs);
/* Optional: run the sequential resolve pass of b200_match_bruteforce_device on a side stream.  The pass is the reference's greedy loop
 * (robust.cc:253-315) -- one warp per problem, ~0.5 ms for 64 pairs of 2000 keypoints with 147 of 148 SMs idle -- so when it is enabled
 * the caller's NEXT kernels on the matcher's stream (e.g. the next batch's extraction) overlap it.  Contract while enabled: d_pairs /
 * d_n_pairs of a call, and the freedom to overwrite that call's input buffers, are reached on the matcher's stream only after the next
 * b200_match_bruteforce_device call on this handle, b200_matcher_join (stream-ordered: the stream waits, the host does not) or
 * b200_matcher_sync.  Off by default. */
int b200_matcher_set_async_resolve(b200_matcher_t h, int enable);
int b200_matcher_join(b200_matcher_t h);
/* Grid-guided projection matchers
 *   mode 0 (B200_GUIDED_LANDMARKS): match::projection::match_frame_and_landmarks     (src/stella_vslam/match/projection.cc:13-93)
 *   mode 1 (B200_GUIDED_LAST_FRAME): match::projection::match_current_and_last_frames (src/stella_vslam/match/projection.cc:95-207)
 * and, on the same search primitive, the occasional (relocalisation / loop closure / mapping / initialisation) variants:
 *   mode 1 with thr = hamm_dist_thr, t_x_right = NULL:        projection::match_frame_and_keyframe  (projection.cc:217-319)
 *   mode 1 with thr = 50, check_orientation = 0:              projection::match_by_Sim3_transform   (projection.cc:321-416)
 *   mode 2 (B200_GUIDED_INDEPENDENT) once per direction, then b200_match_cross_check:
 *                                                             projection::match_keyframes_mutually  (projection.cc:418-630)
 *   mode 3 (B200_GUIDED_FUSE) with thr = 50:                  fuse::detect_duplication              (match/fuse.cc:12-154)
 *   mode 4 (B200_GUIDED_AREA) with thr = 50, lowe_ratio:      area::match_in_consistent_area        (match/area.cc:8-98); a later
 *        query may take a keypoint over from an earlier one (its match_out entry goes back to -1, area.cc:75-81)
 * including the keypoint grid they search: data::assign_keypoints_to_grid / get_keypoints_in_cell
 * (src/stella_vslam/data/common.cc:83-190).  The caller (the adapter) does what needs the map: it walks the landmarks in the
 * reference's order, reprojects them (camera::base::reproject_to_image), predicts the pyramid level and fills one query per
 * landmark; q_valid[q] == 0 marks a landmark the reference skips before the search (will_be_erased, !is_observable_in_tracking,
 * reprojection failed / outside the image, last-frame outlier).  All pointers are HOST buffers.
 * Result: match_out[q] = index of the frame keypoint landmark q is attached to (frm.add_landmark(lm, idx)) or -1, n_matches;
 * t_occupied is updated in place in modes 0, 1 and 3 (a keypoint that received a landmark is not offered to later landmarks,
 * projection.cc:50-53, 163-166, 292-294, 388-390; fuse.cc:88-90); mode 2 reads it only, mode 4 ignores it.  `n_problems` independent frames are processed in one launch sequence. */
enum { B200_GUIDED_LANDMARKS = 0, B200_GUIDED_LAST_FRAME = 1, B200_GUIDED_INDEPENDENT = 2, B200_GUIDED_FUSE = 3, B200_GUIDED_AREA = 4 };
typedef struct b200_guided_problem {
    int32_t n_train;                  /* keypoints of the frame that is searched */
    const float* t_x;                 /* frm_obs_.undist_keypts_[i].pt.x */
    const float* t_y;
    const uint8_t* t_octave;
    const float* t_angle;             /* needed when mode 1 checks orientation, else may be NULL */
    const float* t_x_right;           /* frm_obs_.stereo_x_right_, NULL when empty (monocular) */
    const uint8_t* t_desc;            /* n_train x 32 */
    uint8_t* t_occupied;              /* in/out, n_train: keypoint already carries a landmark with observations; NULL = none (no write-back) */
    float min_x, max_x, min_y, max_y; /* camera::base::img_bounds_ */
    int32_t grid_cols, grid_rows;     /* camera::base::num_grid_cols_ / num_grid_rows_ (64 x 48) */
    int32_t n_queries;                /* landmarks in the reference's iteration order */
    const uint8_t* q_desc;            /* n_queries x 32 (landmark::get_descriptor) */
    const float* q_x;                 /* reprojection */
    const float* q_y;
    const float* q_margin;            /* margin * scale_factors_[level], evaluated in float like the reference */
    const int8_t* q_min_level;        /* octave window; < 0 = unchecked (data/common.cc:160-173) */
    const int8_t* q_max_level;
    const float* q_x_right;           /* reprojected x_right; read only when t_x_right != NULL */
    const float* q_angle;             /* last-frame keypoint angle; read only in mode 1 with check_orientation */
    const uint8_t* q_valid;           /* NULL = all valid */
    const uint8_t* q_has_observation; /* modes 0 / 1: landmark::has_observation() of the query landmark; NULL = all have.  A keypoint that
                                         receives a landmark WITHOUT observations (a temporal landmark of a stereo / RGBD last frame) stays
                                         open to later landmarks, exactly like the reference's test `lm && lm->has_observation()`
                                         (projection.cc:50-53, 163-166); match_out then lists the keypoint for both and the adapter's
                                         in-order add_landmark leaves the later one, as in the reference */
    const double* q_reproj;           /* mode 3 with do_reprojection_matching: n_queries x 2, the reprojection in double (fuse.cc:96-97) */
    const float* inv_level_sigma_sq;  /* mode 3: orb_params_->inv_level_sigma_sq_, n_levels entries */
    int32_t n_levels;
    int32_t do_reprojection_matching; /* mode 3 (fuse.cc:19, :93) */
    int32_t* match_out;               /* out, n_queries */
    int32_t n_matches;                /* out */
} b200_guided_problem_t;
/* thr: the reference's `best <= thr` acceptance (100 for modes 0-2 in the callers cited, 50 for 3-4); lowe_ratio: modes 0 and 4; max_candidates bounds the
 * keypoints one search window may return (0 = default 256); B200_ERR_CAPACITY reports the size that would have been needed. */
int b200_match_guided(b200_matcher_t h, int n_problems, b200_guided_problem_t* problems, int mode, unsigned thr, float lowe_ratio,
                      int check_orientation, int max_candidates);
/* Closing loop of match_keyframes_mutually (projection.cc:614-627) on two mode-2 results: mutual_out[i] = j iff idx2_in_1[i] == j
 * and idx1_in_2[j] == i, else -1.  Host-side, no device work. */
int b200_match_cross_check(const int32_t* idx2_in_1, int n1, const int32_t* idx1_in_2, int n2, int32_t* mutual_out, int32_t* n_mutual);
/* All-pairs matchers with greedy state, other than brute_force_match:
 *   variant 0 (B200_PAIRS_BOW):           match::bow_tree::match_frame_and_keyframe  (src/stella_vslam/match/bow_tree.cc:169-256)
 *                                         match::bow_tree::match_keyframes           (bow_tree.cc:258-366)
 *   variant 1 (B200_PAIRS_TRIANGULATION): match::robust::match_for_triangulation     (src/stella_vslam/match/robust.cc:14-146)
 *                                         match::bow_tree::match_for_triangulation   (bow_tree.cc:11-167)
 *                                         with match::check_epipolar_constraint      (match/base.h:67-79)
 * Side 1 are the rows the reference's outer loop walks (keyframe / keyframe 1), side 2 the candidates (frame / keyframe 2).
 * node1/node2: the BoW node each keypoint belongs to (the key of bow_feat_vec_ that lists it); a row only sees candidates of its
 * own node.  NULL on both sides = every row sees every candidate (robust::).  valid1/valid2 carry the landmark tests of the
 * variant (BOW: row has a live landmark, bow_tree.cc:192-199 / 284-291, candidate likewise for match_keyframes :303-309;
 * TRIANGULATION: neither has a landmark, robust.cc:44-48, 66-69).  All pointers are HOST buffers.
 * Result: match_out[i] = index on side 2 matched to row i, or -1; n_matches. */
enum { B200_PAIRS_BOW = 0, B200_PAIRS_TRIANGULATION = 1 };
typedef struct b200_pairs_problem {
    int32_t n1;
    const uint8_t* desc1;        /* n1 x 32 */
    const float* angle1;         /* needed when check_orientation */
    const uint8_t* valid1;       /* NULL = all */
    const int32_t* node1;        /* NULL = no BoW gating (then node2 must be NULL too) */
    const double* bearing1;      /* TRIANGULATION: n1 x 3 (frm_obs_.bearings_) */
    const float* scale1;         /* TRIANGULATION: orb_params_->scale_factors_[octave] per row */
    const uint8_t* stereo1;      /* TRIANGULATION: stereo_x_right_[i] >= 0; NULL = monocular */
    int32_t n2;
    const uint8_t* desc2;
    const float* angle2;
    const uint8_t* valid2;
    const int32_t* node2;
    const double* bearing2;
    const uint8_t* stereo2;
    double E_12[9];              /* TRIANGULATION: essential matrix, row-major */
    double epiplane_in_keyfrm_2[3]; /* camera centre of keyframe 1 as a bearing in keyframe 2 (robust.cc:22-27) */
    int32_t valid_epiplane;
    float residual_rad_thr;
    int32_t* match_out;          /* out, n1 */
    int32_t n_matches;           /* out */
} b200_pairs_problem_t;
/* max_candidates bounds the gated candidates kept per row (0 = default 64); B200_ERR_CAPACITY reports the size needed. */
int b200_match_pairs(b200_matcher_t h, int n_problems, b200_pairs_problem_t* problems, int variant, float lowe_ratio,
                     int check_orientation, int max_candidates);
/* match::stereo::compute (src/stella_vslam/match/stereo.cc:20-114, helpers :116-251; constructed at system.cc:443 from the two
 * extractors' image_pyramid_).  `left` / `right` are the extractor handles whose last extract produced the two rectified frames
 * (frame index inside that extract's batch; both eyes may also be frames 0 and 1 of ONE handle): their pyramids are read where
 * they already live on the device.  Keypoints / descriptors are HOST buffers (the caller may have filtered them).
 * scale_factors_ / inv_scale_factors_ are the extractor's own (orb_params.cc:37-48).
 * Out: stereo_x_right[i], depths[i] (-1 = no stereo match), n_matched = keypoints that keep one. */
int b200_stereo_compute(b200_matcher_t h, b200_orb_t left, int frame_left, b200_orb_t right, int frame_right,
                        const b200_keypoint_t* keypts_left, const uint8_t* descs_left, int n_left, const b200_keypoint_t* keypts_right,
                        const uint8_t* descs_right, int n_right, float focal_x_baseline, float true_baseline, float* stereo_x_right,
                        float* depths, int32_t* n_matched);
/* data::landmark::compute_descriptor (src/stella_vslam/data/landmark.cc:199-256), for `n_landmarks` landmarks at once
 * (after local BA / fusion every touched landmark is refreshed, local_bundle_adjuster_g2o.cc:387-390, 408).  Landmark l owns the
 * descriptors descs[32 * offsets[l] .. 32 * offsets[l+1]) -- the rows of its observing keyframes that are not about to be erased, in
 * observation order.  best_idx[l] = index (within the landmark) of the descriptor with the smallest median Hamming distance to all of
 * them, first on ties (-1 for a landmark without descriptors); desc_out (optional, n_landmarks x 32) = that descriptor. */
int b200_landmark_descriptors(b200_matcher_t h, int n_landmarks, const uint8_t* descs, const int32_t* offsets, int32_t* best_idx,
                              uint8_t* desc_out);
/* data::landmark::update_mean_normal_and_obs_scale_variance (src/stella_vslam/data/landmark.cc:256-311) for
 * `n_landmarks` landmarks.  Landmark l: position pos_w[3l..], observed from the camera centres
 * cam_centers[3 * offsets[l] .. 3 * offsets[l+1]) (keyfrm->get_trans_wc() of its observations in the order the caller walks them);
 * ref_center[3l..] / ref_scale_factor[l] = centre of its reference keyframe and scale_factors_[octave of its keypoint there];
 * inv_scale_factor_last = inv_scale_factors_[num_levels - 1].  Out: mean_normal (n x 3), max_valid_dist, min_valid_dist. */
int b200_landmark_geometry(b200_matcher_t h, int n_landmarks, const double* pos_w, const int32_t* offsets, const double* cam_centers,
                           const double* ref_center, const float* ref_scale_factor, float inv_scale_factor_last, double* mean_normal,
                           float* max_valid_dist, float* min_valid_dist);
/* Device time (ms, CUDA events) of the two passes of the last b200_match_bruteforce[_device] call while timing is enabled:
 * stage 0 = all-pairs distances + per-row top-K lists, stage 1 = sequential resolve.  Synchronises the streams involved. */
int b200_matcher_enable_timing(b200_matcher_t h, int enable);
int b200_matcher_stage_ms(b200_matcher_t h, int stage, float* ms);
/* Run on the caller's stream (a cudaStream_t; NULL is the legacy default stream); use_own != 0 restores the own stream. */
int b200_matcher_set_stream(b200_matcher_t h, void* stream, int use_own);
int b200_matcher_sync(b200_matcher_t h);

/* ------------------------------------------------------------------------------------------------------------------
 * optimize::local_bundle_adjuster  (src/stella_vslam/optimize/local_bundle_adjuster.h:15-24,
 * optimize/local_bundle_adjuster_g2o.cc:36-431).  The host adapter does the pointer-chasing gather (steps 1-4,
 * :41-304) and the write-back under the map mutex (step 8, :379-430); this entry point is steps 5-7: two rounds of
 * Levenberg-Marquardt with Schur complement over the landmarks (g2o BlockSolver_6_3 semantics), outlier marking in
 * between, on a flattened problem.  All arithmetic is fp64.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t model;          /* 0: perspective-family edge (Perspective / Fisheye / RadialDivision all use the perspective edges on
                               undistorted keypoints, reproj_edge_wrapper.h:64-188); 1: equirectangular (:129-146) */
    double fx, fy, cx, cy;  /* perspective_reproj_edge.h:34 */
    double fxb;             /* focal_x_baseline_ (stereo rows, perspective_reproj_edge.h:148) */
    double cols, rows;      /* equirectangular_reproj_edge.h */
} b200_camera_t;

typedef struct {
    int32_t n_poses, n_points, n_edges, n_cams;
    const double* pose_cw;           /* K x 16 row-major 4x4: keyfrm->get_pose_cw() (shot_vertex_container.h:103-117) */
    const uint8_t* pose_fixed;       /* K: 1 = fixed keyframe (local_bundle_adjuster_g2o.cc:184-190) */
    const double* points;            /* L x 3: lm->get_pos_in_world() */
    const uint8_t* point_fixed;      /* L or NULL: marker corners of keep_fixed_ markers (:272) */
    const int32_t* e_pose;           /* E: keyframe index of the observation */
    const int32_t* e_point;          /* E: landmark index */
    const uint8_t* e_cam;            /* E: index into cams */
    const float* e_obs;              /* E x 3: undist_keypt.pt.x, .y, stereo_x_right (< 0 => monocular edge, reproj_edge_wrapper.h:62) */
    const float* e_inv_sigma_sq;     /* E: inv_level_sigma_sq_[octave] (:238) */
    const float* e_delta;            /* E: Huber delta = sqrt(chi-square) as float (:205-208, 239-241) */
    const uint8_t* e_robust;         /* E or NULL (=1): Huber kernel in the first round (use_huber_loss, :297-299) */
    const uint8_t* e_can_be_outlier; /* E or NULL (=1): 0 for marker-corner edges, which are never outlier-tested (:251-304) */
    const b200_camera_t* cams;
} b200_lba_problem_t;

typedef struct {
    int32_t iterations[2];   /* LM iterations run in the robust / non-robust round */
    int32_t n_outliers;
    double chi2[2];          /* active robust chi-square after each round */
    double lambda_init;
    double lambda_final[2];
} b200_lba_stats_t;

typedef struct b200_lba_s* b200_lba_t;

int b200_lba_create(int device, b200_lba_t* out);
int b200_lba_destroy(b200_lba_t h);
/* local_bundle_adjuster_g2o::optimize steps 5-7.  iters1/iters2: num_first_iter_/num_second_iter_ (5 / 10,
 * local_bundle_adjuster_g2o.h:25-27).  force_stop: the caller's abort flag (mapping_module.cc:124,199-206), polled
 * between LM iterations; may be NULL.  As in the reference it is also WRITTEN: the gain-threshold terminate action
 * sets it when it stops a round (terminate_action.cc:55-72 via g2o's setOptimizerStopFlag), which is what makes the
 * reference skip the second round after an early first-round stop (:317-321).
 * Returns B200_ERR_ABORTED (nothing written) when *force_stop is already set on entry (:308-310).
 * With force_stop == NULL the gain stop of the first round lands in g2o's own auxiliary flag, which the second optimize() resets:
 * the second round always runs.  Deviation: the reference's terminate_action also CLEARS the caller's flag at iteration -1 of each
 * optimize() (terminate_action.cc:46-51), so an abort raised in the few microseconds between the entry test and the first iteration
 * is lost there; here an externally raised flag is never cleared, it stops the solve at the next iteration boundary.
 * At most 166 FREE keyframes (B200_ERR_INVALID beyond; fixed keyframes are not limited).
 * pose_cw_out: K x 16, points_out: L x 3, outlier_out: E (1 = observation to erase, :354-375).  Same code path as the batch of one. */
int b200_lba_solve(b200_lba_t h, const b200_lba_problem_t* problem, int iters1, int iters2, volatile uint8_t* force_stop,
                   double* pose_cw_out, double* points_out, uint8_t* outlier_out, b200_lba_stats_t* stats);
/* Many independent windows in ONE launch sequence (the batched form is what makes local BA GPU-shaped).  The reference
 * runs one optimize() per new keyframe on the mapping thread (mapping_module.cc:199-206); an integrator with several maps / streams
 * (BASELINE config 5) or a backlog of keyframes hands all pending windows to one call.  The windows advance in lockstep, every
 * kernel serves all of them (window = blockIdx.y, one thread-block cluster per window for the reduced-system Cholesky), the
 * Levenberg-Marquardt decisions are taken on the device per window, and the plan (edges sorted by landmark, per-keyframe edge lists)
 * is built on the device: the host copies the caller's arrays and enqueues.
 *   force_stop[w]  : per-window abort flag as in b200_lba_solve (array may be NULL, entries may be NULL)
 *   pose_cw_out[w] / points_out[w] / outlier_out[w] : per-window outputs (outlier_out or its entries may be NULL)
 *   stats          : n_windows entries or NULL
 *   status[w]      : B200_OK, B200_ERR_ABORTED (flag already set on entry: nothing written for that window) or B200_ERR_INVALID
 * Limits: at most 166 FREE keyframes per window (the reduced system is factored on chip); fixed keyframes, landmarks and
 * observations are limited by memory only.  Returns B200_OK if every window that ran is valid. */
int b200_lba_solve_batch(b200_lba_t h, int n_windows, const b200_lba_problem_t* problems, int iters1, int iters2,
                         volatile uint8_t* const* force_stop, double* const* pose_cw_out, double* const* points_out,
                         uint8_t* const* outlier_out, b200_lba_stats_t* stats, int32_t* status);
/* optimize::global_bundle_adjuster (src/stella_vslam/optimize/global_bundle_adjuster.cc: optimize_impl :26-192, optimize :258-420,
 * optimize_for_initialization :201-256; called by module/loop_bundle_adjuster.cc:54 and module/initializer.cc:281): the same flattened
 * problem and kernels as the local bundle adjuster, ONE Levenberg-Marquardt round of `num_iter` iterations (global_bundle_adjuster.h:20-23:
 * 10) with the terminate action at `gain_threshold` (1e-3 in optimize(), the caller's value in optimize_for_initialization), no outlier
 * pass.  Every keyframe of the map is free except the spanning root (:80-81), so the reduced system has 6 x (keyframes - 1) unknowns: up
 * to 1000 it is factored on chip like a local window; beyond that (limit 4000 free keyframes) the dense Cholesky runs panel by panel over
 * the whole GPU from HBM -- two launches per 24 columns -- where the reference uses g2o's CSparse solver (:42-45).
 * use_huber_kernel is the per-edge e_robust array (NULL = Huber on every edge); marker corners as in b200_lba_problem_t.
 * Returns B200_ERR_ABORTED when the CALLER raised *force_stop (the reference returns false then, :340-342, and uses no result: the
 * output buffers are unspecified); a stop
 * by the gain threshold also sets the flag (terminate_action.cc:66-70) but is a normal return. */
int b200_global_ba_solve(b200_lba_t h, const b200_lba_problem_t* problem, int num_iter, double gain_threshold, volatile uint8_t* force_stop,
                         double* pose_cw_out, double* points_out, b200_lba_stats_t* stats);
/* optimize::pose_optimizer::optimize  (src/stella_vslam/optimize/pose_optimizer.h:24-40, pose_optimizer_g2o.cc:38-175; factory
 * defaults num_trials_robust = 2, num_trials = 2, num_each_iter = 10, pose_optimizer_factory.h:18-47): motion-only bundle adjustment
 * of `n_problems` frames in one launch.  Each problem uses the b200_lba_problem_t layout with exactly ONE pose (free), the landmarks
 * the frame observes (all fixed) and one edge per observation (e_pose = 0; e_obs = undistorted x, y, x_right (< 0: monocular edge);
 * e_inv_sigma_sq = inv_level_sigma_sq_[octave]; e_delta = sqrt(chi-square), :84-88); one camera per problem.
 * Out: pose_cw_out [n_problems][16] row-major (the input pose when a frame has fewer than 5 observations, :116-118),
 * outlier_flags = the problems' edges concatenated (outlier_flags.at(idx), :141-160), n_valid[p] = num_init_obs - num_bad_obs. */
int b200_pose_optimize(b200_lba_t h, int n_problems, const b200_lba_problem_t* problems, int num_trials_robust, int num_trials,
                       int num_each_iter, double* pose_cw_out, uint8_t* outlier_flags, uint32_t* n_valid);

/* ------------------------------------------------------------------------------------------------------------------
 * Device-resident tracking chain: the per-frame steady state of tracking_module::track_local_map for `n_frames` independent
 * frames (one per camera / session / replayed log) in ONE launch sequence, reading the extractor's results where they lie
 * in HBM -- nothing of the frame goes back to the host between the stages:
 *   camera::*::undistort_keypoints            camera/perspective.cc:245-275 (frame construction, system.cc:386-395)
 *   tracking_module::search_local_landmarks   tracking_module.cc:533-606: data::frame::can_observe (data/frame.cc:59-84) over the local
 *                                             landmarks, then projection::match_frame_and_landmarks (match/projection.cc:13-93) with
 *                                             lowe_ratio 0.8 and the margin the caller chose (:599-603), HAMMING_DIST_THR_HIGH
 *   pose_optimizer::optimize                  optimize/pose_optimizer_g2o.cc:38-175 on the landmarks the frame now carries
 *                                             (tracking_module::optimize_current_frame_with_local_map)
 * Frame f is frame `frames[f].frame` of the LAST extract on `orb` (host- or device-image variant).  The landmark table of a frame
 * lists, in the reference's iteration order, every landmark the stage touches: the local landmarks AND the landmarks the frame
 * already carries (those with lm_skip = 1: tracking_module.cc:536-551 puts them into curr_landmark_ids and :561-563 skips them).
 * All pointers are HOST buffers; they go up in one copy and the results come back in one copy.  The three handles' arenas are used
 * and everything is enqueued on the extractor's stream; the call returns when the results are in the caller's buffers.
 * Bit-exact against the stage-by-stage host ABI (b200_keypoints_undistort, b200_frame_can_observe, b200_match_guided mode 0) and
 * within 1e-5 for the pose (b200_pose_optimize): same kernels / same device functions. */
typedef struct b200_track_params {
    b200_camera_intrinsics_t cam;
    double focal_x_baseline;       /* camera::base::focal_x_baseline_ (0 for monocular) */
    int32_t monocular;             /* setup_type_ == Monocular: edge threshold sqrt(chi2_2D), else sqrt(chi2_3D) (pose_optimizer_g2o.cc:84-88) */
    float img_bounds[4];           /* min_x, max_x, min_y, max_y */
    int32_t grid_cols, grid_rows;  /* 64 x 48 */
    uint32_t num_levels;
    float log_scale_factor;
    const float* scale_factors;        /* num_levels */
    const float* inv_level_sigma_sq;   /* num_levels */
    float margin;                  /* margin_local_map_projection_(_unstable_) */
    float lowe_ratio;              /* 0.8 (tracking_module.cc:599) */
    uint32_t hamming_thr;          /* HAMMING_DIST_THR_HIGH = 100 */
    float ray_cos_thr;             /* 0.5 (tracking_module.cc:588) */
    int32_t num_trials_robust, num_trials, num_each_iter; /* 2 / 2 / 10 */
    int32_t max_candidates;        /* 0 = 256 keypoints per search window */
} b200_track_params_t;
typedef struct b200_track_frame {
    int32_t frame;                     /* index into the extractor's last batch */
    const double* pose_cw;             /* 16, row-major: curr_frm_.pose_cw_ entering the stage */
    int32_t n_keypoints_in;            /* entries of kp_x_right / kp_landmark (0 when both are NULL); must equal the frame's keypoint count */
    const float* kp_x_right;           /* frm_obs_.stereo_x_right_, NULL when empty */
    const int32_t* kp_landmark;        /* per keypoint: row of the landmark it already carries (not will_be_erased), -1 none; NULL = none */
    int32_t n_landmarks;
    const double* lm_pos_w;            /* 3 per landmark */
    const double* lm_mean_normal;      /* 3 per landmark */
    const float* lm_min_valid_dist;
    const float* lm_max_valid_dist;
    const uint8_t* lm_desc;            /* 32 per landmark */
    const uint8_t* lm_skip;            /* 1 = not searched (already in the frame, will_be_erased, temporal-ratio test :565-585); NULL = none */
    const uint8_t* lm_has_observation; /* landmark::has_observation(); NULL = all have */
    int32_t kp_cap;                    /* capacity of kp_landmark_out / kp_outlier (>= the frame's keypoint count) */
    uint8_t* lm_observable;            /* out, n_landmarks: searched and can_observe() held (the caller's increase_num_observable, :594) */
    int32_t* kp_landmark_out;          /* out, per keypoint: landmark row after the search (frm.add_landmark applied in order), -1 none */
    uint8_t* kp_outlier;               /* out, per keypoint: outlier_flags of the pose optimisation */
    double pose_cw_out[16];            /* out (the input pose when fewer than 5 observations, :116-118) */
    int32_t n_keypoints;               /* out */
    int32_t n_matches;                 /* out: return value of match_frame_and_landmarks */
    uint32_t n_valid;                  /* out: return value of pose_optimizer::optimize */
} b200_track_frame_t;
int b200_track_local_map(b200_orb_t orb, b200_matcher_t matcher, b200_lba_t opt, const b200_track_params_t* prm, int n_frames,
                         b200_track_frame_t* frames);
/* Device time of the last b200_track_local_map, per stage: 0 undistort + can_observe + query build, 1 grid, 2 candidates, 3 resolve,
 * 4 edge build, 5 pose optimisation + scatter, 6 whole chain (CUDA events on the stream). */
int b200_track_stage_ms(b200_matcher_t matcher, int stage, float* ms);

/* Profiling mode: an event after every launch of the following solves (adds a few microseconds per launch; off by default).
 * b200_lba_kernel_ms reports, for the LAST batch, the summed device time and the number of intervals of
 *   kernel 0 plan (5 launches, one interval), 1 landmark pass / build, 2 keyframe rows, 3 Schur rows, 4 reduced-system Cholesky,
 *   5 back-substitution + chi2 of the trial state, 6 (unused since the trial pass was fused into 5), 7 everything after the last
 *   repetition (round tails, outliers, export). */
int b200_lba_enable_profile(b200_lba_t h, int enable);
int b200_lba_kernel_ms(b200_lba_t h, int kernel, float* total_ms, int* launches);
/* Device time (ms, CUDA events) spent in the kernels of the last solve, and the number of kernel launches. */
int b200_lba_last_profile(b200_lba_t h, float* gpu_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* B200VSLAM_H */
