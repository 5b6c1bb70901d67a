// b200vslam.hpp -- C++ host-side mirror of the reference's class surfaces on top of the C ABI (b200vslam.h).
//
// Header-only, no OpenCV/Eigen/g2o dependency: images are raw 8-bit buffers, keypoints are b200_keypoint_t.  The names,
// constructor arguments and error behaviour follow the reference so that the thin adapters in
// stella_vslam_b200/host/reference_adapters/ (which DO include the reference's headers) are one-liners:
//   b200::feature::orb_params      <->  stella_vslam::feature::orb_params      (feature/orb_params.h:11-54)
//   b200::feature::orb_extractor   <->  stella_vslam::feature::orb_extractor   (feature/orb_extractor.h:46-122)
//   b200::match::robust            <->  stella_vslam::match::robust            (match/robust.h, match/base.h:81-91)
//   b200::match::projection / fuse / area / bow_tree / stereo  <->  match/projection.h, fuse.h, area.h, bow_tree.h, stereo.h
//   b200::optimize::local_bundle_adjuster <-> stella_vslam::optimize::local_bundle_adjuster (optimize/local_bundle_adjuster.h:15-24)
//   b200::optimize::pose_optimizer        <-> stella_vslam::optimize::pose_optimizer        (optimize/pose_optimizer.h:24-40)
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200vslam.h"

namespace b200 {

inline void check(int rc, const char* what) {
    if (rc != B200_OK) throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

namespace feature {

enum class descriptor_type { ORB, HASH_SIFT };  // feature/orb_extractor.h:17-44

struct orb_params {  // feature/orb_params.cc:12-71 (float recurrences, not pow)
    std::string name_;
    float scale_factor_ = 1.2f;
    float log_scale_factor_;
    unsigned int num_levels_ = 8, ini_fast_thr_ = 20, min_fast_thr_ = 7;
    std::vector<float> scale_factors_, inv_scale_factors_, level_sigma_sq_, inv_level_sigma_sq_;

    explicit orb_params(const std::string& name = "default ORB feature extraction setting", float scale_factor = 1.2f,
                        unsigned int num_levels = 8, unsigned int ini_fast_thr = 20, unsigned int min_fast_thr = 7)
        : name_(name), scale_factor_(scale_factor), log_scale_factor_(std::log(scale_factor)), num_levels_(num_levels),
          ini_fast_thr_(ini_fast_thr), min_fast_thr_(min_fast_thr) {
        scale_factors_.assign(num_levels, 1.0f);
        inv_scale_factors_.assign(num_levels, 1.0f);
        level_sigma_sq_.assign(num_levels, 1.0f);
        inv_level_sigma_sq_.assign(num_levels, 1.0f);
        float s = 1.0f;
        for (unsigned int l = 1; l < num_levels; ++l) {
            scale_factors_[l] = scale_factor * scale_factors_[l - 1];
            inv_scale_factors_[l] = (1.0f / scale_factor) * inv_scale_factors_[l - 1];
            s = scale_factor * s;
            level_sigma_sq_[l] = s * s;
            inv_level_sigma_sq_[l] = 1.0f / (s * s);
        }
    }
};

class orb_extractor {
public:
    // orb_extractor(const orb_params*, unsigned min_area, descriptor_type, mask_rects) -- orb_extractor.h:51-54
    orb_extractor(const orb_params* params, unsigned int min_area, descriptor_type desc_type = descriptor_type::ORB,
                  const std::vector<std::vector<float>>& mask_rects = {}, int device = 0, int max_batch = 1)
        : orb_params_(params), mask_rects_(mask_rects) {
        if (desc_type == descriptor_type::HASH_SIFT) throw std::runtime_error("cuda_efficient_features is not available");  // orb_extractor.cc:121
        b200_orb_params_t p;
        b200_orb_default_params(&p);
        p.scale_factor = params->scale_factor_;
        p.num_levels = (int32_t)params->num_levels_;
        p.ini_fast_thr = (int32_t)params->ini_fast_thr_;
        p.min_fast_thr = (int32_t)params->min_fast_thr_;
        p.min_area = min_area;
        for (const auto& r : mask_rects) flat_rects_.insert(flat_rects_.end(), r.begin(), r.begin() + 4);
        p.n_mask_rects = (int32_t)mask_rects.size();
        p.mask_rects = flat_rects_.empty() ? nullptr : flat_rects_.data();
        p.device = device;
        p.max_batch = max_batch;
        check(b200_orb_create(&p, &h_), "b200_orb_create");
    }
    ~orb_extractor() { b200_orb_destroy(h_); }
    orb_extractor(const orb_extractor&) = delete;
    orb_extractor& operator=(const orb_extractor&) = delete;

    // extract(in_image, in_image_mask, keypts, out_descriptors) -- orb_extractor.h:60-61.  One CV_8UC1 frame in host memory;
    // descriptors come back as N x 32 bytes (cv::Mat(N, 32, CV_8U) layout).  Empty image -> silent return (orb_extractor.cc:30-32).
    void extract(const uint8_t* image, int width, int height, size_t pitch, const uint8_t* mask, size_t mask_pitch,
                 std::vector<b200_keypoint_t>& keypts, std::vector<uint8_t>& descriptors) {
        extract_batch(image, width, height, pitch, pitch * (size_t)height, 1, mask, mask_pitch, keypts, descriptors, counts_);
        keypts.resize(counts_.empty() ? 0 : counts_[0]);
        descriptors.resize(keypts.size() * 32);
    }
    // `batch` same-sized frames; frame f's results start at f * cap (cap = max_keypoints(width, height)).
    void extract_batch(const uint8_t* images, int width, int height, size_t pitch, size_t frame_stride, int batch, const uint8_t* mask,
                       size_t mask_pitch, std::vector<b200_keypoint_t>& keypts, std::vector<uint8_t>& descriptors, std::vector<int32_t>& counts) {
        keypts.clear();
        descriptors.clear();
        counts.assign(batch > 0 ? batch : 0, 0);
        if (!images || width == 0 || height == 0 || batch == 0) return;
        const int cap = max_keypoints(width, height);
        keypts.resize((size_t)cap * batch);
        descriptors.resize((size_t)cap * batch * 32);
        check(b200_orb_extract(h_, images, width, height, pitch, frame_stride, batch, mask, mask_pitch, keypts.data(), descriptors.data(), cap,
                               counts.data()),
              "b200_orb_extract");
    }
    int max_keypoints(int width, int height) const { return b200_orb_max_keypoints(h_, width, height); }
    // image_pyramid_ (orb_extractor.h:71): level >= 1 of the last extract, tightly packed
    std::vector<uint8_t> pyramid_level(int frame, int level, int* w = nullptr, int* hgt = nullptr) const {
        int lw = 0, lh = 0;
        check(b200_orb_level_info(h_, level, &lw, &lh, nullptr, nullptr), "b200_orb_level_info");
        std::vector<uint8_t> out((size_t)lw * lh);
        check(b200_orb_pyramid_level_host(h_, frame, level, out.data(), (size_t)lw), "b200_orb_pyramid_level_host");
        if (w) *w = lw;
        if (hgt) *hgt = lh;
        return out;
    }
    b200_orb_t handle() const { return h_; }

    const orb_params* orb_params_;                  // orb_extractor.h:64
    std::vector<std::vector<float>> mask_rects_;    // orb_extractor.h:68

private:
    b200_orb_t h_ = nullptr;
    std::vector<float> flat_rects_;
    std::vector<int32_t> counts_;
};

}  // namespace feature

namespace match {

constexpr unsigned int HAMMING_DIST_THR_LOW = 50, HAMMING_DIST_THR_HIGH = 100, MAX_HAMMING_DIST = 256;  // match/base.h:15-17

class base {  // match/base.h:81-91
public:
    base(float lowe_ratio, bool check_orientation) : lowe_ratio_(lowe_ratio), check_orientation_(check_orientation) {}
    virtual ~base() = default;

protected:
    const float lowe_ratio_;
    const bool check_orientation_;
};

class robust final : public base {
public:
    robust(float lowe_ratio, bool check_orientation, int device = 0) : base(lowe_ratio, check_orientation) {
        check(b200_matcher_create(device, &h_), "b200_matcher_create");
    }
    ~robust() override { b200_matcher_destroy(h_); }
    // brute_force_match (match/robust.cc:232-328): frame keypoints (descriptors 32 B each, angles with a byte stride) against the
    // keyframe's; keyfrm_has_landmark[i] != 0 <=> lms_2[i] && !will_be_erased().  Returns (idx_1, idx_2) sorted by idx_1.
    unsigned int brute_force_match(const uint8_t* frm_desc, const void* frm_angles, size_t frm_angle_stride, int n1, const uint8_t* keyfrm_desc,
                                   const void* keyfrm_angles, size_t keyfrm_angle_stride, const uint8_t* keyfrm_has_landmark, int n2,
                                   std::vector<std::pair<int, int>>& matches) const {
        matches.clear();
        if (n1 <= 0 || n2 <= 0) return 0;
        std::vector<int32_t> pairs((size_t)2 * n1);
        const int32_t off = 0;
        int32_t n = 0;
        check(b200_match_bruteforce(h_, 1, frm_desc, frm_angles, frm_angle_stride, &off, &n1, keyfrm_desc, keyfrm_angles, keyfrm_angle_stride,
                                    keyfrm_has_landmark, &off, &n2, lowe_ratio_, check_orientation_ ? 1 : 0, pairs.data(), n1, &n),
              "b200_match_bruteforce");
        matches.reserve(n);
        for (int i = 0; i < n; ++i) matches.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
        return (unsigned int)n;
    }

    // match_for_triangulation (match/robust.cc:14-146): the problem carries bearings, E_12, the epipole and the "has no landmark"
    // masks; returns matched_idx_pairs sorted by idx_1.
    unsigned int match_for_triangulation(b200_pairs_problem_t& problem, std::vector<std::pair<unsigned int, unsigned int>>& matched_idx_pairs) const {
        return run_pairs(h_, problem, B200_PAIRS_TRIANGULATION, lowe_ratio_, check_orientation_, matched_idx_pairs);
    }
    b200_matcher_t handle() const { return h_; }

    static unsigned int run_pairs(b200_matcher_t h, b200_pairs_problem_t& problem, int variant, float lowe_ratio, bool check_orientation,
                                  std::vector<std::pair<unsigned int, unsigned int>>& out) {
        std::vector<int32_t> match((size_t)(problem.n1 > 0 ? problem.n1 : 1), -1);
        problem.match_out = match.data();
        check(b200_match_pairs(h, 1, &problem, variant, lowe_ratio, check_orientation ? 1 : 0, 0), "b200_match_pairs");
        out.clear();
        for (int i = 0; i < problem.n1; ++i)
            if (match[i] >= 0) out.emplace_back((unsigned int)i, (unsigned int)match[i]);
        problem.match_out = nullptr;
        return (unsigned int)problem.n_matches;
    }

private:
    b200_matcher_t h_ = nullptr;
};

// One matcher handle (stream + staging) shared by the guided / all-pairs / stereo mirrors below.
class device_matcher {
public:
    explicit device_matcher(int device = 0) { check(b200_matcher_create(device, &h_), "b200_matcher_create"); }
    ~device_matcher() { b200_matcher_destroy(h_); }
    device_matcher(const device_matcher&) = delete;
    device_matcher& operator=(const device_matcher&) = delete;
    b200_matcher_t get() const { return h_; }

private:
    b200_matcher_t h_ = nullptr;
};

// match::projection (match/projection.h:24-67).  The adapter fills one b200_guided_problem_t per call (landmarks in the reference's
// order, reprojected; see reference_adapters/projection_b200.cc); match_out / n_matches come back in the struct.
class projection final : public base {
public:
    explicit projection(float lowe_ratio = 0.6f, bool check_orientation = true, int device = 0) : base(lowe_ratio, check_orientation), m_(device) {}
    unsigned int match_frame_and_landmarks(b200_guided_problem_t& p) const { return run(p, B200_GUIDED_LANDMARKS, HAMMING_DIST_THR_HIGH, false); }
    unsigned int match_current_and_last_frames(b200_guided_problem_t& p) const {
        return run(p, B200_GUIDED_LAST_FRAME, HAMMING_DIST_THR_HIGH, check_orientation_);
    }
    unsigned int match_frame_and_keyframe(b200_guided_problem_t& p, unsigned int hamm_dist_thr) const {
        return run(p, B200_GUIDED_LAST_FRAME, hamm_dist_thr, check_orientation_);
    }
    unsigned int match_by_Sim3_transform(b200_guided_problem_t& p) const { return run(p, B200_GUIDED_LAST_FRAME, HAMMING_DIST_THR_LOW, false); }
    // match_keyframes_mutually: p12 = landmarks of keyframe 1 searched in keyframe 2, p21 the other direction; mutual[i] = idx_2 or -1
    unsigned int match_keyframes_mutually(b200_guided_problem_t& p12, b200_guided_problem_t& p21, std::vector<int32_t>& mutual) const {
        b200_guided_problem_t both[2] = {p12, p21};
        check(b200_match_guided(m_.get(), 2, both, B200_GUIDED_INDEPENDENT, HAMMING_DIST_THR_HIGH, lowe_ratio_, 0, 0), "b200_match_guided");
        mutual.assign((size_t)(p12.n_queries > 0 ? p12.n_queries : 1), -1);
        int32_t n = 0;
        check(b200_match_cross_check(p12.match_out, p12.n_queries, p21.match_out, p21.n_queries, mutual.data(), &n), "b200_match_cross_check");
        mutual.resize((size_t)p12.n_queries);
        return (unsigned int)n;
    }

private:
    unsigned int run(b200_guided_problem_t& p, int mode, unsigned int thr, bool orientation) const {
        check(b200_match_guided(m_.get(), 1, &p, mode, thr, lowe_ratio_, orientation ? 1 : 0, 0), "b200_match_guided");
        return (unsigned int)p.n_matches;
    }
    device_matcher m_;
};

class fuse final : public base {  // match/fuse.h, fuse.cc:12-154
public:
    explicit fuse(float lowe_ratio = 0.6f, bool check_orientation = true, int device = 0) : base(lowe_ratio, check_orientation), m_(device) {}
    unsigned int detect_duplication(b200_guided_problem_t& p) const {
        check(b200_match_guided(m_.get(), 1, &p, B200_GUIDED_FUSE, HAMMING_DIST_THR_LOW, lowe_ratio_, 0, 0), "b200_match_guided");
        return (unsigned int)p.n_matches;
    }

private:
    device_matcher m_;
};

class area final : public base {  // match/area.h, area.cc:8-98
public:
    explicit area(float lowe_ratio = 0.9f, bool check_orientation = true, int device = 0) : base(lowe_ratio, check_orientation), m_(device) {}
    unsigned int match_in_consistent_area(b200_guided_problem_t& p) const {
        check(b200_match_guided(m_.get(), 1, &p, B200_GUIDED_AREA, HAMMING_DIST_THR_LOW, lowe_ratio_, check_orientation_ ? 1 : 0, 0), "b200_match_guided");
        return (unsigned int)p.n_matches;
    }

private:
    device_matcher m_;
};

class bow_tree final : public base {  // match/bow_tree.h, bow_tree.cc:11-366 (node1 / node2 = BoW node of every keypoint)
public:
    explicit bow_tree(float lowe_ratio = 0.6f, bool check_orientation = true, int device = 0) : base(lowe_ratio, check_orientation), m_(device) {}
    unsigned int match_for_triangulation(b200_pairs_problem_t& p, std::vector<std::pair<unsigned int, unsigned int>>& matched_idx_pairs) const {
        return robust::run_pairs(m_.get(), p, B200_PAIRS_TRIANGULATION, lowe_ratio_, check_orientation_, matched_idx_pairs);
    }
    // match_frame_and_keyframe / match_keyframes: (row, candidate) pairs; the adapter maps them to matched_lms_in_frm / _in_keyfrm_1
    unsigned int match_frame_and_keyframe(b200_pairs_problem_t& p, std::vector<std::pair<unsigned int, unsigned int>>& pairs) const {
        return robust::run_pairs(m_.get(), p, B200_PAIRS_BOW, lowe_ratio_, check_orientation_, pairs);
    }
    unsigned int match_keyframes(b200_pairs_problem_t& p, std::vector<std::pair<unsigned int, unsigned int>>& pairs) const {
        return robust::run_pairs(m_.get(), p, B200_PAIRS_BOW, lowe_ratio_, check_orientation_, pairs);
    }

private:
    device_matcher m_;
};

// match::stereo (match/stereo.h:17-101): built from the two extractors whose pyramids stay on the device (system.cc:443)
class stereo {
public:
    stereo(const feature::orb_extractor& left, const feature::orb_extractor& right, const std::vector<b200_keypoint_t>& keypts_left,
           const std::vector<b200_keypoint_t>& keypts_right, const std::vector<uint8_t>& descs_left, const std::vector<uint8_t>& descs_right,
           float focal_x_baseline, float true_baseline, int frame_left = 0, int frame_right = 0, int device = 0)
        : left_(left), right_(right), kl_(keypts_left), kr_(keypts_right), dl_(descs_left), dr_(descs_right), fxb_(focal_x_baseline),
          baseline_(true_baseline), fl_(frame_left), fr_(frame_right), m_(device) {}
    void compute(std::vector<float>& stereo_x_right, std::vector<float>& depths) const {
        stereo_x_right.assign(kl_.size(), -1.0f);
        depths.assign(kl_.size(), -1.0f);
        int32_t n = 0;
        check(b200_stereo_compute(m_.get(), left_.handle(), fl_, right_.handle(), fr_, kl_.data(), dl_.data(), (int)kl_.size(), kr_.data(), dr_.data(),
                                  (int)kr_.size(), fxb_, baseline_, stereo_x_right.data(), depths.data(), &n),
              "b200_stereo_compute");
    }

private:
    const feature::orb_extractor &left_, &right_;
    const std::vector<b200_keypoint_t>&kl_, &kr_;
    const std::vector<uint8_t>&dl_, &dr_;
    float fxb_, baseline_;
    int fl_, fr_;
    device_matcher m_;
};

}  // namespace match

namespace optimize {

class local_bundle_adjuster {  // optimize/local_bundle_adjuster_g2o.h:16-46 with Mapping.backend: "b200"
public:
    explicit local_bundle_adjuster(unsigned int num_first_iter = 5, unsigned int num_second_iter = 10, int device = 0)
        : num_first_iter_(num_first_iter), num_second_iter_(num_second_iter) {
        check(b200_lba_create(device, &h_), "b200_lba_create");
    }
    ~local_bundle_adjuster() { b200_lba_destroy(h_); }
    local_bundle_adjuster(const local_bundle_adjuster&) = delete;
    // steps 5-7 of local_bundle_adjuster_g2o::optimize on the flattened window; returns false when the abort flag was already set
    // (local_bundle_adjuster_g2o.cc:308-310), in which case nothing is written.
    bool optimize(const b200_lba_problem_t& problem, volatile uint8_t* force_stop_flag, std::vector<double>& pose_cw_out,
                  std::vector<double>& points_out, std::vector<uint8_t>& outlier_out, b200_lba_stats_t* stats = nullptr) const {
        pose_cw_out.resize((size_t)16 * problem.n_poses);
        points_out.resize((size_t)3 * problem.n_points);
        outlier_out.resize((size_t)problem.n_edges);
        const int rc = b200_lba_solve(h_, &problem, (int)num_first_iter_, (int)num_second_iter_, force_stop_flag, pose_cw_out.data(),
                                      points_out.data(), outlier_out.data(), stats);
        if (rc == B200_ERR_ABORTED) return false;
        check(rc, "b200_lba_solve");
        return true;
    }

    // many windows per launch sequence (b200_lba_solve_batch); ok[w] = false for windows whose flag was already set
    void optimize_batch(const std::vector<b200_lba_problem_t>& problems, const std::vector<volatile uint8_t*>& force_stop_flags,
                        std::vector<std::vector<double>>& pose_cw_out, std::vector<std::vector<double>>& points_out,
                        std::vector<std::vector<uint8_t>>& outlier_out, std::vector<bool>& ok) const {
        const size_t n = problems.size();
        pose_cw_out.resize(n); points_out.resize(n); outlier_out.resize(n);
        std::vector<double*> pp(n), qq(n);
        std::vector<uint8_t*> oo(n);
        std::vector<int32_t> status(n, 0);
        for (size_t w = 0; w < n; ++w) {
            pose_cw_out[w].resize((size_t)16 * problems[w].n_poses);
            points_out[w].resize((size_t)3 * problems[w].n_points);
            outlier_out[w].resize((size_t)(problems[w].n_edges > 0 ? problems[w].n_edges : 1));
            pp[w] = pose_cw_out[w].data(); qq[w] = points_out[w].data(); oo[w] = outlier_out[w].data();
        }
        const int rc = b200_lba_solve_batch(h_, (int)n, problems.data(), (int)num_first_iter_, (int)num_second_iter_,
                                            force_stop_flags.empty() ? nullptr : force_stop_flags.data(), pp.data(), qq.data(), oo.data(), nullptr, status.data());
        ok.assign(n, true);
        for (size_t w = 0; w < n; ++w)
            if (status[w] == B200_ERR_ABORTED) ok[w] = false;
            else if (status[w] != B200_OK) check(status[w], "b200_lba_solve_batch");
        if (rc != B200_OK && rc != B200_ERR_ABORTED) check(rc, "b200_lba_solve_batch");
    }

private:
    const unsigned int num_first_iter_, num_second_iter_;
    b200_lba_t h_ = nullptr;
};

class global_bundle_adjuster {  // optimize/global_bundle_adjuster.h:18-62
public:
    explicit global_bundle_adjuster(unsigned int num_iter = 10, bool use_huber_kernel = true, int device = 0)
        : num_iter_(num_iter), use_huber_kernel_(use_huber_kernel) {
        check(b200_lba_create(device, &h_), "b200_lba_create");
    }
    ~global_bundle_adjuster() { b200_lba_destroy(h_); }
    global_bundle_adjuster(const global_bundle_adjuster&) = delete;
    // one LM round over the flattened map (problem.e_robust carries use_huber_kernel per edge); false = aborted by the caller's flag
    bool optimize(const b200_lba_problem_t& problem, volatile uint8_t* force_stop_flag, std::vector<double>& pose_cw_out, std::vector<double>& points_out,
                  double gain_threshold = 1e-3) const {
        pose_cw_out.resize((size_t)16 * problem.n_poses);
        points_out.resize((size_t)3 * problem.n_points);
        const int rc = b200_global_ba_solve(h_, &problem, (int)num_iter_, gain_threshold, force_stop_flag, pose_cw_out.data(), points_out.data(), nullptr);
        if (rc == B200_ERR_ABORTED) return false;
        check(rc, "b200_global_ba_solve");
        return true;
    }
    const unsigned int num_iter_;
    const bool use_huber_kernel_;

private:
    b200_lba_t h_ = nullptr;
};

class pose_optimizer {  // optimize/pose_optimizer.h:24-40 with Tracking.backend: "b200" (pose_optimizer_g2o.h:30-37 defaults)
public:
    explicit pose_optimizer(unsigned int num_trials_robust = 2, unsigned int num_trials = 2, unsigned int num_each_iter = 10, int device = 0)
        : num_trials_robust_(num_trials_robust), num_trials_(num_trials), num_each_iter_(num_each_iter) {
        check(b200_lba_create(device, &h_), "b200_lba_create");
    }
    ~pose_optimizer() { b200_lba_destroy(h_); }
    pose_optimizer(const pose_optimizer&) = delete;
    // one flattened frame (one free pose, fixed landmarks, one edge per observation); returns num_init_obs - num_bad_obs
    unsigned int optimize(const b200_lba_problem_t& frame, double (&optimized_pose)[16], std::vector<bool>& outlier_flags) const {
        std::vector<uint8_t> flags((size_t)(frame.n_edges > 0 ? frame.n_edges : 1));
        uint32_t n_valid = 0;
        check(b200_pose_optimize(h_, 1, &frame, (int)num_trials_robust_, (int)num_trials_, (int)num_each_iter_, optimized_pose, flags.data(), &n_valid),
              "b200_pose_optimize");
        outlier_flags.assign((size_t)frame.n_edges, false);
        for (int e = 0; e < frame.n_edges; ++e) outlier_flags[e] = flags[e] != 0;
        return n_valid;
    }
    b200_lba_t handle() const { return h_; }

private:
    const unsigned int num_trials_robust_, num_trials_, num_each_iter_;
    b200_lba_t h_ = nullptr;
};

}  // namespace optimize

namespace tracking {
// tracking_module::search_local_landmarks + pose_optimizer::optimize for a batch of frames that stay on the GPU (b200_track_local_map)
class local_map_tracker {
public:
    local_map_tracker(const feature::orb_extractor& extractor, const b200_track_params_t& params, int device = 0)
        : ex_(extractor), prm_(params), m_(device), opt_(params.num_trials_robust, params.num_trials, params.num_each_iter, device) {}
    void track(std::vector<b200_track_frame_t>& frames) {
        check(b200_track_local_map(ex_.handle(), m_.get(), opt_.handle(), &prm_, (int)frames.size(), frames.data()), "b200_track_local_map");
    }

private:
    const feature::orb_extractor& ex_;
    b200_track_params_t prm_;
    match::device_matcher m_;
    optimize::pose_optimizer opt_;
};
}  // namespace tracking
}  // namespace b200
