// Replacement bodies for the two per-frame guided matchers
//   match::projection::match_frame_and_landmarks       (src/stella_vslam/match/projection.cc:13-93)
//   match::projection::match_current_and_last_frames   (src/stella_vslam/match/projection.cc:95-207)
// Guard the originals with #ifndef USE_B200 and add this TU to src/stella_vslam/match/CMakeLists.txt.  The map-side walk
// (which landmarks, their reprojection, the level window, validity) stays here, in the reference's order; the grid search, the
// gates, the distances and the sequential "a keypoint takes one landmark" state run in b200_match_guided.
// The other guided matchers bind the same way (same flattening helper, different mode / threshold):
//   match_frame_and_keyframe  -> mode B200_GUIDED_LAST_FRAME, thr = hamm_dist_thr, t_x_right = nullptr, occupied = frm_landmarks[i] != null
//   match_by_Sim3_transform   -> mode B200_GUIDED_LAST_FRAME, thr = HAMMING_DIST_THR_LOW, check_orientation = 0
//   match_keyframes_mutually  -> mode B200_GUIDED_INDEPENDENT once per direction (two problems in one call) + b200_match_cross_check
//   fuse::detect_duplication  -> mode B200_GUIDED_FUSE, thr = HAMMING_DIST_THR_LOW, q_reproj / inv_level_sigma_sq / do_reprojection_matching
//   area::match_in_consistent_area -> mode B200_GUIDED_AREA, thr = HAMMING_DIST_THR_LOW, queries = level-0 keypoints of frame 1
#include "stella_vslam/match/projection.h"
#include "stella_vslam/camera/base.h"
#include "stella_vslam/data/frame.h"
#include "stella_vslam/data/landmark.h"

#include "b200vslam.h"

namespace stella_vslam {
namespace match {
namespace {

b200_matcher_t matcher() {
    static thread_local b200_matcher_t h = nullptr;
    if (!h && b200_matcher_create(0, &h) != B200_OK) throw std::runtime_error(b200_last_error());
    return h;
}

// the searched side of a problem: SoA views of frm_obs_ + camera bounds
struct train_side {
    std::vector<float> x, y, angle;
    std::vector<uint8_t> octave, occupied;
    void fill(const data::frame& frm, b200_guided_problem_t& P) {
        const auto& kps = frm.frm_obs_.undist_keypts_;
        const size_t n = kps.size();
        x.resize(n); y.resize(n); angle.resize(n); octave.resize(n); occupied.resize(n);
        for (size_t i = 0; i < n; ++i) {
            x[i] = kps[i].pt.x; y[i] = kps[i].pt.y; angle[i] = kps[i].angle; octave[i] = static_cast<uint8_t>(kps[i].octave);
            const auto& lm = frm.get_landmark(i);
            occupied[i] = lm && lm->has_observation();  // projection.cc:50-53, 163-166
        }
        P.n_train = static_cast<int32_t>(n);
        P.t_x = x.data(); P.t_y = y.data(); P.t_octave = octave.data(); P.t_angle = angle.data();
        P.t_x_right = frm.frm_obs_.stereo_x_right_.empty() ? nullptr : frm.frm_obs_.stereo_x_right_.data();
        P.t_desc = frm.frm_obs_.descriptors_.data;
        P.t_occupied = occupied.data();
        const auto& b = frm.camera_->img_bounds_;
        P.min_x = b.min_x_; P.max_x = b.max_x_; P.min_y = b.min_y_; P.max_y = b.max_y_;
        P.grid_cols = static_cast<int32_t>(frm.frm_obs_.num_grid_cols_);
        P.grid_rows = static_cast<int32_t>(frm.frm_obs_.num_grid_rows_);
    }
};

struct query_side {
    std::vector<uint8_t> desc, valid;
    std::vector<float> x, y, margin, x_right, angle;
    std::vector<int8_t> lo, hi;
    std::vector<int32_t> out;
    void resize(size_t n) {
        desc.assign(32 * n, 0); valid.assign(n, 0); x.assign(n, 0.f); y.assign(n, 0.f); margin.assign(n, 0.f); x_right.assign(n, 0.f);
        angle.assign(n, 0.f); lo.assign(n, -1); hi.assign(n, -1); out.assign(n, -1);
    }
    void bind(b200_guided_problem_t& P) {
        P.n_queries = static_cast<int32_t>(valid.size());
        P.q_desc = desc.data(); P.q_x = x.data(); P.q_y = y.data(); P.q_margin = margin.data(); P.q_min_level = lo.data(); P.q_max_level = hi.data();
        P.q_x_right = x_right.data(); P.q_angle = angle.data(); P.q_valid = valid.data(); P.match_out = out.data();
    }
};

}  // namespace

unsigned int projection::match_frame_and_landmarks(data::frame& frm, const std::vector<std::shared_ptr<data::landmark>>& local_landmarks,
                                                   eigen_alloc_unord_map<unsigned int, Vec2_t>& lm_to_reproj,
                                                   std::unordered_map<unsigned int, float>& lm_to_x_right,
                                                   std::unordered_map<unsigned int, unsigned int>& lm_to_scale, const float margin) const {
    b200_guided_problem_t P{};
    train_side T;
    T.fill(frm, P);
    query_side Q;
    Q.resize(local_landmarks.size());
    const int num_levels = static_cast<int>(frm.orb_params_->num_levels_);
    for (size_t q = 0; q < local_landmarks.size(); ++q) {
        const auto& lm = local_landmarks[q];
        if (!lm_to_reproj.count(lm->id_) || lm->will_be_erased()) continue;  // :23-28
        const Vec2_t reproj = lm_to_reproj.at(lm->id_);
        const auto level = lm_to_scale.at(lm->id_);
        Q.valid[q] = 1;
        Q.x[q] = static_cast<float>(reproj(0));
        Q.y[q] = static_cast<float>(reproj(1));
        Q.margin[q] = margin * frm.orb_params_->scale_factors_.at(level);  // :36
        Q.lo[q] = static_cast<int8_t>(std::max(0, static_cast<int>(level) - 1));
        Q.hi[q] = static_cast<int8_t>(std::min(num_levels - 1, static_cast<int>(level) + 1));
        if (P.t_x_right) Q.x_right[q] = lm_to_x_right.at(lm->id_);
        const cv::Mat d = lm->get_descriptor();
        std::memcpy(&Q.desc[32 * q], d.data, 32);
    }
    Q.bind(P);
    if (b200_match_guided(matcher(), 1, &P, B200_GUIDED_LANDMARKS, HAMMING_DIST_THR_HIGH, lowe_ratio_, 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    for (size_t q = 0; q < local_landmarks.size(); ++q)
        if (Q.out[q] >= 0) frm.add_landmark(local_landmarks[q], static_cast<unsigned int>(Q.out[q]));  // :87
    return static_cast<unsigned int>(P.n_matches);
}

unsigned int projection::match_current_and_last_frames(data::frame& curr_frm, const data::frame& last_frm, const float margin) const {
    const Mat33_t rot_cw = curr_frm.get_rot_cw();
    const Vec3_t trans_cw = curr_frm.get_trans_cw();
    const Vec3_t trans_wc = -rot_cw.transpose() * trans_cw;
    const Vec3_t trans_lc = last_frm.get_rot_cw() * trans_wc + last_frm.get_trans_cw();
    const bool mono = curr_frm.camera_->setup_type_ == camera::setup_type_t::Monocular;
    const bool assume_forward = mono ? false : trans_lc(2) > curr_frm.camera_->true_baseline_;     // :110-113
    const bool assume_backward = mono ? false : -trans_lc(2) > curr_frm.camera_->true_baseline_;  // :114-117

    b200_guided_problem_t P{};
    train_side T;
    T.fill(curr_frm, P);
    const auto& last_kps = last_frm.frm_obs_.undist_keypts_;
    query_side Q;
    Q.resize(last_kps.size());
    std::vector<std::shared_ptr<data::landmark>> lms(last_kps.size());
    const int num_levels = static_cast<int>(last_frm.orb_params_->num_levels_);
    for (unsigned int idx_last = 0; idx_last < last_kps.size(); ++idx_last) {
        const auto& lm = last_frm.get_landmark(idx_last);
        if (!lm || lm->will_be_erased()) continue;  // :124-129
        Vec2_t reproj;
        float x_right;
        if (!curr_frm.camera_->reproject_to_image(rot_cw, trans_cw, lm->get_pos_in_world(), reproj, x_right)) continue;  // :135-142
        const int level = last_kps[idx_last].octave;
        int lo = std::max(0, level - 1), hi = std::min(num_levels - 1, level + 1);  // :145-158
        if (assume_forward) lo = level;
        else if (assume_backward) hi = level;
        lms[idx_last] = lm;
        Q.valid[idx_last] = 1;
        Q.x[idx_last] = static_cast<float>(reproj(0));
        Q.y[idx_last] = static_cast<float>(reproj(1));
        Q.margin[idx_last] = margin * curr_frm.orb_params_->scale_factors_.at(level);
        Q.lo[idx_last] = static_cast<int8_t>(lo);
        Q.hi[idx_last] = static_cast<int8_t>(hi);
        Q.x_right[idx_last] = x_right;
        Q.angle[idx_last] = last_kps[idx_last].angle;
        const cv::Mat d = lm->get_descriptor();
        std::memcpy(&Q.desc[32 * idx_last], d.data, 32);
    }
    Q.bind(P);
    if (b200_match_guided(matcher(), 1, &P, B200_GUIDED_LAST_FRAME, HAMMING_DIST_THR_HIGH, lowe_ratio_, check_orientation_ ? 1 : 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    for (unsigned int idx_last = 0; idx_last < last_kps.size(); ++idx_last)
        if (Q.out[idx_last] >= 0) curr_frm.add_landmark(lms[idx_last], static_cast<unsigned int>(Q.out[idx_last]));  // :202
    return static_cast<unsigned int>(P.n_matches);
}

}  // namespace match
}  // namespace stella_vslam
