// Replacement bodies for the two per-frame guided matchers
//   match::projection::match_frame_and_landmarks       (src/stella_vslam/match/projection.cc:13-93)
//   match::projection::match_current_and_last_frames   (src/stella_vslam/match/projection.cc:95-207)
// Guard the originals with #ifndef USE_B200 and add this TU to src/stella_vslam/match/CMakeLists.txt.  The map-side walk
// (which landmarks, their reprojection, the level window, validity) stays here, in the reference's order; the grid search, the
// gates, the distances and the sequential "a keypoint takes one landmark" state run in b200_match_guided.
// plus the occasional variants of the same class:
//   match::projection::match_frame_and_keyframe (both overloads) (:209-319), match_by_Sim3_transform (:321-416),
//   match_keyframes_mutually (:418-630)
// fuse::detect_duplication and area::match_in_consistent_area live in fuse_b200.cc / area_b200.cc.
#include <cstring>
#include <stdexcept>

#include "stella_vslam/camera/base.h"
#include "stella_vslam/data/frame.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"
#include "stella_vslam/match/projection.h"

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
    // occupied_at(i): the keypoint is closed to the landmarks of this call before it starts
    template <class Occ>
    void fill(const data::frame_observation& obs, const camera::base* camera, bool with_x_right, Occ occupied_at, b200_guided_problem_t& P) {
        const auto& kps = obs.undist_keypts_;
        const size_t n = kps.size();
        x.resize(n); y.resize(n); angle.resize(n); octave.resize(n); occupied.resize(n);
        for (size_t i = 0; i < n; ++i) {
            x[i] = kps[i].pt.x; y[i] = kps[i].pt.y; angle[i] = kps[i].angle; octave[i] = static_cast<uint8_t>(kps[i].octave);
            occupied[i] = occupied_at(i) ? 1 : 0;
        }
        P.n_train = static_cast<int32_t>(n);
        P.t_x = x.data(); P.t_y = y.data(); P.t_octave = octave.data(); P.t_angle = angle.data();
        P.t_x_right = (!with_x_right || obs.stereo_x_right_.empty()) ? nullptr : obs.stereo_x_right_.data();
        P.t_desc = obs.descriptors_.data;
        P.t_occupied = occupied.data();
        const auto& b = camera->img_bounds_;
        P.min_x = b.min_x_; P.max_x = b.max_x_; P.min_y = b.min_y_; P.max_y = b.max_y_;
        P.grid_cols = static_cast<int32_t>(obs.num_grid_cols_);
        P.grid_rows = static_cast<int32_t>(obs.num_grid_rows_);
    }
    void fill(const data::frame& frm, b200_guided_problem_t& P) {
        fill(frm.frm_obs_, frm.camera_, true,
             [&](size_t i) {
                 const auto& lm = frm.get_landmark(i);
                 return lm && lm->has_observation();  // projection.cc:50-53, 163-166
             },
             P);
    }
};

struct query_side {
    std::vector<uint8_t> desc, valid, has_obs;
    std::vector<float> x, y, margin, x_right, angle;
    std::vector<int8_t> lo, hi;
    std::vector<int32_t> out;
    void resize(size_t n) {
        desc.assign(32 * n, 0); valid.assign(n, 0); has_obs.assign(n, 1); x.assign(n, 0.f); y.assign(n, 0.f); margin.assign(n, 0.f); x_right.assign(n, 0.f);
        angle.assign(n, 0.f); lo.assign(n, -1); hi.assign(n, -1); out.assign(n, -1);
    }
    void bind(b200_guided_problem_t& P) {
        P.n_queries = static_cast<int32_t>(valid.size());
        P.q_desc = desc.data(); P.q_x = x.data(); P.q_y = y.data(); P.q_margin = margin.data(); P.q_min_level = lo.data(); P.q_max_level = hi.data();
        P.q_x_right = x_right.data(); P.q_angle = angle.data(); P.q_valid = valid.data(); P.match_out = out.data();
        P.q_has_observation = has_obs.data();  // projection.cc:50-53, 163-166: only a landmark WITH observations closes its keypoint
    }
    // the part every landmark-projecting matcher repeats (projection.cc:250-281, 345-381, 466-497; fuse.cc:36-77): visibility, ORB
    // scale range, predicted level window.  Returns false when the reference `continue`s.
    bool project(size_t q, const data::landmark& lm, const camera::base* camera, const feature::orb_params* orb_params, const Mat33_t& rot_cw,
                 const Vec3_t& trans_cw, const Vec3_t& cam_center, float margin_px, bool check_normal, Vec2_t* reproj_out = nullptr) {
        const Vec3_t pos_w = lm.get_pos_in_world();
        Vec2_t reproj;
        float xr;
        if (!camera->reproject_to_image(rot_cw, trans_cw, pos_w, reproj, xr)) return false;
        const Vec3_t cam_to_lm_vec = pos_w - cam_center;
        const auto cam_to_lm_dist = cam_to_lm_vec.norm();
        constexpr auto margin_far = 1.3;
        constexpr auto margin_near = 1.0 / margin_far;
        if (cam_to_lm_dist < margin_near * lm.get_min_valid_distance() || margin_far * lm.get_max_valid_distance() < cam_to_lm_dist) return false;
        if (check_normal && cam_to_lm_vec.dot(lm.get_obs_mean_normal()) < 0.5 * cam_to_lm_dist) return false;
        const auto level = lm.predict_scale_level(cam_to_lm_dist, orb_params->num_levels_, orb_params->log_scale_factor_);
        valid[q] = 1;
        x[q] = static_cast<float>(reproj(0));
        y[q] = static_cast<float>(reproj(1));
        x_right[q] = xr;
        margin[q] = margin_px * orb_params->scale_factors_.at(level);
        lo[q] = static_cast<int8_t>(std::max(0, static_cast<int>(level) - 1));
        hi[q] = static_cast<int8_t>(std::min(static_cast<int>(orb_params->num_levels_) - 1, static_cast<int>(level) + 1));
        const cv::Mat d = lm.get_descriptor();
        std::memcpy(&desc[32 * q], d.data, 32);
        if (reproj_out) *reproj_out = reproj;
        return true;
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
        Q.has_obs[q] = lm->has_observation() ? 1 : 0;
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
        Q.has_obs[idx_last] = lm->has_observation() ? 1 : 0;  // temporal landmarks of a stereo / RGBD last frame have none
    }
    Q.bind(P);
    if (b200_match_guided(matcher(), 1, &P, B200_GUIDED_LAST_FRAME, HAMMING_DIST_THR_HIGH, lowe_ratio_, check_orientation_ ? 1 : 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    for (unsigned int idx_last = 0; idx_last < last_kps.size(); ++idx_last)
        if (Q.out[idx_last] >= 0) curr_frm.add_landmark(lms[idx_last], static_cast<unsigned int>(Q.out[idx_last]));  // :202
    return static_cast<unsigned int>(P.n_matches);
}

// projection.cc:209-215
unsigned int projection::match_frame_and_keyframe(data::frame& curr_frm, const std::shared_ptr<data::keyframe>& keyfrm,
                                                  const std::set<std::shared_ptr<data::landmark>>& already_matched_lms, const float margin,
                                                  const unsigned int hamm_dist_thr) const {
    auto lms = curr_frm.get_landmarks();
    auto num_matches = match_frame_and_keyframe(curr_frm.get_pose_cw(), curr_frm.camera_, curr_frm.frm_obs_, curr_frm.orb_params_, lms, keyfrm,
                                                already_matched_lms, margin, hamm_dist_thr);
    curr_frm.set_landmarks(lms);
    return num_matches;
}

// projection.cc:217-319: the keyframe's landmarks projected into the frame; a frame keypoint that already carries ANY landmark is closed
// (:292-294), no stereo gate, orientation between the keyframe keypoint and the frame keypoint (:296-298)
unsigned int projection::match_frame_and_keyframe(const Mat44_t& cam_pose_cw, const camera::base* camera, const data::frame_observation& frm_obs,
                                                  const feature::orb_params* orb_params, std::vector<std::shared_ptr<data::landmark>>& frm_landmarks,
                                                  const std::shared_ptr<data::keyframe>& keyfrm,
                                                  const std::set<std::shared_ptr<data::landmark>>& already_matched_lms, const float margin,
                                                  const unsigned int hamm_dist_thr) const {
    const Mat33_t rot_cw = cam_pose_cw.block<3, 3>(0, 0);
    const Vec3_t trans_cw = cam_pose_cw.block<3, 1>(0, 3);
    const Vec3_t cam_center = -rot_cw.transpose() * trans_cw;
    const auto landmarks = keyfrm->get_landmarks();
    b200_guided_problem_t P{};
    train_side T;
    T.fill(frm_obs, camera, false, [&](size_t i) { return static_cast<bool>(frm_landmarks.at(i)); }, P);
    query_side Q;
    Q.resize(landmarks.size());
    for (unsigned int idx = 0; idx < landmarks.size(); ++idx) {
        const auto& lm = landmarks.at(idx);
        if (!lm || lm->will_be_erased() || already_matched_lms.count(lm)) continue;  // :239-250
        if (!Q.project(idx, *lm, camera, orb_params, rot_cw, trans_cw, cam_center, margin, false)) continue;
        Q.angle[idx] = keyfrm->frm_obs_.undist_keypts_.at(idx).angle;
    }
    Q.bind(P);
    P.q_has_observation = nullptr;  // any assigned landmark closes the keypoint (:292-294, 315)
    if (b200_match_guided(matcher(), 1, &P, B200_GUIDED_LAST_FRAME, hamm_dist_thr, lowe_ratio_, check_orientation_ ? 1 : 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    for (unsigned int idx = 0; idx < landmarks.size(); ++idx)
        if (Q.out[idx] >= 0) frm_landmarks.at(Q.out[idx]) = landmarks.at(idx);  // :315
    return static_cast<unsigned int>(P.n_matches);
}

// projection.cc:321-416
unsigned int projection::match_by_Sim3_transform(const std::shared_ptr<data::keyframe>& keyfrm, const Mat44_t& Sim3_cw,
                                                 const std::vector<std::shared_ptr<data::landmark>>& landmarks,
                                                 std::vector<std::shared_ptr<data::landmark>>& matched_lms_in_keyfrm, const float margin) const {
    const Mat33_t s_rot_cw = Sim3_cw.block<3, 3>(0, 0);
    const auto s_cw = std::sqrt(s_rot_cw.block<1, 3>(0, 0).dot(s_rot_cw.block<1, 3>(0, 0)));
    const Mat33_t rot_cw = s_rot_cw / s_cw;
    const Vec3_t trans_cw = Sim3_cw.block<3, 1>(0, 3) / s_cw;
    const Vec3_t cam_center = -rot_cw.transpose() * trans_cw;
    std::set<std::shared_ptr<data::landmark>> already_matched(matched_lms_in_keyfrm.begin(), matched_lms_in_keyfrm.end());
    already_matched.erase(nullptr);
    b200_guided_problem_t P{};
    train_side T;
    T.fill(keyfrm->frm_obs_, keyfrm->camera_, false, [&](size_t i) { return static_cast<bool>(matched_lms_in_keyfrm.at(i)); }, P);
    query_side Q;
    Q.resize(landmarks.size());
    for (size_t q = 0; q < landmarks.size(); ++q) {
        const auto& lm = landmarks[q];
        if (lm->will_be_erased() || already_matched.count(lm)) continue;  // :337-343
        Q.project(q, *lm, keyfrm->camera_, keyfrm->orb_params_, rot_cw, trans_cw, cam_center, margin, true);
    }
    Q.bind(P);
    P.q_has_observation = nullptr;
    if (b200_match_guided(matcher(), 1, &P, B200_GUIDED_LAST_FRAME, HAMMING_DIST_THR_LOW, lowe_ratio_, 0, 0) != B200_OK)  // no orientation test here
        throw std::runtime_error(b200_last_error());
    for (size_t q = 0; q < landmarks.size(); ++q)
        if (Q.out[q] >= 0) matched_lms_in_keyfrm.at(Q.out[q]) = landmarks[q];  // :411
    return static_cast<unsigned int>(P.n_matches);
}

// projection.cc:418-630: both directions are stateless searches (one call, two problems), then the mutual check
unsigned int projection::match_keyframes_mutually(const std::shared_ptr<data::keyframe>& keyfrm_1, const std::shared_ptr<data::keyframe>& keyfrm_2,
                                                  std::vector<std::shared_ptr<data::landmark>>& matched_lms_in_keyfrm_1, const float& s_12,
                                                  const Mat33_t& rot_12, const Vec3_t& trans_12, const float margin) const {
    const Mat33_t rot_1w = keyfrm_1->get_rot_cw();
    const Vec3_t trans_1w = keyfrm_1->get_trans_cw();
    const Mat33_t rot_2w = keyfrm_2->get_rot_cw();
    const Vec3_t trans_2w = keyfrm_2->get_trans_cw();
    const Mat33_t s_rot_12 = s_12 * rot_12;
    const Mat33_t s_rot_21 = (1.0 / s_12) * rot_12.transpose();
    const Vec3_t trans_21 = -s_rot_21 * trans_12;
    const auto landmarks_1 = keyfrm_1->get_landmarks();
    const auto landmarks_2 = keyfrm_2->get_landmarks();
    std::vector<bool> matched_1(landmarks_1.size(), false), matched_2(landmarks_2.size(), false);
    for (unsigned int idx_1 = 0; idx_1 < landmarks_1.size(); ++idx_1) {  // :436-448
        const auto& lm = matched_lms_in_keyfrm_1.at(idx_1);
        if (!lm) continue;
        const auto idx_2 = lm->get_index_in_keyframe(keyfrm_2);
        if (0 <= idx_2 && idx_2 < static_cast<int>(landmarks_2.size())) {
            matched_1.at(idx_1) = true;
            matched_2.at(idx_2) = true;
        }
    }
    b200_guided_problem_t P[2] = {};
    train_side T[2];
    query_side Q[2];
    // direction 0: landmarks of keyframe 1 searched in keyframe 2 (:453-525); direction 1: the converse (:527-600).  As in the reference,
    // BOTH directions project with keyfrm_2->camera_ (:474, :550) and take the distance from the Sim3-transformed point itself
    const Mat33_t s_rot[2] = {s_rot_21 * rot_1w, s_rot_12 * rot_2w};
    const Vec3_t trans[2] = {s_rot_21 * trans_1w + trans_21, s_rot_12 * trans_2w + trans_12};
    const std::shared_ptr<data::keyframe> target[2] = {keyfrm_2, keyfrm_1};
    const std::vector<std::shared_ptr<data::landmark>>* lms[2] = {&landmarks_1, &landmarks_2};
    const std::vector<bool>* done[2] = {&matched_1, &matched_2};
    for (int dir = 0; dir < 2; ++dir) {
        const auto& kf = target[dir];
        T[dir].fill(kf->frm_obs_, kf->camera_, false, [](size_t) { return false; }, P[dir]);
        P[dir].t_occupied = nullptr;
        Q[dir].resize(lms[dir]->size());
        for (unsigned int idx = 0; idx < lms[dir]->size(); ++idx) {
            const auto& lm = lms[dir]->at(idx);
            if (!lm || lm->will_be_erased() || done[dir]->at(idx)) continue;
            const Vec3_t pos_w = lm->get_pos_in_world();
            const Vec3_t pos_c = s_rot[dir] * pos_w + trans[dir];
            Vec2_t reproj;
            float x_right;
            if (!keyfrm_2->camera_->reproject_to_image(s_rot[dir], trans[dir], pos_w, reproj, x_right)) continue;
            const auto cam_to_lm_dist = pos_c.norm();
            constexpr auto margin_far = 1.3;
            constexpr auto margin_near = 1.0 / margin_far;
            if (cam_to_lm_dist < margin_near * lm->get_min_valid_distance() || margin_far * lm->get_max_valid_distance() < cam_to_lm_dist) continue;
            const auto level = lm->predict_scale_level(cam_to_lm_dist, kf->orb_params_->num_levels_, kf->orb_params_->log_scale_factor_);
            auto& q = Q[dir];
            q.valid[idx] = 1;
            q.x[idx] = static_cast<float>(reproj(0));
            q.y[idx] = static_cast<float>(reproj(1));
            q.margin[idx] = margin * kf->orb_params_->scale_factors_.at(level);
            q.lo[idx] = static_cast<int8_t>(std::max(0, static_cast<int>(level) - 1));
            q.hi[idx] = static_cast<int8_t>(std::min(static_cast<int>(kf->orb_params_->num_levels_) - 1, static_cast<int>(level) + 1));
            const cv::Mat d = lm->get_descriptor();
            std::memcpy(&q.desc[32 * idx], d.data, 32);
        }
        Q[dir].bind(P[dir]);
        P[dir].q_has_observation = nullptr;
    }
    if (b200_match_guided(matcher(), 2, P, B200_GUIDED_INDEPENDENT, HAMMING_DIST_THR_HIGH, lowe_ratio_, 0, 0) != B200_OK)
        throw std::runtime_error(b200_last_error());
    std::vector<int32_t> mutual(landmarks_1.size(), -1);
    int32_t n_mutual = 0;
    b200_match_cross_check(Q[0].out.data(), static_cast<int>(landmarks_1.size()), Q[1].out.data(), static_cast<int>(landmarks_2.size()), mutual.data(),
                           &n_mutual);
    for (unsigned int i = 0; i < landmarks_1.size(); ++i)
        if (mutual[i] >= 0) matched_lms_in_keyfrm_1.at(i) = landmarks_2.at(mutual[i]);  // :614-627
    return static_cast<unsigned int>(n_mutual);
}

}  // namespace match
}  // namespace stella_vslam
