// Replacement body for match::fuse::detect_duplication (src/stella_vslam/match/fuse.cc:12-154), all three instantiations.
// Guard the original with #ifndef USE_B200 and add this TU to src/stella_vslam/match/CMakeLists.txt.  The map-side walk (which
// landmarks, visibility, scale range, viewing angle, level window) stays here in the container's iteration order; the grid search,
// the chi-square reprojection gate, the distances and the "a keypoint is fused once" state run in b200_match_guided (mode FUSE).
#include <cstring>
#include <stdexcept>
#include <unordered_set>

#include "stella_vslam/camera/base.h"
#include "stella_vslam/data/keyframe.h"
#include "stella_vslam/data/landmark.h"
#include "stella_vslam/match/fuse.h"

#include "b200vslam.h"

namespace stella_vslam {
namespace match {

template<typename T>
unsigned int fuse::detect_duplication(const std::shared_ptr<data::keyframe>& keyfrm, const Mat33_t& rot_cw, const Vec3_t& trans_cw,
                                      const T& landmarks_to_check, const float margin,
                                      std::unordered_map<std::shared_ptr<data::landmark>, std::shared_ptr<data::landmark>>& duplicated_lms_in_keyfrm,
                                      std::unordered_map<unsigned int, std::shared_ptr<data::landmark>>& new_connections,
                                      bool do_reprojection_matching) const {
    static thread_local b200_matcher_t h = nullptr;
    if (!h && b200_matcher_create(0, &h) != B200_OK) throw std::runtime_error(b200_last_error());
    const Vec3_t trans_wc = -rot_cw.transpose() * trans_cw;
    duplicated_lms_in_keyfrm.clear();
    const auto& obs = keyfrm->frm_obs_;
    const auto* prm = keyfrm->orb_params_;
    const size_t n = obs.undist_keypts_.size();
    std::vector<float> tx(n), ty(n);
    std::vector<uint8_t> toct(n);
    for (size_t i = 0; i < n; ++i) {
        tx[i] = obs.undist_keypts_[i].pt.x;
        ty[i] = obs.undist_keypts_[i].pt.y;
        toct[i] = static_cast<uint8_t>(obs.undist_keypts_[i].octave);
    }
    // queries in the container's order (vector / id_ordered_set / unordered_set: whatever order the reference's range-for sees)
    std::vector<std::shared_ptr<data::landmark>> lms(landmarks_to_check.begin(), landmarks_to_check.end());
    const size_t nq = lms.size();
    std::vector<uint8_t> desc(32 * nq, 0), valid(nq, 0);
    std::vector<float> qx(nq, 0.f), qy(nq, 0.f), qm(nq, 0.f), qxr(nq, 0.f);
    std::vector<int8_t> lo(nq, -1), hi(nq, -1);
    std::vector<double> reproj2(2 * nq, 0.0);
    std::vector<int32_t> out(nq, -1);
    for (size_t q = 0; q < nq; ++q) {
        const auto& lm = lms[q];
        if (!lm || lm->will_be_erased() || lm->is_observed_in_keyframe(keyfrm)) continue;  // :27-37
        const Vec3_t pos_w = lm->get_pos_in_world();
        Vec2_t reproj;
        float x_right;
        if (!keyfrm->camera_->reproject_to_image(rot_cw, trans_cw, pos_w, reproj, x_right)) continue;  // :42-50
        const Vec3_t cam_to_lm_vec = pos_w - trans_wc;
        const auto cam_to_lm_dist = cam_to_lm_vec.norm();
        const auto margin_far = 1.3;
        const auto margin_near = 1.0 / margin_far;
        if (cam_to_lm_dist < margin_near * lm->get_min_valid_distance() || margin_far * lm->get_max_valid_distance() < cam_to_lm_dist) continue;  // :53-61
        if (cam_to_lm_vec.dot(lm->get_obs_mean_normal()) < 0.5 * cam_to_lm_dist) continue;  // :63-69
        const auto level = lm->predict_scale_level(cam_to_lm_dist, prm->num_levels_, prm->log_scale_factor_);
        valid[q] = 1;
        qx[q] = static_cast<float>(reproj(0));
        qy[q] = static_cast<float>(reproj(1));
        reproj2[2 * q] = reproj(0);
        reproj2[2 * q + 1] = reproj(1);
        qxr[q] = x_right;
        qm[q] = margin * prm->scale_factors_.at(level);
        lo[q] = static_cast<int8_t>(std::max(0, static_cast<int>(level) - 1));
        hi[q] = static_cast<int8_t>(std::min(static_cast<int>(prm->num_levels_) - 1, static_cast<int>(level) + 1));
        const cv::Mat d = lm->get_descriptor();
        std::memcpy(&desc[32 * q], d.data, 32);
    }
    b200_guided_problem_t P{};
    P.n_train = static_cast<int32_t>(n);
    P.t_x = tx.data(); P.t_y = ty.data(); P.t_octave = toct.data();
    P.t_x_right = obs.stereo_x_right_.empty() ? nullptr : obs.stereo_x_right_.data();
    P.t_desc = obs.descriptors_.data;
    const auto& b = keyfrm->camera_->img_bounds_;
    P.min_x = b.min_x_; P.max_x = b.max_x_; P.min_y = b.min_y_; P.max_y = b.max_y_;
    P.grid_cols = static_cast<int32_t>(obs.num_grid_cols_);
    P.grid_rows = static_cast<int32_t>(obs.num_grid_rows_);
    P.n_queries = static_cast<int32_t>(nq);
    P.q_desc = desc.data(); P.q_x = qx.data(); P.q_y = qy.data(); P.q_margin = qm.data(); P.q_min_level = lo.data(); P.q_max_level = hi.data();
    P.q_x_right = qxr.data(); P.q_valid = valid.data(); P.q_reproj = reproj2.data();
    P.inv_level_sigma_sq = prm->inv_level_sigma_sq_.data();
    P.n_levels = static_cast<int32_t>(prm->inv_level_sigma_sq_.size());
    P.do_reprojection_matching = do_reprojection_matching ? 1 : 0;
    P.match_out = out.data();
    if (b200_match_guided(h, 1, &P, B200_GUIDED_FUSE, HAMMING_DIST_THR_LOW, lowe_ratio_, 0, 0) != B200_OK) throw std::runtime_error(b200_last_error());
    for (size_t q = 0; q < nq; ++q) {  // :131-146, in query order like the reference's loop
        if (out[q] < 0) continue;
        auto lm_in_keyfrm = keyfrm->get_landmark(static_cast<unsigned int>(out[q]));
        if (lm_in_keyfrm) {
            if (!lm_in_keyfrm->will_be_erased()) duplicated_lms_in_keyfrm[lms[q]] = lm_in_keyfrm;
        }
        else {
            new_connections.emplace(static_cast<unsigned int>(out[q]), lms[q]);
        }
    }
    return static_cast<unsigned int>(P.n_matches);
}

// the reference's explicit instantiations (fuse.cc:156-184)
template unsigned int fuse::detect_duplication(const std::shared_ptr<data::keyframe>&, const Mat33_t&, const Vec3_t&,
                                               const std::vector<std::shared_ptr<data::landmark>>&, const float,
                                               std::unordered_map<std::shared_ptr<data::landmark>, std::shared_ptr<data::landmark>>&,
                                               std::unordered_map<unsigned int, std::shared_ptr<data::landmark>>&, bool) const;
template unsigned int fuse::detect_duplication(const std::shared_ptr<data::keyframe>&, const Mat33_t&, const Vec3_t&,
                                               const id_ordered_set<std::shared_ptr<data::landmark>>&, const float,
                                               std::unordered_map<std::shared_ptr<data::landmark>, std::shared_ptr<data::landmark>>&,
                                               std::unordered_map<unsigned int, std::shared_ptr<data::landmark>>&, bool) const;
template unsigned int fuse::detect_duplication(const std::shared_ptr<data::keyframe>&, const Mat33_t&, const Vec3_t&,
                                               const std::unordered_set<std::shared_ptr<data::landmark>>&, const float,
                                               std::unordered_map<std::shared_ptr<data::landmark>, std::shared_ptr<data::landmark>>&,
                                               std::unordered_map<unsigned int, std::shared_ptr<data::landmark>>&, bool) const;

}  // namespace match
}  // namespace stella_vslam
